"""CPU, world_size 2 over gloo: the host-side logic of the per-object shard (partition, ragged all-gather of the shape
codes, sharded step == unsharded step) with the ORACLE standing in for the CUDA callbacks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from echoscene_b200 import shard


def test_partition_is_contiguous_and_balanced():
    for n in (0, 1, 5, 16, 17, 1024):
        for w in (1, 2, 3, 8):
            p = shard.partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))
            sizes = [e - b for b, e in p]
            assert max(sizes) - min(sizes) <= 1
    assert shard.partition(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # a toy "denoiser" with the same dependency structure as the shape step: codes depend on the local latents,
        # the update of object i depends on x_i and on ALL codes (the echo exchange)
        W = torch.randn(3 * 4 * 4 * 4, 64)
        x_all = torch.randn(n, 3, 4, 4, 4)
        ranges = shard.partition(n, world)
        counts = [e - b for b, e in ranges]

        def embed(x):
            return x.reshape(x.shape[0], -1) @ W

        def trunk(x, begin, codes_all):
            ctx = codes_all.mean(0).sum()
            mine = codes_all[begin:begin + x.shape[0]].sum(1).view(-1, 1, 1, 1, 1)
            return x * 0.9 + 0.01 * ctx + 0.001 * mine

        b, e = ranges[rank]
        x = x_all[b:e].clone()
        for _ in range(3):
            x = shard.sharded_shape_step(x, ranges[rank], counts, embed, trunk)
        full = shard.gather_latents(x, counts)
        ref = x_all.clone()
        for _ in range(3):
            ref = trunk(ref, 0, embed(ref))
        ret[rank] = float((full - ref).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [16, 5])
def test_sharded_step_equals_unsharded_gloo(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-5 for v in ret.values()), dict(ret)
