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


# ---------------------------------------------------------------------------------------------------- shard by scene
def test_partition_scenes_and_subgraph():
    from echoscene_b200 import synth
    sizes = [5, 3, 8, 3, 6, 4]
    graphs = [synth.make_scene_graph(n, n + 2, 10 + i) for i, n in enumerate(sizes)]
    batch = synth.batch_scene_graphs(graphs)
    o2s = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(sizes)])
    for world in (1, 2, 3, 6, 8):
        parts = shard.partition_scenes(o2s, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(sizes) and parts[-1][3] == sum(sizes)
        assert all(a[1] == b[0] and a[3] == b[2] for a, b in zip(parts, parts[1:]))
        if world <= len(sizes):
            assert all(p[1] > p[0] for p in parts)                      # nobody idles while scenes remain
        tri = [shard.scene_subgraph(batch.triples, p[2], p[3]) for p in parts]
        assert sum(t.shape[0] for t in tri) == batch.triples.shape[0]
        for p, t in zip(parts, tri):                                     # each rank's piece is the batch of its own scenes
            own = graphs[p[0]:p[1]]
            if own:
                assert torch.equal(t, synth.batch_scene_graphs(own).triples)
            else:
                assert t.shape == (0, 3)
    assert shard.partition_scenes(o2s, 2) == [(0, 3, 0, 16), (3, 6, 16, 29)]
    assert shard.partition_scenes(torch.zeros(0, dtype=torch.int64), 2) == [(0, 0, 0, 0), (0, 0, 0, 0)]
    with pytest.raises(ValueError):
        shard.partition_scenes(torch.tensor([0, 1, 0]), 2)
    cross = batch.triples.clone()
    cross[0, 2] = sum(sizes) - 1                                          # an edge from scene 0 into the last scene
    with pytest.raises(ValueError):
        shard.scene_subgraph(cross, 0, 16)


def _scene_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from echoscene_b200 import arch, synth
        from oracle import echoscene_oracle as orc
        sizes = [5, 3, 8, 3, 6]
        graphs = [synth.make_scene_graph(n, n + 2, 20 + i) for i, n in enumerate(sizes)]
        batch = synth.batch_scene_graphs(graphs)
        o2s = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(sizes)])
        cfg = arch.GCNConfig(32, 16, 2, 32, 32, True, "avg", "batch")
        sd = arch.make_state_dict(arch.gcn_specs(cfg), 5)
        gen = torch.Generator().manual_seed(1)
        x_all = torch.randn(batch.n_nodes, 32, generator=gen)
        pred_all = torch.randn(batch.triples.shape[0], 16, generator=gen)

        def chain(x, pred, triples):   # the only cross-object op of a step is the echo GCN: iterate it like a chain
            edges, _ = orc.edges_of(triples)
            for _ in range(3):
                x, pred = orc.graph_triple_conv_net(sd, "", x, pred, edges, cfg.num_layers)
            return x

        parts = shard.partition_scenes(o2s, world)
        sb, se, nb, ne = parts[rank]
        tri = shard.scene_subgraph(batch.triples, nb, ne)
        t_in = (batch.triples[:, 0] >= nb) & (batch.triples[:, 0] < ne)
        with torch.no_grad():
            mine = chain(x_all[nb:ne], pred_all[t_in], tri)               # no collective inside the chain
            full = shard.gather_latents(mine, [p[3] - p[2] for p in parts])
            ref = chain(x_all, pred_all, batch.triples)
        ret[rank] = float((full - ref).abs().max() / ref.abs().max())
    finally:
        dist.destroy_process_group()


def test_scene_shard_equals_batched_gloo():
    """BASELINE config 4 shape of the problem (scenes >= ranks): every rank runs its own scenes with no per-step collective;
    the gathered result equals the run over the whole collated batch."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_scene_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-5 for v in ret.values()), dict(ret)


def test_echo_components_restrict_a_block_diagonal_batch():
    """shard.echo_components: the nodes / re-indexed triples of the connected components that contain a rank's objects."""
    from echoscene_b200 import shard, synth
    gs = [synth.make_scene_graph(3, 4, 1), synth.make_scene_graph(4, 6, 2), synth.make_scene_graph(2, 1, 3)]
    b = synth.batch_scene_graphs(gs)
    nodes, tri, begin = shard.echo_components(b.triples, b.n_nodes, 3, 4)          # the rank owns scene 1 = nodes 3..6
    assert nodes.tolist() == [3, 4, 5, 6] and begin == 0
    assert torch.equal(tri, gs[1].triples)                                          # its own triples, re-based, original order
    nodes, tri, begin = shard.echo_components(b.triples, b.n_nodes, 5, 3)          # a range across scenes 1 and 2
    assert nodes.tolist() == [3, 4, 5, 6, 7, 8] and begin == 2
    want = torch.cat([gs[1].triples, gs[2].triples + torch.tensor([4, 0, 4])])
    assert torch.equal(tri, want)
    assert shard.echo_components(b.triples, b.n_nodes, 0, b.n_nodes) is None        # everything is wanted: nothing to restrict
    assert shard.echo_components(gs[1].triples, 4, 0, 2) is None or len(shard.echo_components(gs[1].triples, 4, 0, 2)[0]) <= 4
    # an isolated node (no triples) is its own component
    nodes, tri, begin = shard.echo_components(torch.zeros(0, 3, dtype=torch.int64), 5, 2, 1)
    assert nodes.tolist() == [2] and tri.shape == (0, 3) and begin == 0
