"""Per-object sharding of the shape branch over the GPUs of one box.

Inside a shape step the only cross-object operation is the 5-layer echo GCN over the (N, 1408) node features, of which
only the 64-d shape code depends on the (sharded) latents.  So objects are partitioned contiguously over ranks; per step
each rank (1) runs `shape_embeddings` on its own latents, (2) takes part in ONE all-gather of the (n_local, 64) fp32
codes — the echo exchange, 256 B per object over NVLink, (3) runs the cheap GCN redundantly on the whole graph and the
UNet trunk + DDIM update on its own objects.  No other collective touches the data path; the final latents are gathered
once after the chain.

The host logic here is backend-agnostic (NCCL on the GPU box, gloo in the CPU tests): the compute callbacks are passed in.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(n_objects: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [begin, end) object ranges; the first n % world ranks take one more object."""
    base, rem = divmod(n_objects, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def all_gather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather of row blocks with (possibly) different row counts per rank -> (sum(counts), D)."""
    world = dist.get_world_size(group)
    assert len(counts) == world
    if len(set(counts)) == 1:
        out = local.new_empty(sum(counts), *local.shape[1:])
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = local.new_zeros(mx, *local.shape[1:])
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def sharded_shape_step(x_local: torch.Tensor, rank_range: Tuple[int, int], counts: Sequence[int],
                       embed: Callable[[torch.Tensor], torch.Tensor],
                       trunk: Callable[[torch.Tensor, int, torch.Tensor], torch.Tensor], group=None) -> torch.Tensor:
    """One DDIM step on this rank's objects.  embed(x_local) -> (n_local, 64); trunk(x_local, obj_begin, codes_all)."""
    codes_local = embed(x_local)
    codes_all = all_gather_rows(codes_local, counts, group)
    return trunk(x_local, rank_range[0], codes_all)


def gather_latents(x_local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Final all-gather of the (n_local, 3, 16, 16, 16) latents before VQ-VAE decode (48 KB per object)."""
    flat = all_gather_rows(x_local.reshape(x_local.shape[0], -1), counts, group)
    return flat.reshape(-1, *x_local.shape[1:])
