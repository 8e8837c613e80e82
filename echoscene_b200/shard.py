"""Per-object sharding of the shape branch over the GPUs of one box.

Inside a shape step the only cross-object operation is the 5-layer echo GCN over the (N, 1408) node features, of which
only the 64-d shape code depends on the (sharded) latents.  So objects are partitioned contiguously over ranks; per step
each rank (1) runs `shape_embeddings` on its own latents, (2) takes part in ONE all-gather of the (n_local, 64) fp32
codes — the echo exchange, 256 B per object over NVLink, (3) runs the cheap GCN on the connected components of the scene graph
that contain its objects (`echo_components`: message passing never leaves a component, so that is all the echo of its objects
depends on -- one scene of a collated batch, the whole graph of a single scene) and the UNet trunk + DDIM update on its own
objects.  No other collective touches the data path; the final latents are gathered once after the chain.

When a batch holds at least as many scenes as there are GPUs (BASELINE config 4: 64 scenes over 8 GPUs) the shard is by
SCENE instead: the batched graph is block-diagonal per scene (collate_fn offsets, dataset/threedfront_dataset.py:698-701),
so a rank that owns whole scenes needs no per-step collective at all -- `partition_scenes` / `scene_subgraph` cut the batch,
`gather_latents` joins the results once after the chain.

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


def partition_scenes(obj_to_scene: torch.Tensor, world: int) -> List[Tuple[int, int, int, int]]:
    """Contiguous scene ranges per rank for a collated batch (`obj_to_scene` (N,) int64, non-decreasing as collate_fn builds
    it).  Scenes are split so that object counts are as even as contiguous ranges allow (greedy on the running total).
    -> per rank (scene_begin, scene_end, node_begin, node_end); ranks beyond the number of scenes get empty ranges."""
    o2s = obj_to_scene.detach().cpu().to(torch.int64)
    n = int(o2s.numel())
    if n and bool((o2s[1:] < o2s[:-1]).any()):
        raise ValueError("obj_to_scene must be non-decreasing (nodes of a scene are contiguous in a collated batch)")
    n_scenes = int(o2s[-1]) + 1 if n else 0
    sizes = torch.bincount(o2s, minlength=n_scenes).tolist() if n else []
    first = [0]
    for c in sizes:
        first.append(first[-1] + c)
    out, sb = [], 0
    for r in range(world):
        # close this rank's range at the scene boundary nearest to its share of the objects, leaving >= 1 scene per later rank
        target = n * (r + 1) / world
        se = sb
        last_allowed = n_scenes - min(world - 1 - r, max(n_scenes - sb - 1, 0)) if r < world - 1 else n_scenes
        while se < last_allowed and (se == sb or abs(first[se + 1] - target) <= abs(first[se] - target)):
            se += 1
        if r == world - 1:
            se = n_scenes
        out.append((sb, se, first[sb], first[se]))
        sb = se
    return out


def scene_subgraph(triples: torch.Tensor, node_begin: int, node_end: int) -> torch.Tensor:
    """Triples of the scenes whose nodes are [node_begin, node_end) of the batched graph, re-based to local node indices.
    Raises if a triple crosses the boundary (the batch would not be block-diagonal)."""
    s, o = triples[:, 0], triples[:, 2]
    s_in = (s >= node_begin) & (s < node_end)
    o_in = (o >= node_begin) & (o < node_end)
    if bool((s_in != o_in).any()):
        raise ValueError("a triple connects nodes of different ranks' scenes: the batched graph is not block-diagonal")
    local = triples[s_in].clone()
    local[:, 0] -= node_begin
    local[:, 2] -= node_begin
    return local


def echo_components(triples: torch.Tensor, n_nodes: int, obj_begin: int, n_local: int):
    """The echo of a node depends only on its connected component of the scene graph (message passing never leaves it), so a rank
    that owns objects [obj_begin, obj_begin + n_local) needs the GCN on the components those objects belong to and on nothing else
    -- in a collated batch (block-diagonal graph) that is the rank's own scenes instead of the whole batch, for any graph it is
    exact.  -> None when the components cover every node, else (nodes, triples_sub, begin_sub):
      nodes        (n_sub,) int64, ascending: the nodes of those components (on the device of `triples`);
      triples_sub  the triples among them, in their original order, re-indexed into `nodes`;
      begin_sub    where the local range starts inside `nodes` (it stays contiguous: `nodes` is sorted and contains all of it).
    One host round trip (union-find over the edges); callers cache the result per (graph, range)."""
    tri = triples.detach().cpu()
    parent = list(range(n_nodes))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for s, _, o in tri.tolist():
        ra, rb = find(s), find(o)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    wanted = {find(i) for i in range(obj_begin, obj_begin + n_local)}
    nodes = [i for i in range(n_nodes) if find(i) in wanted]
    if len(nodes) == n_nodes:
        return None
    remap = torch.full((n_nodes,), -1, dtype=torch.int64)
    remap[torch.tensor(nodes, dtype=torch.int64)] = torch.arange(len(nodes), dtype=torch.int64)
    if tri.shape[0]:
        keep = remap[tri[:, 0]] >= 0                      # an edge lies inside one component: its object node is kept with it
        sub = tri[keep].clone()
        sub[:, 0], sub[:, 2] = remap[sub[:, 0]], remap[sub[:, 2]]
    else:
        sub = tri.clone()
    dev = triples.device
    return torch.tensor(nodes, dtype=torch.int64, device=dev), sub.to(dev), int(remap[obj_begin])

