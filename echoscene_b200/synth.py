"""Synthetic scene graphs / conditioning of the BASELINE.json configs.

Node layout follows dataset/threedfront_dataset.py:339-350: the last node is ``_scene_`` and every
other node has an "in" edge (predicate 0) to it; remaining triples are random [s, p in 1..15, o], s != o.
All draws come from a seeded CPU ``torch.Generator`` so the GPU box and this container agree.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class SceneGraph:
    n_nodes: int
    triples: torch.Tensor        # (T, 3) int64 [s, p, o]


def make_scene_graph(n_nodes: int, n_triples: int, seed: int) -> SceneGraph:
    g = torch.Generator().manual_seed(seed)
    n_obj = n_nodes - 1
    rows = [[i, 0, n_obj] for i in range(n_obj)]
    if n_obj < 2 and n_triples > len(rows):
        raise ValueError("need >= 2 objects for random (s != o) triples")
    while len(rows) < n_triples:
        s = int(torch.randint(0, n_obj, (1,), generator=g))
        o = int(torch.randint(0, n_obj, (1,), generator=g))
        if s == o:
            continue
        p = int(torch.randint(1, 16, (1,), generator=g))
        rows.append([s, p, o])
    return SceneGraph(n_nodes, torch.tensor(rows[:n_triples], dtype=torch.int64))


def batch_scene_graphs(graphs) -> SceneGraph:
    """Disjoint union with index offsets, as collate_fn does (threedfront_dataset.py:698-701)."""
    off, tri = 0, []
    for g in graphs:
        t = g.triples.clone()
        t[:, 0] += off
        t[:, 2] += off
        tri.append(t)
        off += g.n_nodes
    return SceneGraph(off, torch.cat(tri, 0))


def shape_inputs(n_nodes: int, seed: int, context_dim: int = 1280, latent=(3, 16, 16, 16), same_noise: bool = True):
    """uc_s (N,1,ctx) and x_T; rel2shape repeats ONE noise draw for every object (echo2shape.py:507-510)."""
    g = torch.Generator().manual_seed(seed)
    uc = torch.randn(n_nodes, 1, context_dim, generator=g)
    if same_noise:
        x_T = torch.randn(1, *latent, generator=g).repeat(n_nodes, 1, 1, 1, 1)
    else:
        x_T = torch.randn(n_nodes, *latent, generator=g)
    return uc, x_T.contiguous()


def layout_inputs(n_nodes: int, seed: int, obj_dim: int = 640, box_dim: int = 8):
    g = torch.Generator().manual_seed(seed)
    obj_embed = torch.randn(n_nodes, obj_dim, generator=g)
    x_T = torch.randn(n_nodes, box_dim, generator=g)
    return obj_embed, x_T


# ---- workload configuration of bench.py / smoke() (product side: nothing here touches oracle/) -----------------------
WEIGHT_SEED_GCN = 10
WEIGHT_SEED_LAYOUT = 11
WEIGHT_SEED_SHAPE = 12
WEIGHT_SEED_VQVAE = 13
WEIGHT_SEED_SCENE = 14


def layout_cfg():
    from . import arch
    return arch.UNet1DConfig()


def shape_cfg():
    from . import arch
    return arch.UNet3DConfig()


def vqvae_cfg():
    from . import arch
    return arch.VQVAEConfig()


def scene_cfg():
    from . import arch
    return arch.SceneEncoderConfig()


def vqvae_inputs(n: int = 2, seed: int = 6):
    """latents as the DDIM chain leaves them: (n, 3, 16, 16, 16), O(1) values"""
    gen = torch.Generator().manual_seed(seed + 500)
    return torch.randn(n, 3, 16, 16, 16, generator=gen)


def scene_inputs(n_nodes: int, n_triples: int, seed: int, cfg=None):
    """dec_objs (N,) i64 class ids (last node = '_scene_' class 0), triples, CLIP-like text / relation features"""
    cfg = cfg or scene_cfg()
    g = make_scene_graph(n_nodes, n_triples, seed)
    gen = torch.Generator().manual_seed(seed + 600)
    objs = torch.randint(1, cfg.num_objs, (n_nodes,), generator=gen)
    objs[-1] = 0
    text = torch.nn.functional.normalize(torch.randn(n_nodes, cfg.add_dim, generator=gen), dim=1) * 10
    rel = torch.nn.functional.normalize(torch.randn(n_triples, cfg.add_dim, generator=gen), dim=1) * 10
    return g, objs, text, rel
