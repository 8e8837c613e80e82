"""CPU: the index work of echo_graph_create (host side, exported as echo_debug_graph_csr) -- bit-exact against a restatement
of the order in which the reference's scatter_add visits the (triple, role) items of every node (model/graph.py:161-199:
`pooled.scatter_add(0, s_idx_exp, new_s_vecs)` then `pooled.scatter_add(0, o_idx_exp, new_o_vecs)`), on the synthetic scene
graphs, on random multigraphs with self loops / duplicates / isolated nodes, and on the empty graph."""
import ctypes as C

import numpy as np
import pytest
import torch

from echoscene_b200 import _lib, synth


def csr(triples: torch.Tensor, n: int):
    t = int(triples.shape[0])
    tri = np.ascontiguousarray(triples.numpy().astype(np.int64)).reshape(t, 3)
    off = np.full(n + 1, -7, dtype=np.int32)
    items = np.full(max(2 * t, 1), -7, dtype=np.int32)
    rng = np.zeros(2, dtype=np.int64)
    rc = _lib.lib().echo_debug_graph_csr(tri.ctypes.data_as(C.c_void_p) if t else None, t, n, off.ctypes.data_as(C.c_void_p),
                                         items.ctypes.data_as(C.c_void_p), rng.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise _lib.EchoError(_lib.lib().echo_last_error().decode())
    return off, items[:2 * t], rng


def visit_order(triples: torch.Tensor, n: int):
    """items of node v in scatter_add order: every triple whose SUBJECT is v by ascending t, then every triple whose OBJECT is v"""
    per = [[] for _ in range(n)]
    for t, (s, _, o) in enumerate(triples.tolist()):
        per[s].append(2 * t)
    for t, (s, _, o) in enumerate(triples.tolist()):
        per[o].append(2 * t + 1)
    off = np.cumsum([0] + [len(p) for p in per]).astype(np.int32)
    return off, np.asarray([i for p in per for i in p], dtype=np.int32)


@pytest.mark.parametrize("n,t,seed", [(8, 32, 1), (16, 64, 2), (32, 128, 3), (2, 1, 4)])
def test_csr_of_synthetic_scene_graphs(n, t, seed):
    g = synth.make_scene_graph(n, t, seed)
    off, items, rng = csr(g.triples, n)
    w_off, w_items = visit_order(g.triples, n)
    assert np.array_equal(off, w_off) and np.array_equal(items, w_items)
    assert rng.tolist() == [int(g.triples[:, 1].min()), int(g.triples[:, 1].max())]
    assert off[-1] == 2 * t                                    # every triple contributes a subject and an object item


@pytest.mark.parametrize("seed", range(6))
def test_csr_of_random_multigraphs(seed):
    gen = torch.Generator().manual_seed(seed)
    n = int(torch.randint(1, 40, (1,), generator=gen))
    t = int(torch.randint(0, 200, (1,), generator=gen))
    tri = torch.stack([torch.randint(0, n, (t,), generator=gen), torch.randint(0, 16, (t,), generator=gen),
                       torch.randint(0, max(n // 2, 1), (t,), generator=gen)], dim=1)      # nodes >= n/2 are never objects
    if t >= 3:
        tri[0, 2] = tri[0, 0]                                   # self loop: counted as subject AND object (graph.py:191-192)
        tri[2] = tri[1]                                         # duplicate edge
    off, items, rng = csr(tri, n)
    w_off, w_items = visit_order(tri, n)
    assert np.array_equal(off, w_off) and np.array_equal(items, w_items)
    # degree = what the reference divides by (clamped to 1 on the device): ones scattered at s and at o (graph.py:188-195)
    deg = torch.zeros(n).scatter_add(0, tri[:, 0], torch.ones(t)).scatter_add(0, tri[:, 2], torch.ones(t))
    assert np.array_equal(np.diff(off), deg.numpy().astype(np.int32))


def test_csr_edge_cases():
    off, items, rng = csr(torch.zeros(0, 3, dtype=torch.int64), 5)
    assert off.tolist() == [0] * 6 and items.size == 0 and rng.tolist() == [0, -1]
    off, items, rng = csr(torch.zeros(0, 3, dtype=torch.int64), 0)
    assert off.tolist() == [0]
    with pytest.raises(_lib.EchoError, match="out of range"):
        csr(torch.tensor([[0, 1, 5]]), 5)
    with pytest.raises(_lib.EchoError, match="out of range"):
        csr(torch.tensor([[-1, 1, 0]]), 5)
    _, _, rng = csr(torch.tensor([[0, -3, 1], [1, 99, 0]]), 2)          # predicate ids are only recorded; users of the graph check them
    assert rng.tolist() == [-3, 99]
