"""TEST INFRASTRUCTURE -- CPU oracle of the SDF -> mesh stage: marching cubes over one (R, R, R) volume, the stage the
reference runs per object on the CPU with PyMCubes (`mcubes.marching_cubes(sdf_i, level)`, model/diff_utils/util_3d.py:213-218,
followed by `verts / n_cell - 0.5`).

PARITY UNPINNED: PyMCubes (requirements.txt: `PyMCubes`, version not pinned) is a third-party C++ extension that is absent from this
image, so neither its outputs nor its case table can be consulted.  What is restated is the published method (Lorensen & Cline 1987)
with case tables derived from its definition by tools/gen_mc_tables.py (the derived edge table equals the classic one entry for
entry); vertices are shared per grid edge as PyMCubes shares them, in index coordinates (x = first array axis), linearly interpolated
along the edge.  Implementations agree on the surface away from ambiguous faces and differ in vertex / triangle ORDER and in the
choice of diagonals -- so the tests pin the CUDA kernels to this oracle exactly, and the oracle to the method through properties
(watertightness, vertices on the level set, area and volume of analytic shapes).  Only tests/ may import this."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_mc_tables as gen   # noqa: E402

EDGE_TABLE, TRI_TABLE, NUM_TRI = gen.tables()
EDGE_BASE = np.array(gen.EDGE_BASE, dtype=np.int64)
CORNERS = gen.CORNERS


def marching_cubes(vol: np.ndarray, level: float):
    """vol (R, R, R) float32 -> (verts (V, 3) float32 in index coordinates, faces (F, 3) int64).
    Order: vertices by grid edge id = direction * R^3 + (i * R + j) * R + k of the edge's base point; triangles by cell (i, j, k)
    in the same linear order, within a cell in table order."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    R = vol.shape[0]
    assert vol.shape == (R, R, R)
    lv = np.float32(level)
    inside = vol < lv
    cut = np.zeros((3, R, R, R), dtype=bool)
    cut[0, :-1] = inside[:-1] != inside[1:]
    cut[1, :, :-1] = inside[:, :-1] != inside[:, 1:]
    cut[2, :, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    flat = cut.reshape(-1)
    vid = np.cumsum(flat, dtype=np.int64) - flat                      # exclusive scan: the vertex index of every cut edge
    ids = np.nonzero(flat)[0]
    d, p = ids // (R ** 3), ids % (R ** 3)
    i, j, k = p // (R * R), (p // R) % R, p % R
    base = np.stack([i, j, k], axis=1)
    step = np.eye(3, dtype=np.int64)[d]
    v0 = vol[i, j, k]
    q = base + step
    v1 = vol[q[:, 0], q[:, 1], q[:, 2]]
    mu = (lv - v0) / (v1 - v0)                                         # float32
    verts = base.astype(np.float32)
    verts[np.arange(len(ids)), d] = verts[np.arange(len(ids)), d] + mu
    # cells
    ci = np.zeros((R - 1, R - 1, R - 1), dtype=np.int64)
    for b, (dx, dy, dz) in enumerate(CORNERS):
        ci |= inside[dx:R - 1 + dx, dy:R - 1 + dy, dz:R - 1 + dz].astype(np.int64) << b
    faces = []
    cells = np.argwhere(NUM_TRI[ci] > 0)                               # lexicographic (i, j, k) = linear order
    vid3 = vid.reshape(3, R, R, R)
    for (a, b, c) in cells:
        case = ci[a, b, c]
        for t in range(NUM_TRI[case]):
            tri = []
            for e in TRI_TABLE[case, 3 * t:3 * t + 3]:
                di, dj, dk, dd = EDGE_BASE[e]
                tri.append(vid3[dd, a + di, b + dj, c + dk])
            faces.append(tri)
    faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    return verts.astype(np.float32), faces


def sdf_to_mesh(sdf: np.ndarray, level: float = 0.02):
    """sdf_to_mesh of the reference without the renderer objects (util_3d.py:194-237): per object marching cubes at `level`, vertices
    scaled to [-0.5, 0.5) by `verts / n_cell - 0.5`.  sdf (N, 1, R, R, R) -> (list of verts, list of faces)."""
    n_cell = sdf.shape[-1]
    vs, fs = [], []
    for i in range(sdf.shape[0]):
        v, f = marching_cubes(sdf[i, 0], level)
        vs.append((v / np.float32(n_cell) - np.float32(0.5)).astype(np.float32))
        fs.append(f)
    return vs, fs


# ---- properties used by the tests -------------------------------------------------------------------------------------------
def edge_use_counts(faces: np.ndarray):
    """how many triangles use each undirected edge, and whether every directed edge is used at most once (consistent orientation)"""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    und = np.sort(e, axis=1)
    _, counts = np.unique(und, axis=0, return_counts=True)
    _, dcounts = np.unique(e, axis=0, return_counts=True)
    return counts, bool((dcounts == 1).all())


def area_and_volume(verts: np.ndarray, faces: np.ndarray):
    a, b, c = (verts[faces[:, k]].astype(np.float64) for k in range(3))
    n = np.cross(b - a, c - a)
    return 0.5 * np.linalg.norm(n, axis=1).sum(), np.einsum("ij,ij->i", a, n).sum() / 6.0
