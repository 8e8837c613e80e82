"""Generates the marching-cubes case tables of csrc/mesh.cu (echoscene_b200/csrc/mc_tables.inc) -- and is imported by the oracle
(oracle/mesh_oracle.py), which walks the same tables on the CPU.

The reference meshes its SDFs with PyMCubes (mcubes.marching_cubes, model/diff_utils/util_3d.py:217), a third-party package that is
not in this image and whose 256-row triangle table is not reproducible from memory; the tables here are DERIVED instead, from the
definition of the method (Lorensen & Cline 1987):

  cube corners v0..v7 = (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1); edges e0..e11 = v0v1 v1v2 v2v3 v3v0 v4v5 v5v6
  v6v7 v7v4 v0v4 v1v5 v2v6 v3v7 (the usual numbering); bit i of the case index is set when corner i is INSIDE (value < level).
  On every cube face the intersected edges are joined pairwise; a face whose corners alternate (4 intersected edges) joins the two
  edges next to each INSIDE corner -- a rule that depends on the face's corner signs only, so the two cubes sharing the face agree
  and the surface is watertight.  The joined edges form closed loops around the inside corners; every loop is oriented so that its
  normal points from inside to outside and is cut into a triangle fan.

Same surface as any marching-cubes implementation away from ambiguous faces; on ambiguous faces and in the choice of fan diagonals
implementations differ (PyMCubes' included), which is why parity with it is not claimed ("SDF -> mesh").
Usage: python tools/gen_mc_tables.py   (rewrites echoscene_b200/csrc/mc_tables.inc)"""
import os

import numpy as np

CORNERS = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=np.int64)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}
# local edge -> (offset of its base grid point, direction 0/1/2): the edge runs from base to base + unit(direction)
EDGE_BASE = []
for a, b in EDGES:
    pa, pb = CORNERS[a], CORNERS[b]
    lo = np.minimum(pa, pb)
    EDGE_BASE.append((int(lo[0]), int(lo[1]), int(lo[2]), int(np.argmax(np.abs(pb - pa)))))


def case_triangles(c):
    inside = [(c >> i) & 1 for i in range(8)]
    cut = [inside[a] != inside[b] for a, b in EDGES]
    adj = {i: [] for i in range(12) if cut[i]}
    for f in FACES:
        fe = [EDGE_OF[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)]      # edge k joins corner k and k + 1
        hit = [e for e in fe if cut[e]]
        if len(hit) == 2:
            adj[hit[0]].append(hit[1]); adj[hit[1]].append(hit[0])
        elif len(hit) == 4:
            for k in range(4):                                                    # the two edges at each inside corner k: k - 1 and k
                if inside[f[k]]:
                    a, b = fe[(k - 1) % 4], fe[k]
                    adj[a].append(b); adj[b].append(a)
    assert all(len(v) == 2 for v in adj.values()), (c, adj)
    loops, seen = [], set()
    for s in sorted(adj):
        if s in seen:
            continue
        loop, prev, cur = [s], None, s
        seen.add(s)
        while True:
            nxt = [n for n in adj[cur] if n != prev]
            nxt = nxt[0] if nxt else adj[cur][0]
            if adj[cur][0] == adj[cur][1]:
                nxt = adj[cur][0]
            if nxt == s:
                break
            if nxt in seen:
                break
            loop.append(nxt); seen.add(nxt)
            prev, cur = cur, nxt
        loops.append(loop)
    tris = []
    mid = [(CORNERS[a] + CORNERS[b]) / 2.0 for a, b in EDGES]
    for loop in loops:
        assert len(loop) >= 3, (c, loop)
        pts = np.array([mid[e] for e in loop])
        nrm = np.zeros(3)
        for k in range(len(loop)):                                                # Newell normal
            p, q = pts[k], pts[(k + 1) % len(loop)]
            nrm += np.cross(p, q)
        ins = np.array([CORNERS[a] if inside[a] else CORNERS[b] for a, b in (EDGES[e] for e in loop)], dtype=float).mean(0)
        out = np.array([CORNERS[b] if inside[a] else CORNERS[a] for a, b in (EDGES[e] for e in loop)], dtype=float).mean(0)
        if np.dot(nrm, out - ins) < 0:
            loop = loop[::-1]
        # fan apex: a diagonal whose two ends lie on one cube face would coincide with a segment (or diagonal) of the neighbouring
        # cube and make a non-manifold edge; take the first rotation of the loop whose fan has no such diagonal
        best = None
        for r in range(len(loop)):
            rot = loop[r:] + loop[:r]
            bad = sum(1 for k in range(2, len(rot) - 1) if same_face(rot[0], rot[k]))
            if best is None or bad < best[0]:
                best = (bad, rot)
            if bad == 0:
                break
        assert best[0] == 0, (c, loop, "no fan without an in-face diagonal")
        loop = best[1]
        for k in range(1, len(loop) - 1):
            tris.append((loop[0], loop[k], loop[k + 1]))
    return tris


def same_face(e1, e2):
    """do local edges e1 and e2 lie on a common cube face?"""
    for f in FACES:
        fe = {EDGE_OF[frozenset((f[k], f[(k + 1) % 4]))] for k in range(4)}
        if e1 in fe and e2 in fe:
            return True
    return False


def tables():
    """-> edge_table (256,) uint16 (bit e: local edge e is cut), tri_table (256, 16) int8 (-1 terminated), n_tri (256,) int8."""
    edge_table = np.zeros(256, dtype=np.uint16)
    tri_table = -np.ones((256, 16), dtype=np.int8)
    n_tri = np.zeros(256, dtype=np.int8)
    for c in range(256):
        t = case_triangles(c)
        assert len(t) <= 5, (c, len(t))
        n_tri[c] = len(t)
        for k, tri in enumerate(t):
            tri_table[c, 3 * k:3 * k + 3] = tri
            for e in tri:
                edge_table[c] |= 1 << e
    return edge_table, tri_table, n_tri


def main():
    et, tt, nt = tables()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "echoscene_b200", "csrc", "mc_tables.inc")
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_mc_tables.py -- marching-cubes case tables derived from the method's definition (see that file).\n")
        f.write("// bit i of the case index: corner i is inside (value < level); triangles as triples of local edge numbers, -1 terminated.\n")
        f.write("#ifndef MC_TABLE_QUAL\n#define MC_TABLE_QUAL static const\n#endif\n")
        f.write("MC_TABLE_QUAL unsigned short MC_EDGE_TABLE[256] = {\n")
        for r in range(0, 256, 16):
            f.write("  " + ", ".join(f"0x{int(v):03x}" for v in et[r:r + 16]) + ",\n")
        f.write("};\nMC_TABLE_QUAL signed char MC_NUM_TRI[256] = {\n")
        for r in range(0, 256, 32):
            f.write("  " + ", ".join(str(int(v)) for v in nt[r:r + 32]) + ",\n")
        f.write("};\nMC_TABLE_QUAL signed char MC_TRI_TABLE[256][16] = {\n")
        for c in range(256):
            f.write("  {" + ", ".join(f"{int(v):2d}" for v in tt[c]) + "},\n")
        f.write("};\n// local edge -> (di, dj, dk of its base grid point, direction)\nMC_TABLE_QUAL signed char MC_EDGE_BASE[12][4] = {\n")
        for b in EDGE_BASE:
            f.write("  {" + ", ".join(str(v) for v in b) + "},\n")
        f.write("};\n")
    print(f"wrote {path}: max triangles per case {int(nt.max())}, cases with surface {int((nt > 0).sum())}")


if __name__ == "__main__":
    main()
