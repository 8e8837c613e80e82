"""CPU: the marching-cubes oracle (oracle/mesh_oracle.py) and the derived case tables (tools/gen_mc_tables.py) against the METHOD --
PyMCubes, which the reference calls (model/diff_utils/util_3d.py:217), is not available here, so the oracle is held to properties
any correct marching cubes has: a closed, consistently oriented 2-manifold for a closed level set (also on white noise, where every
ambiguous configuration occurs), boundary edges only on the volume's boundary, vertices on the level set of the edge-interpolated
field, area / volume of an analytic sphere, and the classic edge table entry for entry."""
import numpy as np

from oracle import mesh_oracle as mo


def grid(R):
    return np.stack(np.meshgrid(*[np.arange(R, dtype=np.float32)] * 3, indexing="ij"), -1)


def test_derived_edge_table_is_the_classic_one():
    # the first two rows and a few landmarks of the published 256-entry edge table (Lorensen-Cline numbering, Bourke's listing)
    classic = [0x0, 0x109, 0x203, 0x30a, 0x406, 0x50f, 0x605, 0x70c, 0x80c, 0x905, 0xa0f, 0xb06, 0xc0a, 0xd03, 0xe09, 0xf00,
               0x190, 0x99, 0x393, 0x29a, 0x596, 0x49f, 0x795, 0x69c, 0x99c, 0x895, 0xb9f, 0xa96, 0xd9a, 0xc93, 0xf99, 0xe90]
    assert [int(v) for v in mo.EDGE_TABLE[:32]] == classic
    assert int(mo.EDGE_TABLE[255]) == 0 and int(mo.EDGE_TABLE[0x55]) == 0xff and int(mo.EDGE_TABLE[0xaa]) == 0xff   # bottom ring / top ring
    assert all(int(mo.EDGE_TABLE[c]) == int(mo.EDGE_TABLE[255 - c]) for c in range(256))                             # complement symmetry
    assert int(mo.NUM_TRI.max()) == 5 and int((mo.NUM_TRI > 0).sum()) == 254


def test_sphere_is_a_closed_oriented_manifold_with_the_right_area_and_volume():
    R, c, r = 32, np.array([15.3, 16.1, 14.7], dtype=np.float32), 9.5
    vol = (np.linalg.norm(grid(R) - c, axis=-1) - r).astype(np.float32)
    v, f = mo.marching_cubes(vol, 0.0)
    counts, oriented = mo.edge_use_counts(f)
    assert (counts == 2).all() and oriented
    assert len(v) - len(counts) + len(f) == 2                                  # Euler characteristic of a sphere
    area, volume = mo.area_and_volume(v, f)
    assert abs(area / (4 * np.pi * r * r) - 1) < 0.01 and abs(volume / (4 / 3 * np.pi * r ** 3) - 1) < 0.01 and volume > 0   # outward normals
    assert float(np.abs(np.linalg.norm(v - c, axis=1) - r).max()) < 0.03       # vertices on the level set (linear interpolation error)
    assert f.dtype == np.int64 and v.dtype == np.float32 and int(f.max()) == len(v) - 1 and len(np.unique(f)) == len(v)


def test_white_noise_gives_a_closed_manifold():
    R = 20
    rng = np.random.default_rng(3)
    vol = rng.standard_normal((R, R, R)).astype(np.float32)
    vol[0] = vol[-1] = vol[:, 0] = vol[:, -1] = vol[:, :, 0] = vol[:, :, -1] = 5.0          # outside shell: the level set is closed
    v, f = mo.marching_cubes(vol, 0.0)
    counts, oriented = mo.edge_use_counts(f)
    assert len(f) > 10000 and (counts == 2).all() and oriented
    _, volume = mo.area_and_volume(v, f)
    assert volume > 0


def test_open_surface_has_boundary_edges_only_on_the_volume_boundary():
    R = 24
    g = grid(R)
    vol = (g[..., 0] * 0.3 + g[..., 1] * 0.5 + g[..., 2] - 14.2).astype(np.float32)
    v, f = mo.marching_cubes(vol, 0.0)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und, counts = np.unique(np.sort(e, axis=1), axis=0, return_counts=True)
    assert set(counts.tolist()) == {1, 2}
    b = v[und[counts == 1]]                                                    # (n, 2, 3) end points of boundary edges
    on_wall = ((b == 0) | (b == R - 1)).any(axis=2).all(axis=1)
    assert on_wall.all()
    assert float(np.abs(v[:, 0] * 0.3 + v[:, 1] * 0.5 + v[:, 2] - 14.2).max()) < 1e-4      # a linear field is met exactly


def test_empty_and_full_volumes_and_the_scaling_of_sdf_to_mesh():
    R = 8
    for fill in (1.0, -1.0):
        v, f = mo.marching_cubes(np.full((R, R, R), fill, dtype=np.float32), 0.02)
        assert v.shape == (0, 3) and f.shape == (0, 3)
    vol = (np.linalg.norm(grid(R) - 3.5, axis=-1) - 2.0).astype(np.float32)
    vs, fs = mo.sdf_to_mesh(vol[None, None], level=0.02)
    v, f = mo.marching_cubes(vol, 0.02)
    assert np.array_equal(vs[0], v / np.float32(R) - np.float32(0.5)) and np.array_equal(fs[0], f)          # util_3d.py:218-219


def test_case_tables_are_invariant_under_the_rotations_of_the_cube_and_under_complement():
    """The derived tables treat the 24 rotations of the cube alike: rotating a corner configuration permutes its cut edges and keeps its
    triangle count and its loop structure (sorted loop lengths); the complementary configuration has the same cut edges, the same
    triangles up to orientation (reversed: the normal follows inside -> outside) and fan apex."""
    import itertools
    C = mo.CORNERS.astype(float) - 0.5
    rots = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            M = np.zeros((3, 3))
            for r, (p, s) in enumerate(zip(perm, signs)):
                M[r, p] = s
            if round(np.linalg.det(M)) == 1:
                rots.append(M)
    assert len(rots) == 24
    edges = mo.gen.EDGES
    edge_of = {frozenset(e): i for i, e in enumerate(edges)}

    def loops_of(case):
        tris = [tuple(int(v) for v in mo.TRI_TABLE[case, 3 * t:3 * t + 3]) for t in range(mo.NUM_TRI[case])]
        # boundary edges of the triangle set = the loops' segments; group triangles into fans by shared apex chains (connected components)
        comp = list(range(len(tris)))
        for a in range(len(tris)):
            for b in range(a + 1, len(tris)):
                if len(set(tris[a]) & set(tris[b])) >= 2:
                    ra, rb = comp[a], comp[b]
                    comp = [ra if c == rb else c for c in comp]
        sizes = {}
        for c, t in zip(comp, tris):
            sizes.setdefault(c, set()).update(t)
        return sorted(len(v) for v in sizes.values())

    for M in rots:
        corner_map = [int(np.argmin(np.abs(C - (M @ C[i])).sum(1))) for i in range(8)]            # corner i moves to corner_map[i]
        edge_map = [edge_of[frozenset((corner_map[a], corner_map[b]))] for a, b in edges]
        for case in range(256):
            rc = sum(((case >> i) & 1) << corner_map[i] for i in range(8))
            assert mo.NUM_TRI[rc] == mo.NUM_TRI[case]
            cut = {e for e in range(12) if (int(mo.EDGE_TABLE[case]) >> e) & 1}
            assert {edge_map[e] for e in cut} == {e for e in range(12) if (int(mo.EDGE_TABLE[rc]) >> e) & 1}
            assert loops_of(rc) == loops_of(case), (case, rc)
    for case in range(256):
        comp = 255 - case
        assert mo.EDGE_TABLE[case] == mo.EDGE_TABLE[comp]
        # same loops unless the configuration has an ambiguous face (there the inside-corner rule joins the other pair of edges)
        amb = any(((case >> f[0]) & 1) == ((case >> f[2]) & 1) != ((case >> f[1]) & 1) == ((case >> f[3]) & 1) for f in mo.gen.FACES)
        if not amb:
            assert mo.NUM_TRI[case] == mo.NUM_TRI[comp] and loops_of(case) == loops_of(comp), case


def test_random_volumes_always_give_closed_oriented_manifolds():
    """Property test (hypothesis): for any volume whose outer shell is outside, the mesh is a closed, consistently oriented manifold
    whose enclosed volume is positive -- including volumes full of ambiguous faces, isolated voxels and values equal to the level."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra.numpy import arrays

    @settings(max_examples=40, deadline=None)
    @given(arrays(np.float32, (7, 7, 7), elements=st.sampled_from([-1.0, -0.25, 0.0, 0.25, 1.0]).map(np.float32)),
           st.sampled_from([-0.1, 0.0, 0.1]))
    def check(vol, level):
        vol = vol.copy()
        vol[0] = vol[-1] = vol[:, 0] = vol[:, -1] = vol[:, :, 0] = vol[:, :, -1] = 2.0
        v, f = mo.marching_cubes(vol, level)
        if len(f) == 0:
            assert not (vol < np.float32(level)).any()
            return
        counts, oriented = mo.edge_use_counts(f)
        assert (counts % 2 == 0).all() and oriented           # closed; an edge shared by two touching sheets counts 4
        assert np.isfinite(v).all() and v.min() >= 0 and v.max() <= 6
        _, volume = mo.area_and_volume(v, f)
        assert volume > 0
    check()
