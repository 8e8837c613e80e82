"""GPU: echo_mesh_marching_cubes / mesh.marching_cubes / mesh.sdf_to_mesh against the CPU oracle, EXACTLY: vertex
coordinates bit for bit (one fp32 subtraction, division and addition per vertex on both sides), vertex and triangle order and
indices identical (both are functions of the volume only).  The oracle itself is pinned to the method by tests/test_mesh_host.py."""
import numpy as np
import pytest
import torch

from echoscene_b200 import mesh
from echoscene_b200._lib import EchoError
from oracle import mesh_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def grid(R):
    return np.stack(np.meshgrid(*[np.arange(R, dtype=np.float32)] * 3, indexing="ij"), -1)


def volumes():
    rng = np.random.default_rng(5)
    R = 32
    yield "sphere", (np.linalg.norm(grid(R) - np.array([15.3, 16.1, 14.7], dtype=np.float32), axis=-1) - 9.5).astype(np.float32), 0.0
    yield "white noise (every ambiguous configuration)", rng.standard_normal((24, 24, 24)).astype(np.float32), 0.1
    R = 64   # an SDF like the decoder's: union of spheres and a box, truncated at +-0.2, meshed at the reference's level
    g = grid(R) / R - 0.5
    d = np.minimum(np.linalg.norm(g - np.array([0.1, -0.05, 0.0], dtype=np.float32), axis=-1) - 0.22,
                   np.linalg.norm(g - np.array([-0.15, 0.1, 0.12], dtype=np.float32), axis=-1) - 0.15)
    box = np.abs(g - np.array([0.0, 0.0, -0.25], dtype=np.float32)).max(-1) - 0.12
    yield "truncated SDF 64^3", np.clip(np.minimum(d, box), -0.2, 0.2).astype(np.float32), 0.02
    yield "odd size", rng.standard_normal((7, 7, 7)).astype(np.float32), -0.3


@pytest.mark.parametrize("name,vol,level", list(volumes()), ids=lambda v: v if isinstance(v, str) else "")
def test_marching_cubes_equals_the_oracle_exactly(name, vol, level):
    v, f = mesh.marching_cubes(torch.from_numpy(vol).to(DEV), level)
    wv, wf = mo.marching_cubes(vol, level)
    assert v.dtype == torch.float32 and f.dtype == torch.int64 and v.is_cuda and f.is_cuda
    assert tuple(v.shape) == wv.shape and tuple(f.shape) == wf.shape, name
    assert np.array_equal(v.cpu().numpy(), wv), name
    assert np.array_equal(f.cpu().numpy(), wf), name
    v2, f2 = mesh.marching_cubes(torch.from_numpy(vol).to(DEV), level)          # deterministic
    assert torch.equal(v, v2) and torch.equal(f, f2)


def test_empty_volume_bad_input_and_sdf_to_mesh():
    v, f = mesh.marching_cubes(torch.ones(16, 16, 16, device=DEV), 0.02)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(EchoError):
        mesh.marching_cubes(torch.ones(16, 16, 8, device=DEV), 0.0)
    with pytest.raises(EchoError, match="CUDA"):
        mesh.marching_cubes(torch.ones(8, 8, 8), 0.0)
    R = 16
    vols = np.stack([(np.linalg.norm(grid(R) - 7.5, axis=-1) - (3.0 + 0.2 * i)).astype(np.float32) for i in range(18)])[:, None]
    sdf = torch.from_numpy(vols).to(DEV)
    vs, fs = mesh.sdf_to_mesh(sdf)                                              # level 0.02, the first 16 objects (util_3d.py:204-208)
    assert len(vs) == len(fs) == 16
    wv, wf = mo.sdf_to_mesh(vols[:16], 0.02)
    for a, b, c, d in zip(vs, fs, wv, wf):
        assert np.array_equal(a.cpu().numpy(), c) and np.array_equal(b.cpu().numpy(), d)
    assert len(mesh.sdf_to_mesh(sdf, render_all=True)[0]) == 18
    with pytest.raises(EchoError):
        mesh.sdf_to_mesh(sdf[:, 0])
