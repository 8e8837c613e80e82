"""SDF -> triangle mesh on the GPU: the last stage of the reference's evaluation pipeline after the VQ-VAE decode,
``sdf_to_mesh`` (model/diff_utils/util_3d.py:194-237), which runs PyMCubes per object on the CPU.  ``marching_cubes`` is
``mcubes.marching_cubes(volume, isovalue)`` on a CUDA tensor (``echo_mesh_marching_cubes``, csrc/mesh.cu); ``sdf_to_mesh`` keeps the
reference's per-object loop, its ``level`` default and its ``verts / n_cell - 0.5`` scaling, and returns the vertex / face lists the
reference hands to ``pytorch3d.structures.Meshes`` (pytorch3d is not a dependency of this package).

PyMCubes is not available in this environment: the case tables are derived from the method's definition
(tools/gen_mc_tables.py), the surface is the marching-cubes surface, but vertex / triangle ORDER and the diagonals on ambiguous
configurations are this library's own ("SDF -> mesh").  No CPU fallback."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib
from ._lib import EchoError

_ws = {}


def _workspace(R: int, device) -> torch.Tensor:
    key = (R, str(device))
    if key not in _ws:
        n = int(_lib.lib().echo_mesh_workspace_bytes(R))
        if n <= 0:
            raise EchoError(f"marching_cubes: resolution {R} outside [2, 512]")
        _ws[key] = torch.empty((n + 3) // 4, dtype=torch.int32, device=device)
    return _ws[key]


@torch.no_grad()
def marching_cubes(volume: torch.Tensor, isovalue: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """volume (R, R, R) on a CUDA device -> (vertices (V, 3) float32 in index coordinates, triangles (F, 3) int64), both on the
    device.  Two library calls: count (one 8-byte host read), then emit into exactly sized outputs."""
    _lib.require_cuda(volume)
    if volume.dim() != 3 or volume.shape[0] != volume.shape[1] or volume.shape[1] != volume.shape[2]:
        raise EchoError(f"marching_cubes: volume must be (R, R, R), got {tuple(volume.shape)}")
    vol = volume.float().contiguous()
    R = int(vol.shape[0])
    ws = _workspace(R, vol.device)
    counts = torch.zeros(2, dtype=torch.int32, device=vol.device)
    fn = _lib.lib().echo_mesh_marching_cubes
    _lib.check(fn(_lib.ptr(vol), R, float(isovalue), None, 0, None, 0, _lib.ptr(counts), _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr()))
    nv, nf = (int(c) for c in counts.tolist())
    verts = torch.empty(nv, 3, dtype=torch.float32, device=vol.device)
    faces = torch.empty(nf, 3, dtype=torch.int32, device=vol.device)
    if nv or nf:
        _lib.check(fn(_lib.ptr(vol), R, float(isovalue), _lib.ptr(verts) if nv else None, nv, _lib.ptr(faces) if nf else None, nf,
                      _lib.ptr(counts), _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr()))
    return verts, faces.to(torch.int64)


@torch.no_grad()
def sdf_to_mesh(sdf: torch.Tensor, level: float = 0.02, color=None, render_all: bool = False) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """sdf (B, 1, R, R, R) -> (verts list, faces list): util_3d.sdf_to_mesh without the pytorch3d container -- the first
    ``min(B, 16)`` objects unless ``render_all`` (:204-208), marching cubes at ``level``, ``verts / n_cell - 0.5`` (:218-219),
    faces int64 (:222).  ``color`` only fills per-vertex textures in the reference and is accepted for signature compatibility."""
    if sdf.dim() != 5 or sdf.shape[1] != 1:
        raise EchoError(f"sdf_to_mesh: sdf must be (B, 1, R, R, R), got {tuple(sdf.shape)}")
    n_cell = sdf.shape[-1]
    n = sdf.shape[0] if render_all else min(sdf.shape[0], 16)
    verts, faces = [], []
    for i in range(n):
        v, f = marching_cubes(sdf[i, 0], level)
        verts.append(v / n_cell - 0.5)
        faces.append(f)
    return verts, faces
