"""Golden vectors for the greedy object selection of the training forward: the reference's own
Sg2ScDiffModel.select_sdfs (model/EchoScene.py:246-319, sample_type='greedy'), imported in place from /root/reference (build
container only) and called unbound on a bare namespace -- it reads nothing of the model but `diffusion_bs`.  Its trailing `.cuda()`
calls (:310-312) are made no-ops for the call (CPU container).  Pins scene.Sg2ScDiffModel.select_sdfs; writes
tests/golden/select_sdfs.pt.  Usage: python oracle/gen_golden_select.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

# (name, nodes per scene, diffusion_bs): whole scenes while they fit; the third scene does not; a batch that fits exactly; one scene
CASES = [("three_scenes_bs16", [7, 7, 7], 16), ("ragged_bs16", [5, 9, 4, 6], 16), ("exact_fit_bs12", [6, 6, 3], 12), ("single_scene_bs16", [10], 16),
         ("all_fit_bs64", [8, 12, 16], 64)]


def batch(sizes, seed):
    gen = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    o2s = torch.cat([torch.full((s,), i, dtype=torch.int64) for i, s in enumerate(sizes)])
    objs = torch.randint(0, 30, (n,), generator=gen)
    tri, off = [], 0
    for s in sizes:          # collate-style: triples of a scene index that scene's nodes only
        t = torch.randint(0, s, (3 * s, 3), generator=gen)
        t[:, 1] = torch.randint(0, 16, (3 * s,), generator=gen)
        t[:, 0] += off
        t[:, 2] += off
        tri.append(t)
        off += s
    triples = torch.cat(tri)
    sdfs = torch.randn(n, 1, 4, 4, 4, generator=gen)
    uc, c = torch.randn(n, 1, 8, generator=gen), torch.randn(n, 1, 8, generator=gen)
    return o2s, objs, triples, sdfs, uc, c


def main():
    ref_import.load()
    Model = importlib.import_module("model.EchoScene").Sg2ScDiffModel
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # :310-312 move the selection to the GPU
    out = {}
    try:
        for name, sizes, bs in CASES:
            o2s, objs, triples, sdfs, uc, c = batch(sizes, 100 + len(sizes))
            ns = types.SimpleNamespace(diffusion_bs=bs)
            cats, d = Model.select_sdfs(ns, o2s, objs, objs, triples, sdfs, uc, c, sample_type="greedy")
            out[name] = {"obj_cat_selected": cats, "sdf": d["sdf"], "uc_s": d["uc_s"], "c_s": d["c_s"], "scene_ids": torch.from_numpy(np.asarray(d["scene_ids"])),
                         "triples": d["triples"]}
    finally:
        torch.Tensor.cuda = cuda
    torch.save(out, os.path.join(ROOT, "tests", "golden", "select_sdfs.pt"))
    print("select_sdfs(greedy) goldens:", {k: (len(v["obj_cat_selected"]), len(v["triples"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
