"""Whole-scene sampling through the public facade, timed end to end: SGDiff.sample_box_and_shape on a synthetic scene
(scene encoders -> 1000-step DDPM layout chain -> 100-step DDIM shape chain -> VQ-VAE decode), as scripts/eval_3dfront.py:271
calls it.  Random-init weights of the reference architecture (no checkpoints in the image) unless --exp/--epoch are given.

Usage: python tools/sample_scene.py [--config DIR] [--type echoscene|echolayout] [--precision bf16|fp32] [--nodes 16]
                                    [--triples 64] [--scenes 3] [--exp EXP --epoch N] [--out scene.npz]
--config: a directory holding full_mp.yaml, sdfusion-txt2shape_mp.yaml and vqvae_snet.yaml with the reference's key structure
(default: YAML written from the constants of tests/test_sgdiff_host.py at the reference's full sizes)."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from echoscene_b200 import sgdiff  # noqa: E402
from oracle import cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default=None)
ap.add_argument("--type", default="echoscene")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--nodes", type=int, default=16)
ap.add_argument("--triples", type=int, default=64)
ap.add_argument("--scenes", type=int, default=3)
ap.add_argument("--exp", default=None)
ap.add_argument("--epoch", type=int, default=0)
ap.add_argument("--out", default=None)
args = ap.parse_args()

if args.config is None:
    from test_sgdiff_host import MAIN, VQ  # noqa: E402
    from test_zzzz_sgdiff_gpu import DF_FULL  # noqa: E402
    tmp = tempfile.mkdtemp()
    main = dict(MAIN)
    main["shape_branch"] = dict(MAIN["shape_branch"], df_cfg="df.yaml", vq_cfg="vq.yaml")
    for name, body in (("full_mp.yaml", main), ("df.yaml", DF_FULL), ("vq.yaml", VQ)):
        with open(os.path.join(tmp, name), "w") as f:
            yaml.safe_dump(body, f)
    args.config = tmp
vocab = {"object_idx_to_name": ["_scene_"] + [f"c{i}" for i in range(35)], "pred_idx_to_name": ["in"] + [f"p{i}" for i in range(15)]}
t0 = time.time()
model = sgdiff.SGDiff(args.type, os.path.join(args.config, "full_mp.yaml"), vocab, residual=True, precision=args.precision)
if args.exp:
    print("checkpoint:", model.load_networks(args.exp, args.epoch))
model = model.cuda().eval()
print(f"built + moved in {time.time() - t0:.1f} s")
dev = torch.device("cuda")
for i in range(args.scenes):
    g, objs, text, rel = cases.scene_inputs(cases.GraphCase(f"scene{i}", args.nodes, args.triples, 100 + i))
    a = [x.to(dev) for x in (objs, g.triples, text, rel)]
    torch.cuda.synchronize()
    t = time.time()
    out = model.sample_box_and_shape(*a, gen_shape=args.type == "echoscene")
    torch.cuda.synchronize()
    dt = time.time() - t
    shapes = out.get("shapes")
    print(f"scene {i}: {dt:.3f} s  boxes {tuple(out['sizes'].shape)} finite {bool(torch.isfinite(out['sizes']).all())}"
          + (f"  sdf {tuple(shapes.shape)} finite {bool(torch.isfinite(shapes).all())}" if shapes is not None else "")
          + ("  (first scene includes handle creation / CUDA-graph capture)" if i == 0 else ""))
if args.out:
    np.savez(args.out, **{k: v.detach().cpu().numpy() for k, v in out.items() if v is not None})
    print("wrote", args.out)
