"""ms per shape step (N nodes, bf16), chained, CUDA events.  With ECHO_SKIP=<kernel classes> in the environment the
difference to the unskipped run is the in-situ (warm L2, overlapped launches) cost of those kernels; results are then
garbage by construction.  Usage: ECHO_SKIP=gn_apply python tools/time_step.py [--nodes 16] [--steps 50]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echoscene_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=16)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
m, _ = bench.build_model(a.precision, dev)
g = synth.make_scene_graph(a.nodes, 4 * a.nodes, 2)
tri = g.triples.to(dev)
uc, x = synth.shape_inputs(a.nodes, 2, same_noise=True)
uc, x = uc.to(dev), x.to(dev)
y = torch.empty_like(x)
m._ensure(a.nodes, tri.shape[0])
m.frozen = True
for i in range(5):
    m.ddim_step(x, uc, tri, 99 - i, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for i in range(a.steps):
    m.ddim_step(x, uc, tri, 99 - (i % 100), out=y)
e1.record()
torch.cuda.synchronize()
print(f"skip={os.environ.get('ECHO_SKIP', '-'):40s} nodes={a.nodes} ms/step={e0.elapsed_time(e1) / a.steps:.3f}")
