"""The reference's algorithm in eager PyTorch ON THE GPU (cuDNN / cuBLAS), next to this repo's kernels: the oracle
restatement (pinned bit-exactly to the reference modules, oracle/echoscene_oracle.py) run with CUDA tensors.
SURVEY 8(d): "the reference on the same B200 in eager PyTorch ... is the real bar to beat".  Measurement tool only --
nothing in the product path imports this.  Usage: python tools/eager_gpu_baseline.py [--nodes 16] [--steps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import arch, synth  # noqa: E402
from oracle import cases, echoscene_oracle as orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = cases.shape_cfg()
sd = {k: v.to(dev) for k, v in arch.make_state_dict(arch.unet3d_specs(cfg), cases.WEIGHT_SEED_SHAPE).items()}
g = synth.make_scene_graph(a.nodes, 4 * a.nodes, 2)
tri = g.triples.to(dev)
uc, x0 = synth.shape_inputs(a.nodes, 2, same_noise=True)
uc, x0 = uc.to(dev), x0.to(dev)
sch = orc.DDIMSchedule(100)
ts = torch.full((a.nodes,), int(sch.ddim_timesteps[-1]), dtype=torch.int64, device=dev)
a_t, a_prev, _, s1m = [float(v) for v in sch.coeffs(99)]
out = {}
for name, tf32, autocast in [("fp32 strict", False, False), ("fp32 + TF32 (torch default for convs)", True, False),
                             ("bf16 autocast", True, True)]:
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    x = x0.clone()
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for i in range(2 + a.steps):
                if i == 2:
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record()
                e = orc.unet3d_forward(sd, cfg, x, uc, tri, ts).float()
                x = a_prev ** 0.5 * ((x - s1m * e) / a_t ** 0.5) + (1.0 - a_prev) ** 0.5 * e   # samplers/ddim.py:252-261, sigma = 0
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[name] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms}
    except Exception as ex:  # noqa: BLE001
        out[name] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
print(json.dumps({"nodes": a.nodes, "eager_pytorch_on_gpu": out, "torch": torch.__version__}))
