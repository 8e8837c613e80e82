"""The optimizer step of scripts/train_3dfront.py:247-259 at the reference's parameter count (encoders + layout denoiser 164 M + shape
denoiser 430 M, in the tensor shapes of the real state_dicts): the reference's own sequence in torch (clip_grad_norm_ on the shape
denoiser, the per-parameter isnan().any() scrub loop, torch.optim.AdamW.step) against echoscene_b200.train.FusedAdamW (one fused
multi-tensor pass, no host synchronisation).  HBM roofline: 28 bytes per parameter (read p, g, m, v; write p, m, v).
Usage: python tools/time_optimizer.py [--steps 10]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echoscene_b200 import arch, synth, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
specs3 = arch.unet3d_specs(synth.shape_cfg())
specs1 = arch.unet1d_specs(synth.layout_cfg())
shapes3 = [tuple(s.shape) for s in specs3.values() if not s.buffer]
shapes1 = [tuple(s.shape) for s in specs1.values() if not s.buffer]


def make(seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g) * 0.02) for s in shapes1 + shapes3]
    for p in ps:
        p.grad = torch.randn(p.shape, device=dev, generator=g) * 0.01
    return ps, ps[len(shapes1):]


n_params = sum(torch.Size(s).numel() for s in shapes1 + shapes3)
out = {"parameters": n_params, "tensors": len(shapes1) + len(shapes3), "bytes_per_step": 28 * n_params}
pk = bench.peaks()

ref_p, ref_clip = make(1)
opt = torch.optim.AdamW(ref_p, lr=1e-4)


def ref_step():
    torch.nn.utils.clip_grad_norm_(ref_clip, 5.0)
    for group in opt.param_groups:
        for p in group["params"]:
            if p.grad is not None and p.requires_grad and torch.isnan(p.grad).any():
                p.grad[torch.isnan(p.grad)] = 0
    opt.step()


for _ in range(2):
    ref_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    ref_step()
torch.cuda.synchronize()
out["reference_sequence_ms"] = (time.perf_counter() - t0) / a.steps * 1e3
del opt, ref_p, ref_clip
torch.cuda.empty_cache()

our_p, our_clip = make(1)
ours = train.FusedAdamW(our_p, lr=1e-4, clip_params=our_clip, clip_max_norm=5.0)
for _ in range(2):
    ours.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(a.steps):
    ours.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
gbs = 28 * n_params / (ms * 1e-3) / 1e9
out.update({"fused_ms": ms, "fused_gbs": gbs, "hbm_peak_gbs": pk["hbm_gbs"], "hbm_frac": gbs / pk["hbm_gbs"],
            "speedup_vs_reference_sequence": out["reference_sequence_ms"] / ms, "info": ours.info()})
print(json.dumps(out))
