"""One profiled denoiser step (for ncu): build the model, warm up outside the profiler range, then run exactly one step
between cudaProfilerStart/Stop.  Usage under gpurun:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --branch shape
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echoscene_b200 import arch, modules, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--branch", default="shape", choices=["shape", "layout"])
ap.add_argument("--precision", default="bf16")
ap.add_argument("--nodes", type=int, default=16)
args = ap.parse_args()
dev = torch.device("cuda:0")
n, t = args.nodes, 4 * args.nodes
g = synth.make_scene_graph(n, t, 2)
tri = g.triples.to(dev)
if args.branch == "shape":
    m, _ = bench.build_model(args.precision, dev)
    uc, x = synth.shape_inputs(n, 2, same_noise=True)
    uc, x = uc.to(dev), x.to(dev)
    for i in range(2):
        x = m.ddim_step(x, uc, tri, 99 - i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    x = m.ddim_step(x, uc, tri, 97)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    sd = arch.make_state_dict(arch.unet1d_specs(synth.layout_cfg()), synth.WEIGHT_SEED_LAYOUT)
    m = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                            channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                            crossattn_dim=1280, enable_t_emb=True, precision=args.precision)
    m.load_state_dict(sd)
    m = m.to(dev)
    obj_embed, x = synth.layout_inputs(n, 2)
    obj_embed, x = obj_embed.to(dev), x.to(dev)
    noise = torch.randn_like(x)
    for i in range(2):
        x = m.ddpm_step(x, obj_embed, tri, 999 - i, noise)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    x = m.ddpm_step(x, obj_embed, tri, 997, noise)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", float(x.abs().mean()))
