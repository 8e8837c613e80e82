"""VQ-VAE decode (VQVAE.decode_no_quant) objects/s at N objects.  Usage: python tools/time_vqvae.py [bf16|fp32] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import arch, modules  # noqa: E402
from oracle import cases  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
cfg = cases.vqvae_cfg()
dd = dict(double_z=False, z_channels=3, resolution=64, in_channels=1, out_ch=1, ch=64, ch_mult=[1, 2, 4], num_res_blocks=1,
          attn_resolutions=[], dropout=0.0)
m = modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim, precision=prec)
m.load_state_dict(arch.make_state_dict(arch.vqvae_decode_specs(cfg), cases.WEIGHT_SEED_VQVAE))
m = m.to(dev)
z = cases.vqvae_inputs(n, seed=3).to(dev)
m.frozen = False
for _ in range(2):
    out = m.decode_no_quant(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
reps = 5
e0.record()
for _ in range(reps):
    out = m.decode_no_quant(z)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"vqvae decode {prec} N={n}: {ms:.2f} ms  ({ms / n:.2f} ms/object, {723e9 * n / ms / 1e9:.0f} TFLOP/s at 723 GFLOP/object (SURVEY 8f))  "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB torch-side")
