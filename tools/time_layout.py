"""Layout-steps/s (UNet1DModel forward + DDPM update, chained) at N nodes.  ECHO_NO_PDL=1 / ECHO_NO_GRAPH=1 for A/B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
r = bench.layout_rate(torch.device("cuda:0"), bench.peaks(), prec)
print(f"pdl={'off' if os.environ.get('ECHO_NO_PDL') else 'on'} graph={'off' if os.environ.get('ECHO_NO_GRAPH') else 'on'} {prec}: "
      f"{r['ms_per_step']:.3f} ms/step  {r['value']:.1f} steps/s  hbm frac {r['roofline']['frac']:.4f}")
