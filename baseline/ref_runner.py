"""The UNMODIFIED reference (baseline/_ref, see install_ref.py) driven through its own modules for the benched workload.

MEASUREMENT / TEST INFRASTRUCTURE only: bench.py's `--impl reference` arm, its `gpu_eager_baseline` figure and the
reference-binding GPU test import this; nothing under echoscene_b200/ does.  The model is the reference's own
`UNet3DModel` (model/networks/diffusion_shape/openai_model_3d.py) built from the reference's own YAML
(config/sdfusion-txt2shape_mp.yaml), the DDIM tables come from the reference's own ldm_diffusion_util, and the x_prev update
is the arithmetic of samplers/ddim.py:246-261 (the sampler class itself hard-codes `.cuda()` and a tqdm loop, so it is driven
step by step exactly as oracle/gen_golden.py drives it).  Weights: the seeded synthetic state_dict of the bench
(arch.make_state_dict), loaded with strict=True.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "model"))


def load_reference():
    """-> namespace of reference classes imported from baseline/_ref (stubs for absent optional third-party packages)."""
    if not available():
        raise RuntimeError("baseline/_ref is not installed (python baseline/install_ref.py in the build container)")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["ECHOSCENE_REF"] = REF
    from oracle import ref_import
    ref_import.REF_ROOT = REF
    return ref_import.load()


class ReferenceShapeStepper:
    """One DDIM iteration of the shape branch with the reference's own modules: e_t = UNet3DModel(x, uc, triples, t) and
    the eta = 0 update.  device = "cpu" (the --impl reference arm) or "cuda" (gpu_eager_baseline)."""

    def __init__(self, state_dict, device="cpu", ddim_steps=100):
        ref = load_reference()
        with open(os.path.join(REF, "config/sdfusion-txt2shape_mp.yaml")) as f:
            sdf = yaml.safe_load(f)
        ukw = dict(sdf["unet"]["params"], conditioning_key=sdf["model"]["params"]["conditioning_key"])   # network.py:15-17
        self.model = ref.UNet3DModel(**ukw).eval()
        self.model.load_state_dict(state_dict, strict=True)
        self.model = self.model.to(device)
        self.device = torch.device(device)
        u = ref.ldm_util
        self.ddim_ts = u.make_ddim_timesteps("uniform", ddim_steps, 1000, verbose=False)
        betas = u.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
        ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
        self.sig, self.al, self.alp = u.make_ddim_sampling_parameters(ac.numpy(), self.ddim_ts, 0.0, verbose=False)
        self.s1m = np.sqrt(1.0 - self.al)

    @torch.no_grad()
    def e_t(self, x, uc, triples, index):
        ts = torch.full((x.shape[0],), int(self.ddim_ts[index]), dtype=torch.long, device=self.device)
        return self.model(x, uc, triples, ts, context=uc)

    @torch.no_grad()
    def step(self, x, uc, triples, index):
        e_t = self.e_t(x, uc, triples, index)
        b, dev = x.shape[0], self.device
        a_t = torch.full((b, 1, 1, 1, 1), float(self.al[index]), device=dev)
        a_prev = torch.full((b, 1, 1, 1, 1), float(self.alp[index]), device=dev)
        sigma_t = torch.full((b, 1, 1, 1, 1), float(self.sig[index]), device=dev)
        s1m = torch.full((b, 1, 1, 1, 1), float(self.s1m[index]), device=dev)
        pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        return a_prev.sqrt() * pred_x0 + dir_xt
