"""Shared helpers of the parity tests (TEST INFRASTRUCTURE; the only place besides bench.py/smoke that touches oracle/)."""
import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: "within 1e-3 rel fp32".  rel = max|a-b| / max|b|  and  ||a-b||_2 / ||b||_2, both must hold.
FP32_TOL = 1e-3
# ECHO_PREC_BF16: bf16 operands, fp32 accumulation.  MEASURED (profiles/r2_parity_and_baselines_call1.txt, r2_bf16_drift.json):
# one step deviates at most 1.56e-2 max-rel / 1.16e-2 rel-L2 from the fp32 path over all 100 steps of the benched chain (N = 16),
# 1.40e-2 / 1.16e-2 from the oracle at N = 16 and N = 32; the bound is that measurement + ~25 % margin.
BF16_TOL = 2e-2
# free-running 100-step bf16 DDIM chain against the fp32 chain (x_t, N = 16): measured worst 5.9e-3 max-rel / 3.4e-3 rel-L2
BF16_CHAIN_TOL = 1e-2
# single bf16 operators / the VQ-VAE decoder against fp32 references (not part of the benched step)
BF16_OP_TOL = 3e-2


def rel_err(a: torch.Tensor, b: torch.Tensor):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b)
    return float(d.abs().max() / b.abs().max().clamp_min(1e-30)), float(d.norm() / b.norm().clamp_min(1e-30))


def assert_close(a, b, tol=FP32_TOL, what=""):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    mx, l2 = rel_err(a, b)
    assert mx < tol and l2 < tol, f"{what}: max-rel {mx:.3e}, rel-L2 {l2:.3e} exceed {tol:.1e}"
    return mx, l2


def gold(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu")


def to_cuda(sd):
    return {k: v.cuda() for k, v in sd.items()}
