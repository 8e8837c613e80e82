"""Shared helpers of the parity tests (TEST INFRASTRUCTURE; the only place besides bench.py/smoke that touches oracle/)."""
import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: "within 1e-3 rel fp32".  rel = max|a-b| / max|b|  and  ||a-b||_2 / ||b||_2, both must hold.
FP32_TOL = 1e-3
# ECHO_PREC_BF16: bf16 operands, fp32 accumulation.  SURVEY §7 calibration: one step deviates ~1e-2 rel-L2 from fp32.
BF16_TOL = 3e-2
# free-running 100-step bf16 DDIM chain against the fp32 chain (x_t, N = 16): set from profiles/r2_bf16_drift.json + margin
BF16_CHAIN_TOL = 6e-2


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
