"""GPU: the shape step at the BENCHED sizes against the ORACLE (CPU restatement pinned bit-exactly to the reference), in every
precision mode, plus the 100-step DDIM chain of the benched bf16 path against the fp32 parity path (VERDICT r1 task 1).

  N = 16 / T = 64   BASELINE config 2 (what bench.py times)
  N = 32 / T = 128  BASELINE config 3 size

The oracle forward costs ~6 s (N = 16) / ~12 s (N = 32) of host time; it is computed once per size and shared by the modes.
"""
import os
import sys

import pytest
import torch

from echoscene_b200 import _lib, arch, synth
from oracle import cases, echoscene_oracle as orc
from test_model_gpu import shape_model
from util import BF16_CHAIN_TOL, BF16_TOL, FP32_TOL, assert_close

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def shape_sd():
    return arch.make_state_dict(arch.unet3d_specs(cases.shape_cfg()), cases.WEIGHT_SEED_SHAPE)


_oracle_cache = {}


def oracle_step(sd, n, t):
    """inputs + the oracle's e_t for one N-node scene, per-object timesteps spread over the schedule"""
    if (n, t) not in _oracle_cache:
        g = synth.make_scene_graph(n, t, 100 + n)
        uc, x = synth.shape_inputs(n, 200 + n, same_noise=False)
        ts = torch.tensor([991, 501, 1, 251, 11, 741, 331, 91] * (n // 8), dtype=torch.int64)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            want = orc.unet3d_forward(sd, cases.shape_cfg(), x, uc, g.triples, ts)
        _oracle_cache[(n, t)] = (g, uc, x, ts, want)
    return _oracle_cache[(n, t)]


@pytest.mark.parametrize("n,t", [(16, 64), (32, 128)])
@pytest.mark.parametrize("precision", ["fp32", "x3", "bf16"])
def test_shape_step_full_size_vs_oracle(shape_sd, precision, n, t):
    if precision != "fp32" and not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    g, uc, x, ts, want = oracle_step(shape_sd, n, t)
    m = shape_model(shape_sd, precision=precision)
    got = m(x.to(DEV), uc.to(DEV), g.triples.to(DEV), ts.to(DEV))
    tol = BF16_TOL if precision == "bf16" else FP32_TOL
    mx, l2 = assert_close(got, want, tol, f"shape step N={n} {precision} vs oracle")
    print(f"\n[parity] shape step N={n} T={t} {precision}: max-rel {mx:.3e} rel-L2 {l2:.3e} (tol {tol:.0e})")


def test_bf16_chain_drift_n16():
    """The benched path over the benched chain: 100 DDIM steps at N = 16, bf16 tcgen05 against the fp32 parity path (itself
    pinned to the reference at 1e-3): one-step error on identical x_t at EVERY step, and the free-running chain's drift."""
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    from bf16_drift import drift
    d = drift(16, 64, 100, 100, ref="x3", test="bf16")
    print(f"\n[parity] bf16 chain N=16 S=100: worst step {d['worst_step']}, worst chain {d['worst_chain']}, end {d['chain_end']}")
    assert d["worst_step"]["rel_l2"] < BF16_TOL and d["worst_step"]["max_rel"] < BF16_TOL
    assert d["worst_chain"]["rel_l2"] < BF16_CHAIN_TOL and d["worst_chain"]["max_rel"] < BF16_CHAIN_TOL


def test_x3_chain_matches_fp32_chain_n16():
    """The split-precision tensor-core mode (hi/lo bf16 x 3 MMAs into one fp32 TMEM accumulator) against the fp32 FMA path
    over a 10-step chain at the benched size: both hold north_star's 1e-3."""
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    from bf16_drift import drift
    d = drift(16, 64, 10, 100, ref="fp32", test="x3")
    print(f"\n[parity] x3 chain N=16 10 steps: worst step {d['worst_step']}, worst chain {d['worst_chain']}")
    assert d["worst_step"]["rel_l2"] < FP32_TOL and d["worst_step"]["max_rel"] < FP32_TOL
    assert d["worst_chain"]["rel_l2"] < FP32_TOL and d["worst_chain"]["max_rel"] < FP32_TOL
