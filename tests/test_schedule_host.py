"""CPU: the sampler tables the library builds at handle creation (host code of layout.cu / shape.cu, exported as
echo_debug_ddpm_tables / echo_debug_ddim_schedule) against the oracle's schedules, which restate how the reference builds
its buffers (GaussianDiffusion.__init__ diffusion_ddpm.py:133-162; make_beta_schedule / make_ddim_timesteps /
make_ddim_sampling_parameters ldm_diffusion_util.py:43-96).  Same tolerance as the GPU tests of the same tables (1e-6 relative;
the DDIM timesteps are integers: exact)."""
import ctypes as C

import numpy as np
import pytest
import torch

from echoscene_b200 import _lib
from oracle import echoscene_oracle as orc


@pytest.mark.parametrize("time_num,b0,b1", [(1000, 1e-4, 0.02), (10, 1e-4, 0.02), (250, 1e-4, 0.02), (1, 1e-4, 0.02), (100, 5e-4, 0.05)])
def test_ddpm_tables(time_num, b0, b1):
    out = np.zeros((5, time_num), dtype=np.float32)
    assert _lib.lib().echo_debug_ddpm_tables(time_num, b0, b1, out.ctypes.data_as(C.c_void_p)) == 0
    want = orc.DDPMSchedule(b0, b1, time_num).tables().numpy()
    assert out.shape == want.shape and np.isfinite(out[:, 1:]).all()
    # t = 0: posterior variance is 0 -> log clipped at 1e-20 on both sides
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("S,T", [(100, 1000), (250, 1000), (10, 1000), (2, 1000), (1000, 1000), (50, 200)])
def test_ddim_schedule(S, T):
    cap = T
    coef = np.zeros((cap, 4), dtype=np.float32)
    ts = np.zeros(cap, dtype=np.int32)
    n = C.c_int32(0)
    rc = _lib.lib().echo_debug_ddim_schedule(T, S, 0.00085, 0.012, cap, coef.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p),
                                             C.byref(n))
    if S == T:   # c = 1: the last timestep would be T (out of range); the reference indexes out of bounds there too
        assert rc == -1 and b"out of range" in _lib.lib().echo_last_error()
        return
    assert rc == 0
    sch = orc.DDIMSchedule(S, T)
    assert n.value == len(sch.ddim_timesteps) == S
    assert ts[:n.value].tolist() == sch.ddim_timesteps.tolist()                   # S = 100: [1, 11, ..., 991]
    np.testing.assert_allclose(coef[:n.value], sch.table().numpy(), rtol=2e-6, atol=0)


def test_schedule_entry_points_reject_bad_arguments():
    L = _lib.lib()
    assert L.echo_debug_ddpm_tables(10, 1e-4, 0.02, None) == -1
    buf = np.zeros(8, dtype=np.float32)
    assert L.echo_debug_ddpm_tables(0, 1e-4, 0.02, buf.ctypes.data_as(C.c_void_p)) == -1
    n = C.c_int32(0)
    ts = np.zeros(2, dtype=np.int32)
    assert L.echo_debug_ddim_schedule(1000, 100, 0.00085, 0.012, 2, buf.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p),
                                      C.byref(n)) == -1 and b"capacity" in L.echo_last_error()
    assert L.echo_debug_ddim_schedule(10, 100, 0.00085, 0.012, 2, buf.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p),
                                      C.byref(n)) == -1


def test_training_tables_are_bit_equal_to_the_reference_constructors():
    """train.layout_train_tables / shape_train_tables (the q_sample factors of both branches and the shape branch's lvlb_weights)
    against the tensors the reference's own constructors build -- GaussianDiffusion.__init__ and EchoToShape.register_schedule, run by
    oracle/gen_golden_train.py in the build container (tests/golden/train_tables.pt).  Host arithmetic (float64 rounded to fp32 where
    the reference rounds): bit for bit."""
    import os
    from echoscene_b200 import train
    G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_tables.pt"), map_location="cpu")
    a, b = train.layout_train_tables(1000, 1e-4, 0.02)
    assert torch.equal(a, G["layout"]["sqrt_alphas_cumprod"]) and torch.equal(b, G["layout"]["sqrt_one_minus_alphas_cumprod"])
    t = train.shape_train_tables(1000, 0.00085, 0.012)
    for k, v in G["shape"].items():
        assert torch.equal(t[k], v), k
    assert torch.equal(t["logvar"], torch.zeros(1000)) and float(t["lvlb_weights"][0]) == float(t["lvlb_weights"][1])


def test_learning_rate_schedule_is_the_reference_lambda():
    """train.lr_lambda / learning_rate against torch's LambdaLR driven by the reference's own Sg2ScDiffModel.lr_lambda
    (model/EchoScene.py:115-141), restated here from its four branches: the rate in param_groups after `counter` scheduler steps."""
    from echoscene_b200 import train
    lr_init, lr_step, lr_evo = 1e-4, [35000, 70000, 140000], [5e-5, 1e-5, 5e-6]

    def reference_lambda(counter):                       # EchoScene.py:117-128
        if counter < lr_step[0]:
            return 1.0
        elif counter < lr_step[1]:
            return lr_evo[0] / lr_init
        elif counter < lr_step[2]:
            return lr_evo[1] / lr_init
        else:
            return lr_evo[2] / lr_init

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-4)
    probes = {0, 1, 34999, 35000, 35001, 69999, 70000, 139999, 140000, 200000}
    for counter in sorted(probes):
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=reference_lambda, last_epoch=-1)
        for g in opt.param_groups:
            g["lr"] = 1e-4
        sched.last_epoch = counter - 1
        opt.step()
        sched.step()                                     # update_learning_rate, EchoScene.py:138-141
        assert opt.param_groups[0]["lr"] == pytest.approx(train.learning_rate(counter), rel=1e-12), counter
        assert train.lr_lambda(counter, lr_init, lr_step, lr_evo) == reference_lambda(counter)
