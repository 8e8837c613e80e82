"""GPU: the persistent executor of the layout step (csrc/layout_mk.cu: one cooperative kernel per DDPM iteration) against the
per-layer kernels it replaces and against the oracle, over the row classes of its program (1 / <= 8 / 16-row tiles), plus the
diagnostics that show WHICH path ran (persistent kernel steps, CUDA-graph replays)."""
import ctypes as C

import pytest
import torch

from echoscene_b200 import _lib, arch, synth
from oracle import cases, echoscene_oracle as orc
from test_model_gpu import layout_model
from util import FP32_TOL, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def layout_sd():
    return arch.make_state_dict(arch.unet1d_specs(cases.layout_cfg()), cases.WEIGHT_SEED_LAYOUT)


def info(m):
    out = (C.c_int64 * 6)()
    _lib.check(_lib.lib().echo_debug_layout_info(m._handle, out))
    d = dict(zip(("mk_steps", "graph_replays", "stages", "ops", "graph_kernels", "ctas"), list(out)))
    assert d["ctas"] >= 0, "the persistent executor's barrier watchdog fired"
    return d


def scene(n, t, seed):
    g = synth.make_scene_graph(n, t, seed) if n > 1 else synth.SceneGraph(1, torch.zeros(0, 3, dtype=torch.int64))
    obj_embed, x = synth.layout_inputs(n, seed + 7)
    noise = torch.randn(n, 8, generator=torch.Generator().manual_seed(seed + 9))
    return g.triples.to(DEV), obj_embed.to(DEV), x.to(DEV), noise.to(DEV)


@pytest.fixture()
def per_layer_mode():
    yield
    _lib.lib().echo_debug_set_layout_mode(0)


@pytest.mark.parametrize("n,t,step", [(16, 64, 999), (8, 32, 500), (32, 128, 3), (1, 0, 0), (5, 4, 17), (24, 100, 731), (2, 1, 250)])
def test_persistent_step_matches_per_layer_kernels(layout_sd, per_layer_mode, n, t, step):
    m = layout_model(layout_sd)
    tri, obj, x, noise = scene(n, t, 40 + n)
    got = m.ddpm_step(x, obj, tri, step, noise)
    i = info(m)
    if i["ctas"] == 0:
        pytest.skip("cooperative launch unavailable on this device")
    assert i["mk_steps"] == 1 and i["stages"] > 50, i
    again = m.ddpm_step(x, obj, tri, step, noise)
    assert torch.equal(got, again), "the persistent step is not run-to-run deterministic"
    _lib.lib().echo_debug_set_layout_mode(1)
    want = m.ddpm_step(x, obj, tri, step, noise)
    assert info(m)["mk_steps"] == 2, "mode 1 must not use the persistent kernel"
    # same fp32 arithmetic, different summation order inside each dot product
    assert_close(got, want, 2e-5, f"persistent vs per-layer layout step N={n} T={t}")


def test_graph_replay_is_active_on_the_default_stream(layout_sd, per_layer_mode):
    """capture used to start on the caller's stream, which fails on the legacy default stream; the replayed graph is now
    captured on a handle-owned stream and its use is observable."""
    m = layout_model(layout_sd)
    tri, obj, x, noise = scene(16, 64, 3)
    _lib.lib().echo_debug_set_layout_mode(1)
    a = m.ddpm_step(x, obj, tri, 10, noise)
    b = m.ddpm_step(x, obj, tri, 10, noise)
    i = info(m)
    assert i["graph_replays"] == 2 and i["graph_kernels"] > 100, i
    assert torch.equal(a, b)


@pytest.mark.parametrize("n,t", [(16, 64), (8, 32)])
def test_persistent_step_vs_oracle(layout_sd, n, t):
    cfg = cases.layout_cfg()
    m = layout_model(layout_sd, time_num=1000)
    g = synth.make_scene_graph(n, t, 60 + n)
    obj_embed, x = synth.layout_inputs(n, 70 + n)
    noise = torch.randn(n, 8, generator=torch.Generator().manual_seed(n))
    step = 421
    got = m.ddpm_step(x.to(DEV), obj_embed.to(DEV), g.triples.to(DEV), step, noise.to(DEV))
    with torch.no_grad():
        eps = orc.unet1d_forward(layout_sd, cfg, x, obj_embed, g.triples, torch.full((n,), step, dtype=torch.int64)).squeeze(-1)
        want = orc.ddpm_update(orc.DDPMSchedule(time_num=1000), x, eps, step, noise)
    assert_close(got, want, FP32_TOL, f"persistent layout step N={n} vs oracle")


def test_program_is_rebuilt_when_the_graph_changes(layout_sd, per_layer_mode):
    m = layout_model(layout_sd)
    outs = {}
    for n, t in [(16, 64), (8, 32), (16, 64), (32, 128), (16, 48)]:
        tri, obj, x, noise = scene(n, t, 5)
        outs.setdefault((n, t), []).append(m.ddpm_step(x, obj, tri, 77, noise))
    assert torch.equal(outs[(16, 64)][0], outs[(16, 64)][1])
    _lib.lib().echo_debug_set_layout_mode(1)
    for (n, t), got in outs.items():
        tri, obj, x, noise = scene(n, t, 5)
        assert_close(got[0], m.ddpm_step(x, obj, tri, 77, noise), 2e-5, f"N={n} T={t} after program rebuilds")


def test_long_chain_stays_in_lockstep(layout_sd, per_layer_mode):
    """200 chained iterations (the epoch counters advance every launch) against the per-layer path on the same noise."""
    m = layout_model(layout_sd)
    tri, obj, x0, _ = scene(16, 64, 11)
    noise = torch.randn(200, 16, 8, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    x = x0.clone()
    for i in range(200):
        x = m.ddpm_step(x, obj, tri, 999 - i, noise[i])
    _lib.lib().echo_debug_set_layout_mode(1)
    y = x0.clone()
    for i in range(200):
        y = m.ddpm_step(y, obj, tri, 999 - i, noise[i])
    assert_close(x, y, 1e-3, "200-step layout chain, persistent vs per-layer")
