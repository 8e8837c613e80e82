#!/usr/bin/env python
"""bench.py — denoiser-steps/sec of the EchoScene shape-branch hot path (BASELINE.json metric).

A "step" = one DDIM iteration of the shape branch for one scene of N = 16 nodes: UNet3DModel forward (echo message
passing + 3-D UNet over 3x16^3 latents) + the x_prev update, chained (step i feeds step i+1).  Workload = BASELINE.json
configs[1] ("echoscene N=16 nodes, 64^3 SDF latent, 100-step DDIM, bf16, 1xB200").

  python bench.py [--gpus N --steps K --warmup W]           this repo's CUDA path (C ABI of libechoscene_b200.so)
  python bench.py --impl reference ...                      the reference's algorithm on the host CPU cores

With N > 1 GPUs (torchrun, one rank per GPU): the batch is N scenes of 16 nodes (collate-style disjoint union); objects
are sharded 16 per rank; every step each rank embeds its objects, ONE NCCL all-gather exchanges the (16, 64) fp32 shape
codes (the echo exchange), every rank runs the GCN on the whole graph and the UNet trunk on its own objects.  value =
scene-steps/s over the whole job (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_OBJECT_STEP = 557.8e9      # algorithmic FLOPs per object per shape step (2*MAC, reference convention)
N_NODES, N_TRIPLES, DDIM_STEPS = 16, 64, 100


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_model(precision: str, dev):
    from echoscene_b200 import arch, modules, synth
    sd = arch.make_state_dict(arch.unet3d_specs(synth.shape_cfg()), synth.WEIGHT_SEED_SHAPE)
    m = modules.UNet3DModel(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                            attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3,
                            use_spatial_transformer=True, transformer_depth=1, context_dim=1280, legacy=False,
                            messsage_passing=True, conditioning_key="crossattn", enable_t_emb=True, precision=precision,
                            ddim_steps=DDIM_STEPS)
    m.load_state_dict(sd, strict=True)
    return m.to(dev), sd


def layout_rate(dev, pk, precision, steps=200):
    """Secondary figure (SURVEY §8d asks for it next to the headline): layout-steps/s at N = 16 — UNet1DModel forward +
    DDPM update per step, chained; HBM roofline = live weight bytes once per step."""
    from echoscene_b200 import arch, modules, synth
    sd = arch.make_state_dict(arch.unet1d_specs(synth.layout_cfg()), synth.WEIGHT_SEED_LAYOUT)
    m = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                            channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                            crossattn_dim=1280, enable_t_emb=True, precision=precision, time_num=1000)
    m.load_state_dict(sd)
    m = m.to(dev)
    g = synth.make_scene_graph(N_NODES, N_TRIPLES, 2)
    tri = g.triples.to(dev)
    obj_embed, x = synth.layout_inputs(N_NODES, 2)
    obj_embed, x = obj_embed.to(dev), x.to(dev)
    noise = torch.randn(steps, N_NODES, 8, device=dev)
    m._ensure(N_NODES, N_TRIPLES)
    m.frozen = True
    for i in range(5):
        x = m.ddpm_step(x, obj_embed, tri, 999 - i, noise[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        x = m.ddpm_step(x, obj_embed, tri, 994 - i, noise[i])
    e1.record()
    torch.cuda.synchronize()
    m.frozen = False
    ms = e0.elapsed_time(e1) / steps
    live_bytes = 115.30e6 * 4    # live parameters read once per step; the layout branch streams fp32 weights in both modes
    ach = live_bytes / (ms * 1e-3) / 1e9
    import ctypes
    info = (ctypes.c_int64 * 6)()
    _lib_ = __import__("echoscene_b200._lib", fromlist=["_lib"])
    _lib_.check(_lib_.lib().echo_debug_layout_info(m._handle, info))
    return {"value": 1e3 / ms, "unit": "layout-steps/s", "ms_per_step": ms, "n_nodes": N_NODES, "dtype": "f32",
            "executor": {"persistent_kernel_steps": int(info[0]), "graph_replays": int(info[1]), "program_stages": int(info[2]),
                         "program_ops": int(info[3]), "ctas": int(info[5])},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "algorithmic_bytes": live_bytes,
                         "traffic": 330.3e6,
                         "traffic_source": "static: dram__bytes_read + write of one persistent-kernel launch, profiles/r2_ncu_layout_mk_final.txt "
                                           "(below the algorithmic bytes: the time path's weights are tabulated per handle, DESIGN section 4)"}}


def vqvae_decode_rate(dev, precision):
    """Secondary figure (the step right after the shape chain): VQVAE.decode_no_quant over the scene's 16 latents."""
    from echoscene_b200 import arch, modules, synth
    cfg = synth.vqvae_cfg()
    dd = dict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=1, out_ch=cfg.out_ch, ch=cfg.ch,
              ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    m = modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim, precision=precision)
    m.load_state_dict(arch.make_state_dict(arch.vqvae_decode_specs(cfg), synth.WEIGHT_SEED_VQVAE))
    m = m.to(dev)
    z = synth.vqvae_inputs(N_NODES, seed=3).to(dev)
    for _ in range(2):
        m.decode_no_quant(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        out = m.decode_no_quant(z)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    ms = e0.elapsed_time(e1) / 3
    return {"ms_per_scene": ms, "objects": N_NODES, "dtype": precision, "unit": "ms per 16-object decode",
            "tflops_algorithmic": 723e9 * N_NODES / (ms * 1e-3) / 1e12}


def optional_figure(fn, *a):
    """A secondary figure never costs the headline line: any failure becomes {"error": ...}."""
    try:
        return fn(*a)
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:200]}


def layout_batched_rate(dev, scenes=64, steps=20):
    """Secondary figure (BASELINE config 4's shape on the layout branch): one DDPM step over a collated batch of `scenes` scenes of
    16 nodes / 64 triples -- more than 64 rows, so every Linear runs as the fp32-grade 3 x TF32 tensor-core GEMM (csrc/sgemm_x3.cu)
    behind its materialised prologue instead of the weight-streaming few-row kernel; the step is a replayed CUDA graph."""
    from echoscene_b200 import arch, modules, synth
    sd = arch.make_state_dict(arch.unet1d_specs(synth.layout_cfg()), synth.WEIGHT_SEED_LAYOUT)
    m = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                            channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                            crossattn_dim=1280, enable_t_emb=True, precision="fp32", time_num=1000)
    m.load_state_dict(sd)
    m = m.to(dev)
    g = synth.batch_scene_graphs([synth.make_scene_graph(N_NODES, N_TRIPLES, 2 + i) for i in range(scenes)])
    n, tri = g.n_nodes, g.triples.to(dev)
    gen = torch.Generator().manual_seed(5)
    obj_embed, x = torch.randn(n, 640, generator=gen).to(dev), torch.randn(n, 8, generator=gen).to(dev)
    noise = torch.randn(steps + 5, n, 8, device=dev)
    m._ensure(n, tri.shape[0])
    m.frozen = True
    for i in range(5):
        x = m.ddpm_step(x, obj_embed, tri, 999 - i, noise[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        x = m.ddpm_step(x, obj_embed, tri, 994 - i, noise[5 + i])
    e1.record()
    torch.cuda.synchronize()
    m.frozen = False
    ms = e0.elapsed_time(e1) / steps
    flops = 2.0 * 115.30e6 * n          # every live parameter is one multiply-add per row
    return {"scenes": scenes, "nodes": n, "triples": int(tri.shape[0]), "ms_per_batched_step": ms, "value": scenes * 1e3 / ms,
            "unit": "scene-steps/s", "dtype": "f32 (3 x TF32 split on the tensor cores)", "tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
            "seconds_per_scene_for_1000_steps": ms / scenes}


def sdf_to_mesh_time(dev, objects=16, R=64):
    """Secondary figure (the stage after the decode): mesh.sdf_to_mesh over the scene's 16 decoded SDFs -- marching cubes
    at level 0.02 on the GPU (the reference runs PyMCubes per object on the CPU; it is not available here, so there is no reference
    time next to this one).  Synthetic SDFs: truncated spheres of growing radius."""
    from echoscene_b200 import mesh
    g = torch.stack(torch.meshgrid(*[torch.arange(R, dtype=torch.float32)] * 3, indexing="ij"), -1) / R - 0.5
    sdf = torch.stack([(g - torch.tensor([0.01 * i, 0.0, 0.0])).norm(dim=-1) - 0.2 - 0.005 * i for i in range(objects)])
    sdf = sdf.clamp(-0.2, 0.2)[:, None].to(dev)
    for _ in range(2):
        out = mesh.sdf_to_mesh(sdf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        out = mesh.sdf_to_mesh(sdf)
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_scene": e0.elapsed_time(e1) / 5, "objects": objects, "resolution": R, "vertices": int(sum(len(v) for v in out[0])),
            "triangles": int(sum(len(f) for f in out[1])), "unit": "ms per 16-object scene (two library calls and one 8-byte host read per object)"}


def scene_encode_time(dev):
    """Secondary figure (the stage right before the two chains): Sg2ScDiffModel.sample's encoders (init_encoder ->
    manipulate -> rel_s_mlp x2) for the 16-node / 64-triple scene as ONE echo_scene_encode call, fp32.  HBM-bound weight
    streaming like the layout step: 10 GraphTripleConv layers + rel_s_mlp = 28.9 M parameters read once per call."""
    from echoscene_b200 import arch, modules, synth
    cfg = synth.scene_cfg()
    m = modules.SceneEncoder()
    m.load_state_dict(arch.make_state_dict(arch.scene_encoder_specs(cfg), synth.WEIGHT_SEED_SCENE))
    m = m.to(dev)
    g, objs, text, rel = synth.scene_inputs(N_NODES, N_TRIPLES, 2)
    a = [t.to(dev) for t in (objs, g.triples, text, rel)]
    for _ in range(3):
        out = m.encode(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 20
    e0.record()
    for _ in range(reps):
        out = m.encode(*a)
    e1.record()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v).all() for v in out.values())
    ms = e0.elapsed_time(e1) / reps
    params = sum(v.numel() for k, v in m.state_dict().items() if v.dtype == torch.float32)
    return {"ms_per_scene": ms, "n_nodes": N_NODES, "n_triples": N_TRIPLES, "dtype": "f32", "live_parameters": params,
            "unit": "ms per scene encode (one C call; includes the object-id range check's host read)",
            "hbm_gbs_algorithmic": params * 4 / (ms * 1e-3) / 1e9}


def conv_kernel_roofline(dev, pk):
    """The dominant kernel timed alone: tcgen05 implicit-GEMM conv 224@16^3 -> 224, N=16 objects (7 of these per step,
    SURVEY Appendix E).  CUDA events on the launching stream; L2 flushed between launches."""
    from echoscene_b200 import _lib
    if not _lib.lib().echo_has_tcgen05():
        return None
    n, c = N_NODES, 224
    x = torch.randn(n, 16, 16, 16, c, device=dev)
    w = torch.randn(c, c, 3, 3, 3, device=dev) * 0.01
    b = torch.zeros(c, device=dev)
    out = torch.empty(n, 16, 16, 16, c, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _lib.lib()

    def call():
        _lib.check(L.echo_op_conv3d(x.data_ptr(), n, 16, 16, 16, c, w.data_ptr(), b.data_ptr(), c, 3, 1, 1, out.data_ptr(),
                                    _lib.PREC_BF16, _lib.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    # the op entry point converts fp32<->bf16 around the kernel: the library's event probe brackets the kernel launch alone
    import ctypes
    L.echo_debug_probe_begin(n * 4096, c, c, 3)
    for _ in range(10):
        flush.zero_()
        call()
    torch.cuda.synchronize()
    avg = ctypes.c_double(0.0)
    cnt = int(L.echo_debug_probe_end(ctypes.byref(avg)))
    if not cnt:
        return None
    ms = float(avg.value)
    flops = 2.0 * n * 4096 * 27 * c * c
    ach = flops / (ms * 1e-3) / 1e12
    return {"kernel": "same kernel alone, L2 flushed between launches", "ms": ms, "achieved": ach, "peak": pk["bf16_tflops"],
            "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"}


def _workload(n_nodes=N_NODES, n_triples=N_TRIPLES, seed=2):
    from echoscene_b200 import arch, synth
    sd = arch.make_state_dict(arch.unet3d_specs(synth.shape_cfg()), synth.WEIGHT_SEED_SHAPE)
    g = synth.make_scene_graph(n_nodes, n_triples, seed)
    uc, x = synth.shape_inputs(n_nodes, seed, same_noise=True)
    return sd, g, uc, x


def cpu_reference_rate(steps: int, warmup: int, budget_s: float = None):
    """The reference on the host cores at the FULL benched config (one scene of N = 16 nodes, T = 64 triples): its own
    UNet3DModel + the DDIM update of samplers/ddim.py, chained.  baseline/_ref holds the unmodified reference (kind
    "reference"); when it was never installed the oracle port (pinned bit-exactly to it) is timed instead (kind "port").
    `budget_s`: stop timing early once the run has used that many seconds (at least one step is always timed).
    -> (steps/s, seconds per step, kind, steps timed)"""
    t_begin = time.perf_counter()
    torch.set_num_threads(os.cpu_count() or 1)
    sd, g, uc, x = _workload()
    from baseline import ref_runner
    if ref_runner.available():
        stepper = ref_runner.ReferenceShapeStepper(sd, "cpu", DDIM_STEPS)
        step_fn, kind = (lambda xx, i: stepper.step(xx, uc, g.triples, i)), "reference"
    else:
        from echoscene_b200 import synth
        from oracle import echoscene_oracle as orc
        cfg, sch = synth.shape_cfg(), orc.DDIMSchedule(DDIM_STEPS)

        def step_fn(xx, i):
            ts = torch.full((N_NODES,), int(sch.ddim_timesteps[i]), dtype=torch.int64)
            return orc.ddim_update(sch, xx, orc.unet3d_forward(sd, cfg, xx, uc, g.triples, ts), i)[0]
        kind = "port"
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x = step_fn(x, DDIM_STEPS - 1 - (i % DDIM_STEPS))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                if budget_s is not None and time.perf_counter() - t_begin + dt > budget_s:
                    break
    assert torch.isfinite(x).all()
    t = sum(times) / len(times)
    return 1.0 / t, t, kind, len(times)


def gpu_eager_baseline(dev, steps=3):
    """Secondary figure, the number to beat: the UNMODIFIED reference modules (baseline/_ref) in eager PyTorch on
    this same GPU, same workload, chained DDIM steps -- with torch's defaults (TF32 convs via cuDNN, fp32 matmuls) and with
    TF32 off (the precision the 1e-3 parity contract is stated in).  The reference's isnan host syncs stay in."""
    from baseline import ref_runner
    if not ref_runner.available():
        return {"unavailable": "baseline/_ref not installed"}
    sd, g, uc, x0 = _workload()
    stepper = ref_runner.ReferenceShapeStepper(sd, dev, DDIM_STEPS)
    uc, tri, x0 = uc.to(dev), g.triples.to(dev), x0.to(dev)
    out = {"impl": "reference modules (baseline/_ref), eager PyTorch, same GPU", "n_nodes": N_NODES, "timed_steps": steps}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, conv_tf32 in (("torch_default_tf32_convs", True), ("strict_fp32", False)):
            torch.backends.cudnn.allow_tf32 = conv_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            x = x0.clone()
            for i in range(2):
                x = stepper.step(x, uc, tri, DDIM_STEPS - 1 - i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for i in range(steps):
                x = stepper.step(x, uc, tri, DDIM_STEPS - 3 - i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def x3_parity_mode_rate(dev, steps=10):
    """Secondary figure: the same chained shape step in ECHO_PREC_X3 -- fp32 activations, every contraction on tcgen05 with hi/lo
    bf16 operands and three MMAs per k-step into one fp32 TMEM accumulator: north_star's 1e-3 parity contract ON the tensor
    cores (measured 8e-5 against the oracle at N = 16 / 32, tests/test_parity_full_gpu.py).  Like-for-like with the
    reference's own GPU default (TF32 convolutions)."""
    from echoscene_b200 import synth
    m, _ = build_model("x3", dev)
    g = synth.make_scene_graph(N_NODES, N_TRIPLES, 2)
    tri = g.triples.to(dev)
    uc, x = synth.shape_inputs(N_NODES, 2, same_noise=True)
    uc, x = uc.to(dev), x.to(dev)
    y = torch.empty_like(x)
    m._ensure(N_NODES, N_TRIPLES)
    m.frozen = True
    for i in range(3):
        m.ddim_step(x, uc, tri, DDIM_STEPS - 1 - i, out=y)
        x, y = y, x
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        m.ddim_step(x, uc, tri, DDIM_STEPS - 4 - i, out=y)
        x, y = y, x
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    ms = e0.elapsed_time(e1) / steps
    ach = FLOP_PER_OBJECT_STEP * N_NODES / (ms * 1e-3) / 1e12
    return {"precision": "x3 (split bf16 hi/lo, 3 tcgen05 MMAs per k-step, fp32 activations)", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "timed_steps": steps, "tflops_algorithmic": ach, "parity_vs_oracle": "<= 1e-3 (measured 8e-5)"}


def config3_rate(dev, steps=10):
    """Secondary figure, BASELINE config 3: "echoscene N=32 nodes, 64^3 SDF latent, 250-step ... bf16, 1xB200" -- one 32-node scene
    (T = 128 triples) on the 250-step schedule (timesteps range(0, 1000, 4) + 1), chained steps.  Parity of this size against the
    oracle: tests/test_parity_full_gpu.py (N = 32 / T = 128)."""
    from echoscene_b200 import arch, modules, synth
    n, t, S = 32, 128, 250
    sd = arch.make_state_dict(arch.unet3d_specs(synth.shape_cfg()), synth.WEIGHT_SEED_SHAPE)
    m = modules.UNet3DModel(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                            attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3, use_spatial_transformer=True,
                            transformer_depth=1, context_dim=1280, legacy=False, messsage_passing=True, conditioning_key="crossattn",
                            enable_t_emb=True, precision="bf16", ddim_steps=S)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    g = synth.make_scene_graph(n, t, 5)
    tri = g.triples.to(dev)
    uc, x = synth.shape_inputs(n, 5, same_noise=True)
    uc, x = uc.to(dev), x.to(dev)
    y = torch.empty_like(x)
    m._ensure(n, t)
    m.frozen = True
    for i in range(3):
        m.ddim_step(x, uc, tri, S - 1 - i, out=y)
        x, y = y, x
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        m.ddim_step(x, uc, tri, S - 4 - i, out=y)
        x, y = y, x
    e1.record()
    torch.cuda.synchronize()
    m.frozen = False
    assert torch.isfinite(x).all()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "echoscene N=32 nodes, T=128 triples, 250-step schedule, bf16", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "timed_steps": steps, "seconds_per_250_step_chain": 250 * ms * 1e-3,
            "tflops_algorithmic": FLOP_PER_OBJECT_STEP * n / (ms * 1e-3) / 1e12}


def strong_scaling_line(m, dev, world, rank, steps):
    """north_star's literal partition: ONE N = 16 scene over the G GPUs, 16 / G objects per rank, NCCL all-gather of the
    (16 / G, 64) codes every step (the echo exchange).  value = steps/s of that one scene (strong scaling)."""
    import torch.distributed as dist
    from echoscene_b200 import synth
    if N_NODES % world:
        return {"skipped": f"{N_NODES} objects do not split over {world} ranks"}
    k = N_NODES // world
    g = synth.make_scene_graph(N_NODES, N_TRIPLES, 2)
    tri = g.triples.to(dev)
    uc, x_all = synth.shape_inputs(N_NODES, 2, same_noise=True)
    uc = uc.to(dev)
    x = x_all[rank * k:(rank + 1) * k].contiguous().to(dev)
    y = torch.empty_like(x)
    codes_all = torch.empty(N_NODES, 64, device=dev)
    xs = torch.cuda.Stream(device=dev)

    def step(xin, xout, i):
        index = DDIM_STEPS - 1 - (i % DDIM_STEPS)
        xs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(xs):
            codes = m.embed_local(xin, N_NODES, tri.shape[0])
            dist.all_gather_into_tensor(codes_all, codes)
        m.trunk_local(xin, rank * k, codes_all, uc, tri, index=index, out=xout, codes_stream=xs)

    m.frozen = False
    m._ensure(N_NODES, tri.shape[0], k)
    m.frozen = True
    for i in range(3):
        step(x, y, i)
        x, y = y, x
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        step(x, y, i)
        x, y = y, x
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    m.frozen = False
    ms = float(t.item()) / steps
    return {"scaling": "strong", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "scenes": 1, "objects_per_rank": k, "steps": steps,
            "exchange": f"NCCL all-gather of ({k},64) fp32 codes per step"}


def scene_sharded_line(m, dev, world, rank, steps, scenes_per_rank=8):
    """BASELINE config 4 shape: a batch of 8 x G scenes of N = 16, whole scenes per rank (block-diagonal graph: the echo never
    crosses scenes), NO data-path collective; every rank steps its 8 scenes (128 objects) as one batched call.  value =
    scene-steps/s over the job (weak scaling; 64 scenes at G = 8)."""
    import torch.distributed as dist
    from echoscene_b200 import synth
    graphs = [synth.make_scene_graph(N_NODES, N_TRIPLES, 100 + rank * scenes_per_rank + i) for i in range(scenes_per_rank)]
    batch = synth.batch_scene_graphs(graphs)
    ucs, xs_ = zip(*[synth.shape_inputs(N_NODES, 100 + rank * scenes_per_rank + i, same_noise=True) for i in range(scenes_per_rank)])
    uc, x, tri = torch.cat(ucs).to(dev), torch.cat(xs_).to(dev), batch.triples.to(dev)
    y = torch.empty_like(x)
    m.frozen = False
    m._ensure(batch.n_nodes, tri.shape[0])
    m.frozen = True
    for i in range(2):
        m.ddim_step(x, uc, tri, DDIM_STEPS - 1 - i, out=y)
        x, y = y, x
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        m.ddim_step(x, uc, tri, DDIM_STEPS - 3 - i, out=y)
        x, y = y, x
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    m.frozen = False
    assert torch.isfinite(x).all()
    ms = float(t.item()) / steps
    return {"scaling": "weak", "value": scenes_per_rank * world * 1e3 / ms, "unit": "scene-steps/s", "ms_per_batched_step": ms,
            "scenes": scenes_per_rank * world, "scenes_per_rank": scenes_per_rank, "objects_per_rank": scenes_per_rank * N_NODES,
            "steps": steps, "exchange": "none (scene-sharded: the collated graph is block-diagonal per scene)",
            "tflops_algorithmic_per_gpu": FLOP_PER_OBJECT_STEP * N_NODES * scenes_per_rank / (ms * 1e-3) / 1e12}


REFERENCE_BUDGET_S = 200.0   # the whole --impl reference run must end "within a few minutes"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every timed step is a FULL N = 16 step (~6 s on 16 host threads); when K + W of them would not fit the budget the
    # run is shortened and the line reports the step count it actually timed
    warm = 1
    rate, t, kind, steps = cpu_reference_rate(args.steps, warm, REFERENCE_BUDGET_S)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "denoiser-steps/sec", "value": rate, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "echoscene N=16 nodes, 64^3 SDF (3x16^3 latent), 100-step DDIM: shape denoiser step "
                                   "(UNet3DModel forward incl. echo message passing + DDIM update)",
                       "n_nodes": N_NODES, "n_triples": N_TRIPLES, "ddim_steps": DDIM_STEPS, "scenes": 1},
            "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": cores, "kind": kind,
                             "sample": f"{steps} timed full steps (N = 16 objects, {t:.2f} s each, torch fp32, {cores} threads) "
                                       + ("of the unmodified reference modules in baseline/_ref" if kind == "reference"
                                          else "of the oracle port (baseline/_ref absent)")},
            "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-graph-gcn", action="store_true",
                    help="sharded step: run the echo GCN on the whole batched graph on every rank (round-1 behaviour) instead of on the "
                         "connected components of the rank's own objects")
    ap.add_argument("--no-graph", action="store_true", help="launch every step's kernels directly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the denoiser hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from echoscene_b200 import _lib, synth
    pk = peaks()
    precision = args.precision
    if precision == "bf16" and not _lib.lib().echo_has_tcgen05():
        raise SystemExit("bf16 precision needs the sm_100a tcgen05 kernels (B200)")
    m, sd = build_model(precision, dev)

    # ---- synthetic workload: `world` scenes of 16 nodes, batched graph, 16 objects per rank ----
    graphs = [synth.make_scene_graph(N_NODES, N_TRIPLES, 2 + r) for r in range(world)]
    batch = synth.batch_scene_graphs(graphs)
    ucs, xs = zip(*[synth.shape_inputs(N_NODES, 2 + r, same_noise=True) for r in range(world)])
    uc_all = torch.cat(ucs).to(dev)
    tri = batch.triples.to(dev)
    n_total = batch.n_nodes
    obj_begin = rank * N_NODES
    x_host = xs[rank].contiguous().pin_memory()
    x = x_host.to(dev, non_blocking=True)
    x_next = torch.empty_like(x)
    codes_all = torch.empty(n_total, 64, device=dev)
    L = _lib.lib()

    xstream = torch.cuda.Stream(device=dev) if world > 1 else None

    def step(xin, xout, i):
        index = DDIM_STEPS - 1 - (i % DDIM_STEPS)
        if world == 1:
            m.ddim_step(xin, uc_all, tri, index, out=xout)
        else:
            # embed + the echo exchange ((16,64) fp32 per rank, NCCL over NVLink) on their own stream; the trunk starts at
            # once and only its echo chain (GCN -> cross-attention vectors, needed at input block 4) waits for the codes
            xstream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(xstream):
                codes = m.embed_local(xin, n_total, tri.shape[0])
                dist.all_gather_into_tensor(codes_all, codes)
            m.trunk_local(xin, obj_begin, codes_all, uc_all, tri, index=index, out=xout, codes_stream=xstream,
                          restrict_to_components=not args.full_graph_gcn)

    m._ensure(n_total, tri.shape[0], N_NODES if world > 1 else None)
    m.frozen = True
    for i in range(max(args.warmup, 3)):
        step(x, x_next, i)
        x, x_next = x_next, x
    torch.cuda.synchronize()

    # ---- the step as a replayed CUDA graph: its ~300 launches (and, sharded, the NCCL all-gather) are captured ONCE with the DDIM
    #      index read from the device (ECHO_INDEX_FROM_DEVICE), two graphs for the x <-> x_next ping-pong; a chain is then
    #      "set the index, replay", which takes the host out of the loop (rank skew at 8 GPUs) ----
    graphs, launches_per_step, graph_note = None, None, "disabled (--no-graph)"
    if not args.no_graph:
        from echoscene_b200._lib import INDEX_FROM_DEVICE

        def step_dev(xin, xout):
            if world == 1:
                m.ddim_step(xin, uc_all, tri, INDEX_FROM_DEVICE, out=xout)
            else:
                xstream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(xstream):
                    codes = m.embed_local(xin, n_total, tri.shape[0])
                    dist.all_gather_into_tensor(codes_all, codes)
                m.trunk_local(xin, obj_begin, codes_all, uc_all, tri, index=INDEX_FROM_DEVICE, out=xout, codes_stream=xstream,
                              restrict_to_components=not args.full_graph_gcn)
        try:
            m.set_step_index(DDIM_STEPS - 1)
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                for a_, b_ in ((x, x_next), (x_next, x)):     # warm the device-index path on the capture stream
                    step_dev(a_, b_)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
            graphs = []
            for a_, b_ in ((x, x_next), (x_next, x)):
                g_ = torch.cuda.CUDAGraph()
                L.echo_launch_count_reset()
                with torch.cuda.graph(g_, stream=cap):
                    step_dev(a_, b_)
                launches_per_step = int(L.echo_launch_count())
                graphs.append(g_)
            torch.cuda.synchronize()
            graph_note = "replayed CUDA graph per step (index on the device)"
        except Exception as e:   # noqa: BLE001  -- capture unsupported here: time the direct launches and say so
            graphs, graph_note = None, f"capture failed, direct launches: {e!r}"[:200]
            torch.cuda.synchronize()
    ok = torch.tensor([1 if graphs else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # every rank replays, or none does (the all-gather is collective)
    if not int(ok.item()):
        graphs = None

    if graphs:
        direct_step = step

        def step(xin, xout, i):   # noqa: F811  (xin / xout alternate exactly as captured)
            m.set_step_index(DDIM_STEPS - 1 - (i % DDIM_STEPS))
            graphs[0 if xin.data_ptr() == x_ptr0 else 1].replay()
        x_ptr0 = x.data_ptr()
        for i in range(2):
            step(x, x_next, i)
            x, x_next = x_next, x
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    L.echo_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(x, x_next, i)
        x, x_next = x_next, x
    e1.record()
    barrier()
    launches = int(L.echo_launch_count())
    if graphs:   # the library's launchers ran at capture time: kernels inside one captured step x steps, + the index writes
        launches = args.steps * (launches_per_step + 1)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps * world / (ms * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of x_t, D2H of x_prev every step) ----
    out_host = torch.empty_like(x_host)
    x_host.copy_(x.cpu())
    barrier()
    e0.record()
    for i in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        step(x, x_next, i)
        out_host.copy_(x_next, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        x_host, out_host = out_host, x_host
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.steps * world / (float(t.item()) * 1e-3)
    m.frozen = False

    # ---- dominant kernel timed IN SITU: CUDA events (recorded by the library on its launching stream) around every launch
    #      of the 3x3x3 conv 224@16^3 -> 224 while a few more real steps run: warm, between its actual neighbours ----
    probe = None
    if precision == "bf16":
        m.frozen = True
        rows_local = N_NODES * 16 * 16 * 16
        L.echo_debug_probe_begin(rows_local, 224, 224, 3)
        probe_step = direct_step if graphs else step   # the event probe brackets launches the library makes itself
        for i in range(10):
            probe_step(x, x_next, i)
            x, x_next = x_next, x
        torch.cuda.synchronize()
        import ctypes
        avg = ctypes.c_double(0.0)
        n_probe = int(L.echo_debug_probe_end(ctypes.byref(avg)))
        m.frozen = False
        if n_probe:
            probe = (n_probe, float(avg.value))

    if rank == 0:
        assert torch.isfinite(x).all(), "non-finite latent after the timed chain"
        ach = FLOP_PER_OBJECT_STEP * N_NODES * value / 1e12       # whole job
        per_gpu = ach / world
        peak = pk["bf16_tflops_sustained"]
        step_roof = {"achieved": per_gpu, "peak": peak, "unit": "TFLOP/s", "frac": per_gpu / peak,
                     "definition": "557.8 GFLOP/object/step (SURVEY 8d, reference FLOP count) x 16 objects x steps/s per GPU"}
        flop_launch = 2.0 * N_NODES * 4096 * 27 * 224 * 224
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel: Conv3d 3x3x3 224->224 @16^3 x 16 objects (7 launches per step, SURVEY App. E)",
                "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r1_ncu_conv224.txt (ncu --set full):
                # 32.18 MB + 0.16 MB -- the bf16 input and the weights once; the output (29 MB) stays in the 126 MB L2
                "traffic": 32.34e6,
                "traffic_source": "static: one ncu --set full capture of this launch, profiles/r1_ncu_conv224.txt (not re-measured in this run)",
                "algorithmic_bytes": 2.0 * (2 * N_NODES * 4096 * 224) + 2.0 * 27 * 224 * 224,
                "flop_per_launch": flop_launch,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']}): kernel timed inside a long step",
                "step": step_roof}
        if probe:
            roof["launches_timed"] = probe[0]
            roof["ms_per_launch"] = probe[1]
            roof["achieved"] = flop_launch / (probe[1] * 1e-3) / 1e12
            roof["frac"] = roof["achieved"] / peak
        if precision != "bf16":
            roof["note"] = "fp32 FMA parity mode: not a tensor-core run"
            roof["achieved"], roof["frac"] = step_roof["achieved"], step_roof["frac"]
        kr = conv_kernel_roofline(dev, pk) if world == 1 else None
        if kr:
            roof["alone"] = kr
        line = {"metric": "denoiser-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "echoscene N=16 nodes, 64^3 SDF (3x16^3 latent), 100-step DDIM: shape denoiser step "
                                       "(UNet3DModel forward incl. echo message passing + DDIM update)",
                           "n_nodes": N_NODES, "n_triples": N_TRIPLES, "ddim_steps": DDIM_STEPS, "scenes": world,
                           "sharding": "1 scene" if world == 1 else f"{world} scenes batched, 16 objects per rank, NCCL all-gather of (16,64) fp32 codes per step; echo GCN on "
                                       + ("the whole batched graph" if args.full_graph_gcn else "the connected components of the rank's objects"),
                           "l2": "not flushed: one step streams 0.84 GB of bf16 weights + >2 GB of activations (>> 126 MB L2)",
                           "weights": "random init (reference initialisers, zero-init tensors re-drawn), seed 12"},
                "clocks": clocks, "gpu_launches": launches, "launch_mode": graph_note,
                "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": x.numel() * 4, "d2h_bytes_per_step": x.numel() * 4},
                "roofline": roof}
        if world == 1:
            line["layout_branch"] = layout_rate(dev, pk, "bf16" if precision == "bf16" else "fp32")
            line["vqvae_decode"] = vqvae_decode_rate(dev, "bf16" if precision == "bf16" else "fp32")
            # the reference runs the chains back to back (EchoScene.py:403-418): 1000 DDPM layout steps, 100 DDIM shape
            # steps, one decode
            line["full_chain_seconds_per_scene"] = {
                "layout_1000_ddpm_steps": 1000.0 / line["layout_branch"]["value"], "shape_100_ddim_steps": DDIM_STEPS / value,
                "vqvae_decode": line["vqvae_decode"]["ms_per_scene"] * 1e-3,
                "total": 1000.0 / line["layout_branch"]["value"] + DDIM_STEPS / value + line["vqvae_decode"]["ms_per_scene"] * 1e-3}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, tcpu, kind, _ = cpu_reference_rate(2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": cores, "kind": kind,
                                    "sample": f"2 timed full steps (N = 16 objects, {tcpu:.2f} s each, torch fp32, {cores} threads) of "
                                              + ("the unmodified reference modules in baseline/_ref" if kind == "reference"
                                                 else "the oracle port (baseline/_ref absent)")}
            line["gpu_eager_baseline"] = optional_figure(gpu_eager_baseline, dev)
        if world == 1:
            # last, and never fatal: a secondary figure must not cost the headline line
            line["scene_encode"] = optional_figure(scene_encode_time, dev)
            line["layout_branch_batched_64_scenes"] = optional_figure(layout_batched_rate, dev)
            line["sdf_to_mesh"] = optional_figure(sdf_to_mesh_time, dev)
            if precision == "bf16":
                line["parity_mode_x3"] = optional_figure(x3_parity_mode_rate, dev)
                line["config3_n32_s250"] = optional_figure(config3_rate, dev)
    # ---- the other partitions of SURVEY 8(e), measured in the same run (every rank takes part; rank 0 reports) ----
    extra = {}
    if precision == "bf16":
        try:
            extra["config4_scene_sharded"] = scene_sharded_line(m, dev, world, rank, 10)
        except Exception as e:   # noqa: BLE001
            extra["config4_scene_sharded"] = {"error": repr(e)[:200]}
        if world > 1:
            try:
                extra["strong_scaling_one_scene"] = strong_scaling_line(m, dev, world, rank, 30)
            except Exception as e:   # noqa: BLE001
                extra["strong_scaling_one_scene"] = {"error": repr(e)[:200]}
    if rank == 0:
        line.update(extra)
        try:   # derived: the three chains of a scene when scenes are sampled as collated batches (sample_scenes), per GPU
            lb, c4, dec = line["layout_branch_batched_64_scenes"], extra["config4_scene_sharded"], line["vqvae_decode"]
            per_rank = c4["scenes_per_rank"]
            line["full_chain_seconds_per_scene_batched"] = {
                "layout_1000_ddpm_steps_64_scene_batches": lb["seconds_per_scene_for_1000_steps"],
                "shape_100_ddim_steps_8_scene_batches": c4["ms_per_batched_step"] * 1e-3 * DDIM_STEPS / per_rank,
                "vqvae_decode": dec["ms_per_scene"] * 1e-3,
                "total": lb["seconds_per_scene_for_1000_steps"] + c4["ms_per_batched_step"] * 1e-3 * DDIM_STEPS / per_rank
                         + dec["ms_per_scene"] * 1e-3,
                "note": "derived from layout_branch_batched_64_scenes, config4_scene_sharded and vqvae_decode of this run"}
        except (KeyError, TypeError):
            pass
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        if graphs:
            # NCCL kernels captured in CUDA graphs keep the communicator busy at teardown (destroy_process_group never returned in
            # a 2-GPU run): drop the graphs, drain the device, and leave without the collective shutdown
            del graphs
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
