"""Pin the oracle against the reference and write tests/golden/ fixtures — TEST INFRASTRUCTURE.

Run in the BUILD container only (needs /root/reference):  python oracle/gen_golden.py

For each seeded case in oracle/cases.py it
  1. builds the reference module (GraphTripleConvNet / UNet1DModel / UNet3DModel, imported in place from
     /root/reference), loads ``arch.make_state_dict(specs, seed)`` with strict=True (this also proves that
     echoscene_b200/arch.py reproduces the reference's state_dict keys and shapes),
  2. runs the reference forward on CPU fp32 (the reference's own samplers for the chains),
  3. runs the oracle restatement (oracle/echoscene_oracle.py) on the same inputs and records max-abs / rel-L2
     deviation in tests/golden/PINNING.json,
  4. stores the REFERENCE outputs as .pt fixtures.  Inputs/weights are not stored: they are regenerated from
     the integer seeds by oracle/cases.py.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echoscene_b200 import arch                      # noqa: E402
from oracle import cases, echoscene_oracle as orc    # noqa: E402
from oracle import ref_import                        # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _dev(a: torch.Tensor, b: torch.Tensor):
    d = (a.double() - b.double())
    return {"max_abs": float(d.abs().max()), "rel_l2": float(d.norm() / b.double().norm().clamp_min(1e-30)),
            "ref_abs_max": float(b.abs().max())}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    ref = ref_import.load()
    os.makedirs(GOLD, exist_ok=True)
    pin = {"reference": "ymxlzgy/echoscene @ /root/reference", "torch": torch.__version__, "cases": {}}
    with open(os.path.join(ref_import.REF_ROOT, "config/full_mp.yaml")) as f:
        full = yaml.safe_load(f)
    with open(os.path.join(ref_import.REF_ROOT, "config/sdfusion-txt2shape_mp.yaml")) as f:
        sdf = yaml.safe_load(f)

    # ---- 1. GraphTripleConvNet (layout echo geometry: 768 -> 1280) --------------------------------------------
    lcfg = cases.layout_cfg()
    gcfg = lcfg.gcn()
    net = ref.GraphTripleConvNet(input_dim_obj=gcfg.input_dim_obj, input_dim_pred=gcfg.input_dim_pred,
                                 num_layers=gcfg.num_layers, hidden_dim=gcfg.hidden_dim, residual=True,
                                 pooling="avg", mlp_normalization="batch", output_dim=gcfg.output_dim).eval()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    net.load_state_dict(sd, strict=True)
    g, obj, pred = cases.gcn_inputs(cases.GCN_CASE, gcfg)
    edges, _ = orc.edges_of(g.triples)
    with torch.no_grad():
        r_obj, r_pred = net(obj, pred, edges)
        o_obj, o_pred = orc.graph_triple_conv_net(sd, "", obj, pred, edges)
        l0_obj, l0_pred = net.gconvs[0](obj, pred, edges)
    pin["cases"]["gcn_obj"] = _dev(o_obj, r_obj)
    pin["cases"]["gcn_pred"] = _dev(o_pred, r_pred)
    torch.save({"obj": r_obj, "pred": r_pred, "layer0_obj": l0_obj, "layer0_pred": l0_pred,
                "gather_s": obj[edges[:, 0]], "gather_o": obj[edges[:, 1]]},
               os.path.join(GOLD, "gcn_layout_n8.pt"))

    # ---- 2. UNet1DModel one step + 10-step DDPM chain (BASELINE config 1) -----------------------------------
    kw = dict(full["layout_branch"]["denoiser_kwargs"])
    m1 = ref.UNet1DModel(**kw).eval()
    sd1 = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
    m1.load_state_dict(sd1, strict=True)
    g, obj_embed, x, t = cases.layout_step_inputs(cases.LAYOUT_CASE, lcfg)
    with torch.no_grad():
        r = m1(x, obj_embed, g.triples, t, None)
        o = orc.unet1d_forward(sd1, lcfg, x, obj_embed, g.triples, t)
    pin["cases"]["layout_step"] = _dev(o, r)
    step_out = r

    # chain through the reference's own sampler with injected noise
    dk = dict(full["layout_branch"]["diffusion_kwargs"])
    dk["time_num"] = cases.LAYOUT_CHAIN_STEPS
    g, obj_embed, x_T, noises = cases.layout_chain_inputs(cases.LAYOUT_CASE, lcfg, cases.LAYOUT_CHAIN_STEPS)
    dp = ref.DiffusionPoint(m1, full["layout_branch"], **dk)   # echo2layout.py:25-30
    it = iter([x_T] + list(noises))

    def noise_fn(size=None, dtype=None, device=None):
        return next(it).clone()
    with torch.no_grad():
        r_chain = dp.gen_samples_sg((cases.LAYOUT_CASE.n_nodes, lcfg.in_channels), "cpu", obj_embed,
                                    triples=g.triples, condition=None, noise_fn=noise_fn, clip_denoised=False)
    with torch.no_grad():
        o_chain = orc.layout_chain(sd1, lcfg, obj_embed, g.triples, x_T, noises, cases.LAYOUT_CHAIN_STEPS)
    pin["cases"]["layout_chain10"] = _dev(o_chain, r_chain)
    sch = orc.DDPMSchedule(time_num=1000)
    gd = ref.DiffusionPoint(m1, full["layout_branch"], **full["layout_branch"]["diffusion_kwargs"]).diffusion
    tab_dev = max(float((getattr(gd, n).float() - getattr(sch, n)).abs().max()) for n in
                  ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                   "posterior_mean_coef2", "posterior_log_variance_clipped"))
    pin["cases"]["ddpm_tables_1000"] = {"max_abs": tab_dev}
    torch.save({"step": step_out, "chain": r_chain}, os.path.join(GOLD, "layout_n8.pt"))

    # ---- 3. UNet3DModel one step + 3-step DDIM chain --------------------------------------------------------
    scfg = cases.shape_cfg()
    ukw = dict(sdf["unet"]["params"], conditioning_key=sdf["model"]["params"]["conditioning_key"])  # network.py:15-17
    m3 = ref.UNet3DModel(**ukw).eval()
    sd3 = arch.make_state_dict(arch.unet3d_specs(scfg), cases.WEIGHT_SEED_SHAPE)
    m3.load_state_dict(sd3, strict=True)
    g, uc, x, t = cases.shape_step_inputs(cases.SHAPE_CASE, scfg)
    t0 = time.time()
    with torch.no_grad():
        r = m3(x, uc, g.triples, t, context=uc)
    t_ref = time.time() - t0
    with torch.no_grad():
        o = orc.unet3d_forward(sd3, scfg, x, uc, g.triples, t)
        emb = orc._linear(sd3, "time_embed.2", torch.nn.functional.silu(
            orc._linear(sd3, "time_embed.0", orc.timestep_embedding(t, scfg.model_channels))))
        r_latent = m3.shape_messsage_passing(uc, g.triples, x, emb, enable_t_emb=True)
    pin["cases"]["shape_step"] = {**_dev(o, r), "ref_seconds": t_ref}
    step3 = r

    # DDIM: reference sampler tables + update, driven step by step (the sampler class hard-codes .cuda())
    sch = orc.DDIMSchedule(100)
    ddim_ts = ref.ldm_util.make_ddim_timesteps("uniform", 100, 1000, verbose=False)
    betas = ref.ldm_util.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
    sig, al, alp = ref.ldm_util.make_ddim_sampling_parameters(ac.numpy(), ddim_ts, 0.0, verbose=False)
    pin["cases"]["ddim_tables_100"] = {
        "timesteps_equal": bool((ddim_ts == sch.ddim_timesteps).all()),
        "alphas_max_abs": float(np.abs(al - sch.alphas).max()),
        "alphas_prev_max_abs": float(np.abs(alp - sch.alphas_prev).max()),
        "sigma_max": float(np.abs(sig).max())}

    g, uc, x_T, _ = cases.shape_step_inputs(cases.SHAPE_CHAIN_CASE, scfg, same_noise=True)
    xr = x_T
    with torch.no_grad():
        for i, step in enumerate(np.flip(ddim_ts)[: cases.SHAPE_CHAIN_STEPS]):
            index = len(ddim_ts) - i - 1
            ts = torch.full((xr.shape[0],), int(step), dtype=torch.long)
            e_t = m3(xr, uc, g.triples, ts, context=uc)
            # samplers/ddim.py:246-261
            b = xr.shape[0]
            a_t = torch.full((b, 1, 1, 1, 1), al[index])
            a_prev = torch.full((b, 1, 1, 1, 1), alp[index])
            sigma_t = torch.full((b, 1, 1, 1, 1), sig[index])
            s1m = torch.full((b, 1, 1, 1, 1), np.sqrt(1.0 - al)[index])
            pred_x0 = (xr - s1m * e_t) / a_t.sqrt()
            dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
            xr = a_prev.sqrt() * pred_x0 + dir_xt
        xo = orc.shape_chain(sd3, scfg, uc, g.triples, x_T, 100, cases.SHAPE_CHAIN_STEPS)
    pin["cases"]["shape_chain3"] = _dev(xo, xr)
    torch.save({"step": step3, "latent": r_latent, "chain": xr}, os.path.join(GOLD, "shape.pt"))

    pin["param_counts"] = {"unet1d": arch.count_params(arch.unet1d_specs(lcfg)),
                           "unet3d": arch.count_params(arch.unet3d_specs(scfg))}
    with open(os.path.join(GOLD, "PINNING.json"), "w") as f:
        json.dump(pin, f, indent=1)
    print(json.dumps(pin, indent=1))
    worst = max(v.get("rel_l2", 0.0) for v in pin["cases"].values())
    assert worst < 1e-5, f"oracle deviates from the reference: {worst}"


if __name__ == "__main__":
    main()
