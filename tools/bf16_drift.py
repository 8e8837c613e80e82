"""bf16 (tcgen05) shape path against the fp32 parity path over a whole DDIM chain, at the benched config (N = 16, T = 64,
S = 100) -- VERDICT r1 task 1(b).  For every step i:
  * `step_*`  : e_t of both precisions evaluated on the SAME x_t (the fp32 chain's), i.e. the one-step error;
  * `chain_*` : the free-running bf16 chain's x against the fp32 chain's x after step i, i.e. the accumulated drift.
rel = max|a-b| / max|b| and ||a-b||_2 / ||b||_2 (tests/util.rel_err).  Writes profiles/r2_bf16_drift.json.
Usage: python tools/bf16_drift.py [--nodes 16] [--steps 100] [--ref fp32|x3] [--out profiles/r2_bf16_drift.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from echoscene_b200 import arch, modules, synth  # noqa: E402
from util import rel_err  # noqa: E402


def shape_model(sd, precision, ddim_steps):
    m = modules.UNet3DModel(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                            attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3,
                            use_spatial_transformer=True, transformer_depth=1, context_dim=1280, legacy=False,
                            messsage_passing=True, conditioning_key="crossattn", enable_t_emb=True, precision=precision,
                            ddim_steps=ddim_steps)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def drift(n_nodes=16, n_triples=64, steps=100, ddim_steps=100, ref="fp32", test="bf16", seed=2):
    sd = arch.make_state_dict(arch.unet3d_specs(synth.shape_cfg()), synth.WEIGHT_SEED_SHAPE)
    g = synth.make_scene_graph(n_nodes, n_triples, seed)
    uc, x_T = synth.shape_inputs(n_nodes, seed, same_noise=True)
    uc, tri, x_T = uc.cuda(), g.triples.cuda(), x_T.cuda()
    m_ref, m_b = shape_model(sd, ref, ddim_steps), shape_model(sd, test, ddim_steps)
    _, ts = m_ref.schedule_tables()
    x_ref, x_b = x_T.clone(), x_T.clone()
    rows = []
    for i in range(steps):
        index = ddim_steps - 1 - i
        t = torch.full((n_nodes,), int(ts[index]), dtype=torch.int64, device="cuda")
        e_ref = m_ref(x_ref, uc, tri, t)
        e_b = m_b(x_ref, uc, tri, t)
        smx, sl2 = rel_err(e_b, e_ref)
        x_ref = m_ref.ddim_step(x_ref, uc, tri, index)
        x_b = m_b.ddim_step(x_b, uc, tri, index)
        cmx, cl2 = rel_err(x_b, x_ref)
        rows.append({"step": i, "ddim_index": index, "t": int(ts[index]), "step_max_rel": smx, "step_rel_l2": sl2,
                     "chain_max_rel": cmx, "chain_rel_l2": cl2})
    assert torch.isfinite(x_b).all() and torch.isfinite(x_ref).all()
    return {"config": {"n_nodes": n_nodes, "n_triples": n_triples, "chain_steps": steps, "ddim_steps": ddim_steps,
                       "reference_mode": ref, "tested_mode": test, "weights_seed": synth.WEIGHT_SEED_SHAPE},
            "worst_step": {"max_rel": max(r["step_max_rel"] for r in rows), "rel_l2": max(r["step_rel_l2"] for r in rows)},
            "worst_chain": {"max_rel": max(r["chain_max_rel"] for r in rows), "rel_l2": max(r["chain_rel_l2"] for r in rows)},
            "chain_end": {"max_rel": rows[-1]["chain_max_rel"], "rel_l2": rows[-1]["chain_rel_l2"]},
            "per_step": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=16)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--ref", default="fp32")
    ap.add_argument("--test", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_bf16_drift.json"))
    a = ap.parse_args()
    d = drift(a.nodes, 4 * a.nodes, a.steps, 100, a.ref, a.test)
    with open(a.out, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps({k: d[k] for k in ("config", "worst_step", "worst_chain", "chain_end")}))
