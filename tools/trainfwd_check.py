"""The training forward, value against value (VERDICT r1 task 7, "loss parity"): the UNMODIFIED reference (baseline/_ref) builds
its own SGDiff('echoscene', config/full_mp.yaml), goes to model.train() and runs one `forward_mani` on a collated batch of scenes
(scripts/train_3dfront.py:237-241) -- eager PyTorch on this GPU, TF32 off; then the B200 components (echoscene_b200.sgdiff.SGDiff
built from the same YAML files, loaded with the same weights) run `train().forward_mani` on the same batch under the same seeds.
Every entry of the loss dictionary (shape loss_simple / loss_vlb / loss_total, layout loss.bbox / .trans / .size / .angle) and both
totals must agree within north_star's 1e-3.  One process: nothing is patched, the two models only share the inputs and the seeds.

  python tools/trainfwd_check.py [--scenes 3 --nodes 7] [--out gpurun_out/trainfwd.json]

TEST / MEASUREMENT INFRASTRUCTURE: imports baseline/_ref, never imported by the package.
Three forward_mani calls are compared: a manipulation batch (replace_latent), an addition (one node missing from the encoder-side
scene, replace_latent = False: zero row inserted, touched rows only) and the layout-only model SGDiff('echolayout').  With
random-initialised weights the losses are dominated by the noise target, so the denoiser OUTPUTS (and every other stage that
carries a BatchNorm1d) are compared on identical inputs first (`stages`).
What this pins: BatchNorm1d on batch statistics in all four GCNs and rel_s_mlp, the greedy object selection, the VQ-VAE encode,
q_sample of both branches (one timestep per object / per scene), both denoisers, the loss terms, and the order of the random
draws.  What it does not: gradients (no backward pass on the B200 side).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refbind_check as rb   # noqa: E402  (reference construction helpers)


def batch(n_scenes, nodes, dev):
    """A collated batch as threedfront_dataset.collate_fn builds it: node / triple tensors concatenated, triple indices offset."""
    from echoscene_b200 import synth
    objs, triples, text, rel, o2s = [], [], [], [], []
    off = 0
    for s in range(n_scenes):
        g, o, t, r = synth.scene_inputs(nodes, 3 * nodes, 40 + s)
        tr = g.triples.clone()
        tr[:, 0] += off
        tr[:, 2] += off
        objs.append(o); triples.append(tr); text.append(t); rel.append(r); o2s += [s] * nodes
        off += nodes
    gen = torch.Generator().manual_seed(77)
    n = off
    boxes = torch.randn(n, 6, generator=gen)
    angles = (torch.rand(n, generator=gen) - 0.5) * 6.0
    sdfs = (torch.rand(n, 1, 64, 64, 64, generator=gen) - 0.5) * 0.4              # truncated SDF range (dataset.trunc_thres 0.2)
    return dict(objs=torch.cat(objs).to(dev), triples=torch.cat(triples).to(dev), text=torch.cat(text).to(dev), rel=torch.cat(rel).to(dev),
                o2s=torch.tensor(o2s, dtype=torch.int64), boxes=boxes.to(dev), angles=angles.to(dev), sdfs=sdfs.to(dev))


def without_node(b, k):
    """The encoder-side scene of an addition: node k and its triples removed, later node indices shifted down (what the dataset
    hands over as enc_* next to missing_nodes = [k], threedfront_dataset.py)."""
    keep_n = torch.ones(len(b["objs"]), dtype=torch.bool, device=b["objs"].device)
    keep_n[k] = False
    tr = b["triples"]
    keep_t = (tr[:, 0] != k) & (tr[:, 2] != k)
    tr = tr[keep_t].clone()
    tr[:, 0] -= (tr[:, 0] > k).long()
    tr[:, 2] -= (tr[:, 2] > k).long()
    return dict(objs=b["objs"][keep_n], triples=tr, text=b["text"][keep_n], rel=b["rel"][keep_t])


def call(model, b, manipulated, missing=(), layout_only=False):
    np.random.seed(5)
    torch.manual_seed(4321)
    e = b
    for k in sorted(missing, reverse=True):
        e = without_node(e, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if layout_only:   # SGDiff.forward_mani routes the same argument list to Sg2BoxDiffModel.forward (SGDiff.py:41-43)
        out = model.forward_mani(e["objs"], e["triples"], e["text"], e["rel"], b["objs"], b["objs"], b["triples"], b["boxes"], b["angles"],
                                 None, b["text"], b["rel"], b["o2s"], list(missing), list(manipulated))
    else:
        out = model.forward_mani(e["objs"], e["triples"], e["text"], e["rel"], b["objs"], b["objs"], b["triples"], b["boxes"], b["angles"],
                                 b["sdfs"], b["text"], b["rel"], b["o2s"], list(missing), list(manipulated))
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def loss_table(tag, r, m):
    """-> (rows, worst relative deviation) over the loss dictionary and the two totals of one forward_mani call on each side."""
    (_, r_shape, r_layout, r_dict), (_, m_shape, m_layout, m_dict) = r, m
    r_vals = {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in r_dict.items()}
    m_vals = {k: float(v) for k, v in m_dict.items()}
    for name, rv, mv in (("Shape_loss", r_shape, m_shape), ("Layout_loss", r_layout, m_layout)):
        if torch.is_tensor(rv):
            r_vals[name], m_vals[name] = float(rv.detach()), float(mv)
    rows, worst = {}, 0.0
    print(f"[{tag}]")
    for k, rv in r_vals.items():
        mv = m_vals[k]
        rel = abs(rv - mv) / abs(rv) if rv != 0 else abs(mv)
        rows[k] = {"reference": rv, "b200": mv, "rel": rel}
        worst = max(worst, rel)
        print(f"  {k:16s} reference {rv:.8f}   b200 {mv:.8f}   rel {rel:.2e}")
    return rows, worst


def stage_checks(ref, mine, b, dev):
    """Each BatchNorm-carrying stage on its own, both sides under train(), identical inputs: localises a disagreement."""
    def rel(a, c):
        return float((a.double() - c.double()).abs().max() / a.double().abs().max().clamp_min(1e-30))
    out = {}
    n = len(b["objs"])
    gen = torch.Generator().manual_seed(9)
    with torch.no_grad():
        r = ref.diff.init_encoder(b["objs"], b["triples"], b["text"], b["rel"])
        m = mine.encoder.init_encoder(b["objs"], b["triples"], b["text"], b["rel"])
        out["init_encoder.latent"] = rel(r[2], m[2])
        lat = torch.cat([r[2], torch.randn(n, 64, generator=gen).to(dev)], dim=1)
        r2 = ref.diff.manipulate(lat, b["objs"], b["triples"], b["text"], b["rel"])
        m2 = mine.encoder.manipulate(lat, b["objs"], b["triples"], b["text"], b["rel"])
        out["manipulate.latent"] = rel(r2[0], m2[0])
        out["rel_s_mlp"] = rel(ref.diff.rel_s_mlp(r2[0]), mine.encoder.rel_s(r2[0]))
        # layout denoiser, one timestep per object
        x = torch.randn(n, 8, generator=gen).to(dev)
        t = torch.randint(0, 1000, (n,), generator=gen).to(dev)
        rl = ref.diff.LayoutDiff.df._denoise(x, r2[2], b["triples"], t, r2[0])
        ml = mine.layout._denoise(x, r2[2], b["triples"], t, r2[0])
        out["layout_denoiser"] = rel(rl, ml)
        # shape denoiser on the first two scenes
        k = 2 * (n // max(1, len(set(b["o2s"].tolist()))))
        tr = b["triples"][(b["triples"][:, 0] < k) & (b["triples"][:, 2] < k)]
        z = torch.randn(k, 3, 16, 16, 16, generator=gen).to(dev)
        uc = ref.diff.rel_s_mlp(r2[2]).unsqueeze(1)[:k]
        c = ref.diff.rel_s_mlp(r2[0]).unsqueeze(1)[:k]
        ref.diff.ShapeDiff.df.train()
        rs = ref.diff.ShapeDiff.apply_model(z, uc, tr, t[:k], c)
        ms = mine.unet3d(z, uc, tr, t[:k], context=c)
        out["shape_denoiser"] = rel(rs, ms)
    for k_, v in out.items():
        print(f"  stage {k_:22s} max-rel {v:.2e}")
    return out


def main(args):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    import tempfile
    workdir = tempfile.mkdtemp(prefix="trainfwd_")          # the synthetic VQ-VAE checkpoint and the reference's log directory
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    rb.setup_reference()
    args.device, args.layout_steps = "cuda", 1000
    ref, cfg = rb.build(args, workdir)
    print("re-drawn zero-initialised tensors:", rb.redraw_zero_init(ref))
    ref = ref.to(dev)
    b = batch(args.scenes, args.nodes, dev)
    manipulated = [1, args.nodes + 2]

    ref.train()
    # echo2shape.py:317 indexes the CPU tensor `logvar` with CUDA timesteps, which the torch of this image (2.11) refuses; the
    # attribute is moved to the GPU on the instance -- the reference's files stay untouched and the values (all 0) are the same
    ref.diff.ShapeDiff.logvar = ref.diff.ShapeDiff.logvar.to(dev)
    call(ref, b, manipulated)
    r1, r_sec = call(ref, b, manipulated)
    # an addition (one node missing from the encoder-side scene) with replace_latent = False: only the touched rows are replaced
    ref.diff.replace_all_latent = False
    r2, _ = call(ref, b, [args.nodes + 3], missing=[2])
    ref.diff.replace_all_latent = True

    # ---- the B200 arm: its own components from the same YAML files, the reference's weights as a checkpoint dict
    from echoscene_b200 import scene, sgdiff
    ckpt = {k: v.detach() for k, v in torch.nn.Module.state_dict(ref.diff).items() if torch.is_tensor(v)}
    ckpt["shape_df"] = ref.diff.ShapeDiff.df.state_dict()
    ckpt["vqvae"] = ref.diff.ShapeDiff.vqvae.state_dict()
    kw = dict(replace_latent=True, with_changes=True, residual=True, gconv_pooling="avg", with_angles=True, clip=True, separated=False,
              precision=args.precision, config_dir=os.path.join(rb.REF, "config"))
    mine = sgdiff.SGDiff("echoscene", cfg, ref.vocab, with_vq_encoder=True, **kw)
    info = scene.load_reference_checkpoint(ckpt, encoder=mine.encoder, unet1d=mine.unet1d, unet3d=mine.unet3d, vqvae=mine.vqvae)
    print("loaded:", info["loaded"])
    mine = mine.cuda().train()
    stages = stage_checks(ref, mine, b, dev)
    call(mine, b, manipulated)
    m1, m_sec = call(mine, b, manipulated)
    mine.diff.replace_all_latent = False
    m2, _ = call(mine, b, [args.nodes + 3], missing=[2])

    rows1, w1 = loss_table("three scenes, two manipulated nodes, replace_latent", r1, m1)
    rows2, w2 = loss_table("one missing node + one manipulated node, touched rows only", r2, m2)
    same_sel = bool(torch.equal(r1[0].cpu(), m1[0].cpu()) and torch.equal(r2[0].cpu(), m2[0].cpu()))
    print(f"objects selected for the shape branch: {len(r1[0])} / {len(b['objs'])}, identical: {same_sel}")
    print(f"forward_mani: reference (eager, autograd tape recorded) {r_sec * 1e3:.1f} ms, b200 (values only) {m_sec * 1e3:.1f} ms")
    del ref, mine
    torch.cuda.empty_cache()

    # ---- the layout-only model (SGDiff('echolayout'), model/EchoLayout.py:247-289)
    refl, cfgl = rb.build(args, workdir, model_type="echolayout")
    rb.redraw_zero_init(refl)
    refl = refl.to(dev).train()
    r3, _ = call(refl, b, manipulated, missing=[2], layout_only=True)
    minel = sgdiff.SGDiff("echolayout", cfgl, refl.vocab, **kw)
    ckl = {k: v.detach() for k, v in torch.nn.Module.state_dict(refl.diff).items() if torch.is_tensor(v)}
    print("loaded:", scene.load_reference_checkpoint(ckl, encoder=minel.encoder, unet1d=minel.unet1d)["loaded"])
    minel = minel.cuda().train()
    m3, _ = call(minel, b, manipulated, missing=[2], layout_only=True)
    rows3, w3 = loss_table("layout-only model (echolayout), one missing + two manipulated nodes", r3, m3)

    worst = max(w1, w2, w3)
    print(f"worst deviation {worst:.3e}")
    res = {"stages": stages, "losses": rows1, "losses_addition": rows2, "losses_layout_only": rows3, "worst_rel": worst,
           "selected_identical": same_sel, "scenes": args.scenes, "nodes_per_scene": args.nodes,
           "manipulated_nodes": manipulated, "precision": args.precision, "reference_ms": r_sec * 1e3, "b200_ms": m_sec * 1e3,
           "note": "reference = unmodified baseline/_ref under model.train(), eager fp32 (TF32 off) on the same GPU, building its autograd "
                   "tape; b200 = forward values only (no tape, no backward)"}
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    return 0 if (worst < args.tol and same_sel) else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=3)
    ap.add_argument("--nodes", type=int, default=7)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--tol", type=float, default=1e-3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trainfwd.json"))
    sys.exit(main(ap.parse_args()))
