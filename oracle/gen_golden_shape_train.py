"""Training-mode golden vectors of the SHAPE denoiser (groundwork for the trunk's backward pass): the
reference's UNet3DModel under .train() -- shape_code_graph_cov's BatchNorm1d layers on batch statistics, blocks behind the gradient
checkpoint wrapper -- imported in place from /root/reference (build container only), on one scene of three objects with one
timestep per object (echo2shape.py:359), differentiated by torch autograd for the loss of p_losses (mean squared error against the
noise, echo2shape.py:314-319 with logvar = 0).

Stored (tests/golden/shape_train.pt): the forward output, and per parameter a DIGEST of its gradient (L2 norm + eight entries; the
full gradient is the size of the model, 430 M values).  Pins oracle.unet3d_forward(batch_stats=True) and autograd over it -- the
oracle of the conv dgrad / wgrad, GroupNorm, attention and GEGLU backward kernels to come -- against the reference's autograd.
Usage: python oracle/gen_golden_shape_train.py   (about two minutes and ~20 GB of host memory)"""
import os
import sys
import time

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from echoscene_b200 import arch, synth  # noqa: E402
from oracle import cases, echoscene_oracle as orc, ref_import  # noqa: E402
from oracle.gen_golden_layout_train import digest  # noqa: E402

N_OBJ, N_TRI = 3, 6


def inputs(scfg):
    g = synth.make_scene_graph(N_OBJ, N_TRI, 81)
    gen = torch.Generator().manual_seed(83)
    uc = torch.randn(N_OBJ, 1, scfg.context_dim, generator=gen)
    x = torch.randn(N_OBJ, scfg.in_channels, scfg.image_size, scfg.image_size, scfg.image_size, generator=gen)
    t = torch.tensor([907, 412, 33]).long()
    noise = torch.randn(x.shape, generator=gen)
    return g, uc, x, t, noise


def oracle_grads(sd, scfg, g, uc, x, t, noise):
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    out = orc.unet3d_forward(leaf, scfg, x, uc, g.triples, t, batch_stats=True)
    loss = ((noise - out) ** 2).mean()
    loss.backward()
    return out.detach(), float(loss.detach()), {k: v.grad for k, v in leaf.items() if torch.is_tensor(v) and v.requires_grad}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    ref = ref_import.load()
    with open(os.path.join(ref_import.REF_ROOT, "config/sdfusion-txt2shape_mp.yaml")) as f:
        sdf = yaml.safe_load(f)
    scfg = cases.shape_cfg()
    ukw = dict(sdf["unet"]["params"], conditioning_key=sdf["model"]["params"]["conditioning_key"])
    m = ref.UNet3DModel(**ukw)
    sd = arch.make_state_dict(arch.unet3d_specs(scfg), cases.WEIGHT_SEED_SHAPE)
    m.load_state_dict(sd, strict=True)
    m.train()
    g, uc, x, t, noise = inputs(scfg)
    t0 = time.time()
    out = m(x, uc, g.triples, t, context=uc)
    loss = ((noise - out) ** 2).mean()
    loss.backward()
    t_ref = time.time() - t0
    ref_grads = {k: p.grad for k, p in m.named_parameters()}
    del m
    t0 = time.time()
    o_out, o_loss, o_grads = oracle_grads(sd, scfg, g, uc, x, t, noise)
    t_orc = time.time() - t0
    worst_out = float((o_out - out.detach()).abs().max() / out.detach().abs().max())
    scale = max(float(gr.abs().max()) for gr in ref_grads.values() if gr is not None)
    worst, wk, missing, noise_level, devs = 0.0, "", [], [], []
    for k, gr in ref_grads.items():
        if gr is None:
            missing.append(k)
            continue
        og = o_grads[k]
        if float(gr.abs().max()) < 1e-6 * scale:
            noise_level.append(k)
            assert float(og.abs().max()) < 1e-5 * scale, k
            continue
        r = float((og.double() - gr.double()).norm() / gr.double().norm().clamp_min(1e-30))
        devs.append((r, k, float(gr.abs().max()) / scale))
        if r > worst:
            worst, wk = r, k
    for r, k, rel_size in sorted(devs, reverse=True)[:6]:
        print(f"  {r:.2e}  {k}  (max |grad| = {rel_size:.1e} of the largest)")
    sizable = [d for d in devs if d[2] > 1e-4]
    print(f"  parameters whose gradient is above 1e-4 of the largest: {len(sizable)}, worst rel-L2 {max(d[0] for d in sizable):.2e}")
    gold = {"out": out.detach(), "loss": float(loss.detach()), "grads": {k: digest(gr) for k, gr in ref_grads.items() if gr is not None},
            "no_grad": missing, "noise_level": noise_level, "grad_scale": scale}
    path = os.path.join(ROOT, "tests", "golden", "shape_train.pt")
    torch.save(gold, path)
    print(f"UNet3DModel under .train() ({N_OBJ} objects; reference fwd+bwd {t_ref:.0f} s, oracle {t_orc:.0f} s): oracle(batch_stats=True) forward "
          f"vs reference max-rel {worst_out:.3e}; loss {float(loss.detach()):.6f} vs {o_loss:.6f}; autograd over the oracle vs the reference's "
          f"autograd, worst per-parameter rel-L2 {worst:.3e} ({wk}) over {len(gold['grads']) - len(noise_level)} parameters "
          f"({len(noise_level)} mathematically zero, {len(missing)} without gradient: {missing}); fixture {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
