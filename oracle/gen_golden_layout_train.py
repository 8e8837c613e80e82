"""Training-mode golden vectors of the LAYOUT denoiser (groundwork for the trunk's backward pass): the
reference's UNet1DModel under .train() -- box_graph_cov's BatchNorm1d layers on batch statistics, every block behind the
reference's gradient-checkpoint wrapper (`use_checkpoint: true`) -- imported in place from /root/reference (build container only),
on a collated batch of two scenes with one timestep per scene (get_loss_iter, diffusion_ddpm.py:597-608), differentiated by torch
autograd for the loss the training step uses (mean squared error against the noise).

Stored (tests/golden/layout_train.pt): the forward output, and for every parameter a DIGEST of its gradient -- its L2 norm and
eight entries at fixed positions -- because the full gradient is the size of the model (164 M values).  Pins
oracle.unet1d_forward(batch_stats=True) and autograd over it (the oracle of the trunk's backward) against the reference's autograd.
Usage: python oracle/gen_golden_layout_train.py"""
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from echoscene_b200 import arch, synth  # noqa: E402
from oracle import cases, echoscene_oracle as orc, ref_import  # noqa: E402


def inputs(lcfg):
    g = synth.batch_scene_graphs([synth.make_scene_graph(6, 18, 71), synth.make_scene_graph(8, 24, 72)])
    gen = torch.Generator().manual_seed(73)
    n = g.n_nodes
    obj_embed = torch.randn(n, lcfg.obj_embed_dim, generator=gen)
    x = torch.randn(n, lcfg.in_channels, generator=gen)
    t = torch.cat([torch.full((6,), 412), torch.full((8,), 37)]).long()           # one timestep per scene
    noise = torch.randn(n, lcfg.in_channels, generator=gen)
    return g, obj_embed, x, t, noise


def digest(grad: torch.Tensor):
    flat = grad.detach().reshape(-1)
    idx = (torch.arange(8, dtype=torch.int64) * (flat.numel() - 1)) // 7
    return {"norm": float(flat.double().norm()), "samples": flat[idx].clone()}


def oracle_grads(sd, lcfg, g, obj_embed, x, t, noise):
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    out = orc.unet1d_forward(leaf, lcfg, x, obj_embed, g.triples, t, batch_stats=True).squeeze(-1)
    loss = ((noise - out) ** 2).mean()
    loss.backward()
    return out.detach(), float(loss.detach()), {k: v.grad for k, v in leaf.items() if torch.is_tensor(v) and v.requires_grad}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    ref = ref_import.load()
    with open(os.path.join(ref_import.REF_ROOT, "config/full_mp.yaml")) as f:
        full = yaml.safe_load(f)
    lcfg = cases.layout_cfg()
    kw = dict(full["layout_branch"]["denoiser_kwargs"])
    m = ref.UNet1DModel(**kw)
    sd = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
    m.load_state_dict(sd, strict=True)
    m.train()
    g, obj_embed, x, t, noise = inputs(lcfg)
    out = m(x, obj_embed, g.triples, t, None).squeeze(-1)
    loss = ((noise - out) ** 2).mean()
    loss.backward()
    ref_grads = {k: p.grad for k, p in m.named_parameters()}
    o_out, o_loss, o_grads = oracle_grads(sd, lcfg, g, obj_embed, x, t, noise)
    worst_out = float((o_out - out.detach()).abs().max() / out.detach().abs().max())
    worst, missing, noise_level = 0.0, [], []
    scale = max(float(gr.abs().max()) for gr in ref_grads.values() if gr is not None)
    for k, gr in ref_grads.items():
        if gr is None:
            missing.append(k)
            continue
        og = o_grads[k]
        if float(gr.abs().max()) < 1e-6 * scale:
            # mathematically zero: a bias feeding BatchNorm on batch statistics, or attn1.to_q / to_k (softmax over ONE token is 1
            # whatever q and k are, attention.py:203-217 with L = 1): rounding noise on both sides
            noise_level.append(k)
            assert float(og.abs().max()) < 1e-5 * scale, k
            continue
        worst = max(worst, float((og.double() - gr.double()).norm() / gr.double().norm().clamp_min(1e-30)))
    gold = {"out": out.detach(), "loss": float(loss), "grads": {k: digest(gr) for k, gr in ref_grads.items() if gr is not None},
            "no_grad": missing, "noise_level": noise_level, "grad_scale": scale}
    torch.save(gold, os.path.join(ROOT, "tests", "golden", "layout_train.pt"))
    print(f"UNet1DModel under .train(): oracle(batch_stats=True) forward vs reference max-rel {worst_out:.3e}; loss {float(loss):.6f} vs "
          f"{o_loss:.6f}; autograd over the oracle vs the reference's autograd, worst per-parameter rel-L2 {worst:.3e} over "
          f"{len(gold['grads']) - len(noise_level)} parameters ({len(noise_level)} mathematically zero, {len(missing)} without gradient: {missing}); fixture "
          f"{os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'layout_train.pt')) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
