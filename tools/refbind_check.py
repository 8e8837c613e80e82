"""The reference-side binding, run for real (VERDICT r1 task 6): the UNMODIFIED reference (baseline/_ref) builds its own
SGDiff('echoscene', config/full_mp.yaml) and samples a scene twice -- once as it is (eager PyTorch, TF32 off), once after
echoscene_b200.integrate.patch_reference() has rebound its hot-path classes to libechoscene_b200.so -- from the same
synthetic checkpoint, scene and RNG seed.  Each arm runs in its own process (the patch is global):

  python tools/refbind_check.py --arm reference --out gpurun_out/refbind_ref.pt
  python tools/refbind_check.py --arm patched   --out gpurun_out/refbind_b200.pt --like gpurun_out/refbind_ref.pt
  python tools/refbind_check.py --compare gpurun_out/refbind_ref.pt gpurun_out/refbind_b200.pt

TEST / MEASUREMENT INFRASTRUCTURE: imports baseline/_ref, never imported by the package.
The layout chain is cut to --layout-steps DDPM steps and the shape chain to the reference's own debug setting (misc.debug = 1:
7 DDIM steps, echo2shape.py:116-120) so that the eager arm finishes in a minute; the code paths are the full ones.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")


class AttrDict(dict):
    """dict with attribute access, recursively (what the reference expects of an OmegaConf node)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k) from None
        return v

    def __setattr__(self, k, v):
        self[k] = v


def wrap(x):
    if isinstance(x, dict):
        return AttrDict({k: wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


class _Loader(yaml.SafeLoader):
    """SafeLoader that reads `1e-4` as a float, as OmegaConf's loader does (plain PyYAML wants `1.0e-4` and returns a string)."""


import re  # noqa: E402
_Loader.add_implicit_resolver("tag:yaml.org,2002:float", re.compile(
    r"^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?|[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)|\.[0-9_]+(?:[eE][-+][0-9]+)?"
    r"|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"), list("-+0123456789."))


def load_yaml(path):
    with open(path) as f:
        return wrap(yaml.load(f, Loader=_Loader))


def setup_reference():
    from baseline import ref_runner
    ref_runner.load_reference()          # sys.path + stubs for absent optional packages
    om = sys.modules["omegaconf"]
    om.OmegaConf.load = staticmethod(load_yaml)
    om.OmegaConf.create = staticmethod(wrap)
    if os.path.join(REF, "scripts") not in sys.path:
        sys.path.insert(0, os.path.join(REF, "scripts"))


def build(args, workdir, model_type="echoscene"):
    """-> (SGDiff instance of the reference, config)."""
    import importlib
    cfg = load_yaml(os.path.join(REF, "config", "full_mp.yaml"))
    cfg.hyper.device = args.device
    cfg.hyper.batch_size = 16
    cfg.hyper.logs_dir = cfg.hyper.results_dir = os.path.join(workdir, "logs")
    cfg.layout_branch.diffusion_kwargs.time_num = args.layout_steps
    cfg.shape_branch.df_cfg = os.path.join(REF, "config", "sdfusion-txt2shape_mp.yaml")
    cfg.shape_branch.vq_cfg = os.path.join(REF, "config", "vqvae_snet.yaml")
    cfg.shape_branch.vq_ckpt = os.path.join(workdir, "vqvae_synth.pth")
    cfg.misc.debug = 1                    # the reference's own 7-step DDIM setting
    vq_conf = load_yaml(cfg.shape_branch.vq_cfg)
    if not os.path.exists(cfg.shape_branch.vq_ckpt):
        VQVAE = importlib.import_module("model.networks.vqvae_networks.network").VQVAE
        mp = vq_conf.model.params
        torch.manual_seed(5)
        vq = VQVAE(mp.ddconfig, mp.n_embed, mp.embed_dim)
        torch.save({"vqvae": vq.state_dict()}, cfg.shape_branch.vq_ckpt)
    vocab = {"object_idx_to_name": ["_scene_"] + [f"c{i}" for i in range(35)], "pred_idx_to_name": ["in"] + [f"p{i}" for i in range(15)],
             "object_idx_to_name_grained": ["x"]}
    # the mesh renderer (pytorch3d, absent here) is built in EchoToShape.__init__ but only used by visualisation code
    for modname in ("model.networks.diffusion_shape.diff_utils.util_3d", "model.networks.diffusion_shape.echo2shape"):
        mod = importlib.import_module(modname)
        if hasattr(mod, "init_mesh_renderer"):
            mod.init_mesh_renderer = lambda *a, **k: None
    # EchoToShape.rel2shape seeds its x_T from the wall clock (echo2shape.py:502): pin the clock it sees so that both arms draw the
    # same noise (the reference's file is untouched; only the `time` name its module looks up is replaced for this process)
    import types
    importlib.import_module("model.networks.diffusion_shape.echo2shape").time = types.SimpleNamespace(time=lambda: 1700000000)
    SGDiff = importlib.import_module("model.SGDiff").SGDiff
    torch.manual_seed(11)
    model = SGDiff(model_type, cfg, vocab, replace_latent=True, with_changes=True, residual=True, gconv_pooling="avg",
                   with_angles=True, clip=True, separated=False)
    return model, cfg


def redraw_zero_init(model):
    """The reference zero-initialises the last conv of every ResBlock, proj_out and out.2: with random weights the whole UNet
    would be the identity on its input.  Re-draw them N(0, 0.02) so that parity is not vacuous (as oracle/cases.py does)."""
    g = torch.Generator().manual_seed(21)
    n = 0
    import itertools
    # EchoToShape is not an nn.Module: its denoiser's parameters are not among the model's
    shape_df = model.diff.ShapeDiff.df.named_parameters() if hasattr(model.diff, "ShapeDiff") else []
    for name, p in itertools.chain(model.named_parameters(), shape_df):
        if p.dim() >= 1 and float(p.detach().abs().max()) == 0.0:
            with torch.no_grad():
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
            n += 1
    for name, b in model.named_buffers():
        if name.endswith("running_mean"):
            b.copy_(torch.randn(b.shape, generator=g) * 0.1)
        elif name.endswith("running_var"):
            b.copy_(torch.rand(b.shape, generator=g) + 0.5)
    return n


def scene(args, dev):
    from echoscene_b200 import synth
    g, objs, text, rel = synth.scene_inputs(args.nodes, 4 * args.nodes, 3)
    return objs.to(dev), g.triples.to(dev), text.to(dev), rel.to(dev)


def run(args):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    workdir = os.path.dirname(os.path.abspath(args.out))
    os.makedirs(workdir, exist_ok=True)
    setup_reference()
    patched = None
    if args.arm == "patched":
        from echoscene_b200 import integrate
        patched = integrate.patch_reference(precision=args.precision)
    model, cfg = build(args, workdir)
    dev = torch.device(args.device)
    if args.like:
        like = torch.load(args.like, map_location="cpu")
        missing = torch.nn.Module.load_state_dict(model.diff, like["diff_state"], strict=True)   # (the model overrides state_dict / load_state_dict for its checkpoints)
        model.diff.ShapeDiff.df.load_state_dict(like["shape_df_state"], strict=True)
        model.diff.ShapeDiff.vqvae.load_state_dict(like["vqvae_state"], strict=False)
        print("loaded the reference arm's weights:", missing)
    else:
        print("re-drawn zero-initialised tensors:", redraw_zero_init(model))
    model = model.to(dev).eval()
    objs, triples, text, rel = scene(args, dev)
    out = {}
    # record the latents the DDIM chain hands to the VQ-VAE decoder (a near-tie of the codebook search may quantise a 1e-7
    # difference to another code, so the latents are the robust comparison; the decoded SDFs are reported next to them)
    vq = model.diff.ShapeDiff.vqvae_module
    seen = {}
    inner = vq.decode_no_quant

    def recording_decode(h, *a, **k):
        seen["latents"] = h.detach().float().cpu()
        return inner(h, *a, **k)
    vq.decode_no_quant = recording_decode
    with torch.no_grad():
        for rep in range(2):   # the second pass is the timed one
            torch.manual_seed(1234)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = model.sample_box_and_shape(objs, triples, text, rel, gen_shape=True)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            out["seconds"] = time.perf_counter() - t0
    out["result"] = {k: v.detach().float().cpu() for k, v in res.items() if torch.is_tensor(v)}
    out["result"].update(seen)
    if args.like and "latents" in like.get("result", {}):
        # the decoder on IDENTICAL input: the other arm's latents through this arm's decode_no_quant (the end-to-end `shapes` differ
        # wherever a 1e-6 difference of the latents falls on a codebook near-tie; that is the quantiser, not the decoder)
        with torch.no_grad():
            out["result"]["shapes_of_reference_latents"] = inner(like["result"]["latents"].to(dev)).detach().float().cpu()
    out["arm"], out["patched"] = args.arm, patched
    if not args.like:
        out["diff_state"] = {k: v.detach().cpu() for k, v in torch.nn.Module.state_dict(model.diff).items() if torch.is_tensor(v)}
        out["shape_df_state"] = {k: v.detach().cpu() for k, v in model.diff.ShapeDiff.df.state_dict().items()}
        out["vqvae_state"] = {k: v.detach().cpu() for k, v in model.diff.ShapeDiff.vqvae.state_dict().items()}
    torch.save(out, args.out)
    print(f"[{args.arm}] sample_box_and_shape: {out['seconds']:.2f} s; outputs: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out["result"].items()))


def compare(a, b):
    A, B = torch.load(a, map_location="cpu"), torch.load(b, map_location="cpu")
    worst = 0.0
    if "shapes_of_reference_latents" in B["result"]:
        A["result"]["shapes_of_reference_latents"] = A["result"]["shapes"]
    for k, va in A["result"].items():
        vb = B["result"][k]
        rel = float((va.double() - vb.double()).abs().max() / va.double().abs().max().clamp_min(1e-30))
        l2 = float((va.double() - vb.double()).norm() / va.double().norm().clamp_min(1e-30))
        note = ""
        if k == "shapes" and "shapes_of_reference_latents" in B["result"]:
            # informational: end to end, latents that differ by ~5e-6 fall on different codebook entries at near-ties
            off = float(((va.double() - vb.double()).abs() > 1e-3 * va.double().abs().max()).double().mean())
            note = f"  (end to end, incl. codebook near-ties: {100 * off:.2f} % of voxels off by > 1e-3; not judged)"
        else:
            worst = max(worst, rel, l2)
        print(f"  {k:14s} {tuple(va.shape)}  max-rel {rel:.3e}  rel-L2 {l2:.3e}{note}")
    print(f"reference arm {A['seconds']:.2f} s, patched arm {B['seconds']:.2f} s  ({A['seconds'] / B['seconds']:.1f}x);  worst deviation {worst:.3e}")
    return worst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["reference", "patched"])
    ap.add_argument("--out")
    ap.add_argument("--like", default=None, help="take the weights saved by the other arm")
    ap.add_argument("--compare", nargs=2)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--nodes", type=int, default=6)
    ap.add_argument("--layout-steps", type=int, default=20)
    a = ap.parse_args()
    if a.compare:
        sys.exit(0 if compare(*a.compare) < 1e-3 else 1)
    run(a)
