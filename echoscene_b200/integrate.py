"""Reference-side binding: swap the hot-path classes inside an importable checkout of ymxlzgy/echoscene.

The reference has no plugin/FFI boundary on the denoising path; its seam is the Python class surface.
``patch_reference()`` rebinds, in the reference's own module namespaces, exactly the names its constructors look up:

  model.graph.GraphTripleConv / GraphTripleConvNet                           (model/graph.py:89, 214)
  model.networks.diffusion_layout.denoise_net.UNet1DModel                    (denoise_net.py:451; used by echo2layout.py:15)
  model.networks.diffusion_layout.echo2layout.UNet1DModel / DiffusionPoint   (echo2layout.py:3-4, 15-30)
  model.networks.diffusion_shape.openai_model_3d.UNet3DModel                 (openai_model_3d.py:452)
  model.networks.diffusion_shape.network.UNet3DModel / DiffusionUNet         (network.py:9-17)
  model.networks.diffusion_shape.echo2shape.DDIMSampler                      (echo2shape.py:46, 122)
  model.model_utils.VQVAE (looked up by load_vqvae, model_utils.py:7-32)      decode path only: rel2shape's
                                                                             decode_no_quant (echo2shape.py:522)

  model.EchoScene.Sg2ScDiffModel.init_encoder / manipulate                   (EchoScene.py:143-157, 181-195) -> one
                                                                             echo_scene_* call each (scene_encoders=True)

After that, ``scripts/eval_3dfront.py`` builds ``SGDiff`` as before (same YAML, same checkpoints: state_dict keys are
identical), and ``Sg2ScDiffModel.sample`` runs both chains through libechoscene_b200.so.  The one-time scene encoders
 run through ``modules.SceneEncoder``: the model's ``init_encoder`` / ``manipulate`` methods are rebound to
single C-ABI calls on a ``SceneEncoder`` that SHARES the model's own parameters (no copy; rebuilt when they are replaced),
so ``sample`` / ``sample_with_changes`` / ``sample_with_additions`` run unchanged; ``self.rel_s_mlp(x)`` is routed to
echo_scene_rel_s the first time the encoders run (its parameters stay where they are).  ``echoscene_b200.scene`` /
``echoscene_b200.sgdiff`` are the same surface without any reference code.

Training (`train_3dfront.py`) needs an autograd tape through every module, and the denoiser trunks have no backward pass on
this path yet: the classes handed to the reference raise in ``train()`` mode (``train_forward_values = False``) instead of
returning tape-less values.  (The package's own facade, ``sgdiff.SGDiff``, does compute the training forward's loss VALUES in
``train()`` mode: ``forward_mani``.)
"""
from __future__ import annotations

import importlib
from typing import Optional


def scene_encoder_of(model):
    """The ``modules.SceneEncoder`` that runs ``model``'s (a reference Sg2ScDiffModel / Sg2BoxDiffModel instance) encoder
    sub-modules.  Its parameters ARE the model's tensors (``load_state_dict(assign=True)``); it is rebuilt when the model's
    parameters were replaced (``.cuda()``, ``load_networks``).  Kept outside the model's module tree, so the reference's
    ``state_dict`` / checkpoints do not change."""
    import torch.nn as nn
    from . import modules

    # the layout-only model embeds the predicates of `manipulate` with its own table (EchoLayout.py:154)
    man_dc = bool(getattr(type(model), "_echo_man_dc_preds", False))
    pre = modules.SceneEncoder.PREFIXES + (("pred_embeddings_man_dc.",) if man_dc else ())
    sd = {k: v for k, v in nn.Module.state_dict(model, keep_vars=True).items() if k.startswith(pre)}
    key = tuple(v.data_ptr() for v in sd.values())
    cached = model.__dict__.get("_echo_scene_encoder")
    if cached is not None and cached[0] == key:
        return cached[1]
    if "rel_s_mlp.3.weight" in sd and tuple(sd["rel_s_mlp.3.weight"].shape) != (1280, 960):
        from ._lib import EchoError
        raise EchoError("rel_s_mlp of the 'concat' conditioning variant ([640, 1280, 4096], EchoScene.py:98-99) is outside the hot path")
    enc = modules.SceneEncoder(num_objs=model.obj_embeddings_ec.weight.shape[0] - 1,
                               num_preds=model.pred_embeddings_ec.weight.shape[0], embedding_dim=model.embedding_dim,
                               gconv_num_layers=model.gconv_net_ec.num_layers,
                               residual=any(k.endswith("linear_projection.weight") for k in sd), use_clip=bool(model.clip),
                               with_rel_s=any(k.startswith("rel_s_mlp.") for k in sd),   # the layout-only model has none
                               man_dc_preds=man_dc)
    enc.load_state_dict(sd, strict=True, assign=True)
    enc.eval()
    object.__setattr__(model, "_echo_scene_encoder", (key, enc))     # not a registered sub-module
    if enc.with_rel_s:
        _route_rel_s(model)
    return enc


def _route_rel_s(model):
    """``self.rel_s_mlp(x)`` in Sg2ScDiffModel.sample* (EchoScene.py:405-410) -> echo_scene_rel_s.  The nn.Sequential keeps its
    parameters (state_dict unchanged); only its ``forward`` is rebound, and only for eval mode: under ``model.train()`` the
    reference's own forward runs (autograd)."""
    mlp = model.rel_s_mlp
    if "_echo_forward" in mlp.__dict__:
        return
    original = mlp.forward

    def forward(x, _model=model, _original=original):
        if _model.training or mlp.training:
            return _original(x)
        return scene_encoder_of(_model).rel_s(x)

    object.__setattr__(mlp, "_echo_forward", original)
    object.__setattr__(mlp, "forward", forward)


def _eval_only(model):
    if model.training:   # the SceneEncoder is outside the module tree, so model.train() cannot reach its own eval check
        from ._lib import EchoError
        raise EchoError("init_encoder / manipulate on the B200 path are eval-mode only (BatchNorm running statistics, no "
                        "autograd); patch_reference(scene_encoders=False) keeps the reference's methods for training")


def _init_encoder(self, objs, triples, enc_text_feat, enc_rel_feat):
    """Sg2ScDiffModel.init_encoder (EchoScene.py:143-157) as one echo_scene_init_encoder call."""
    _eval_only(self)
    return scene_encoder_of(self).init_encoder(objs, triples, enc_text_feat, enc_rel_feat)


def _manipulate(self, latent_f, objs, triples, dec_text_feat, dec_rel_feat):
    """Sg2ScDiffModel.manipulate (EchoScene.py:181-195) as one echo_scene_manipulate call."""
    _eval_only(self)
    return scene_encoder_of(self).manipulate(latent_f, objs, triples, dec_text_feat, dec_rel_feat)


def patch_reference(precision: str = "fp32", ddim_steps: Optional[int] = None, vqvae_decode: bool = True,
                    scene_encoders: bool = True) -> dict:
    """Call AFTER the reference checkout is importable (sys.path) and BEFORE SGDiff(...) is constructed.
    ``vqvae_decode=False`` keeps the reference's VQVAE (needed for training, which encodes: echo2shape.py:349);
    ``scene_encoders=False`` keeps the reference's init_encoder / manipulate (they then run layer by layer on the patched
    GraphTripleConvNet)."""
    from . import modules, samplers

    def _with_precision(cls):
        class _P(cls):
            train_forward_values = False   # the reference's training loop needs an autograd tape: refuse train() instead of returning values

            def __init__(self, *a, **k):
                k.setdefault("precision", precision)
                super().__init__(*a, **k)
        _P.__name__ = cls.__name__
        _P.__qualname__ = cls.__qualname__
        return _P

    def _eval_only_class(cls):
        class _E(cls):
            train_forward_values = False
        _E.__name__ = cls.__name__
        _E.__qualname__ = cls.__qualname__
        return _E

    GTC, GTCN = _eval_only_class(modules.GraphTripleConv), _eval_only_class(modules.GraphTripleConvNet)

    U1, U3 = _with_precision(modules.UNet1DModel), _with_precision(modules.UNet3DModel)

    class DiffusionUNet(modules.DiffusionUNet):
        def __init__(self, unet_params, vq_conf=None, conditioning_key=None):
            super().__init__(unet_params, vq_conf=vq_conf, conditioning_key=conditioning_key, precision=precision)

    VQ = _with_precision(modules.VQVAE)
    done = {}
    targets = [
        ("model.graph", {"GraphTripleConv": GTC, "GraphTripleConvNet": GTCN}),
        ("model.networks.diffusion_layout.denoise_net", {"UNet1DModel": U1}),
        ("model.networks.diffusion_layout.echo2layout", {"UNet1DModel": U1, "DiffusionPoint": samplers.DiffusionPoint}),
        ("model.networks.diffusion_shape.openai_model_3d", {"UNet3DModel": U3}),
        ("model.networks.diffusion_shape.network", {"UNet3DModel": U3, "DiffusionUNet": DiffusionUNet}),
        ("model.networks.diffusion_shape.echo2shape", {"DDIMSampler": samplers.DDIMSampler, "DiffusionUNet": DiffusionUNet}),
        ("model.EchoScene", {"GraphTripleConvNet": GTCN}),
        ("model.EchoLayout", {"GraphTripleConvNet": GTCN}),
    ]
    if vqvae_decode:
        targets.append(("model.model_utils", {"VQVAE": VQ}))
    for modname, names in targets:
        try:
            mod = importlib.import_module(modname)
        except Exception as e:   # optional module of the checkout missing: report, do not hide
            done[modname] = f"not patched: {e!r}"
            continue
        for n, obj in names.items():
            if hasattr(mod, n):
                setattr(mod, n, obj)
        done[modname] = sorted(names)
    if scene_encoders:
        for modname, cls in (("model.EchoScene", "Sg2ScDiffModel"), ("model.EchoLayout", "Sg2BoxDiffModel")):
            try:
                c = getattr(importlib.import_module(modname), cls)
            except Exception as e:
                done[modname + "." + cls] = f"not patched: {e!r}"
                continue
            c.init_encoder, c.manipulate = _init_encoder, _manipulate
            c._echo_man_dc_preds = cls == "Sg2BoxDiffModel"
            done[modname + "." + cls] = ["init_encoder", "manipulate"]
    return done
