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

After that, ``scripts/eval_3dfront.py`` builds ``SGDiff`` as before (same YAML, same checkpoints: state_dict keys are
identical), and ``Sg2ScDiffModel.sample`` runs both chains through libechoscene_b200.so.  Only the *denoiser-step*
classes are replaced; the one-time scene encoders keep using whatever ``model.graph`` provides — i.e. the CUDA
GraphTripleConvNet too, since EchoScene.py imports it from model.graph (EchoScene.py:5).

Training (`train_3dfront.py`) needs the backward pass, which is outside this round's scope: the patched classes raise
in ``train()`` mode instead of silently computing something else.
"""
from __future__ import annotations

import importlib
from typing import Optional


def patch_reference(precision: str = "fp32", ddim_steps: Optional[int] = None, vqvae_decode: bool = True) -> dict:
    """Call AFTER the reference checkout is importable (sys.path) and BEFORE SGDiff(...) is constructed.
    ``vqvae_decode=False`` keeps the reference's VQVAE (needed for training, which encodes: echo2shape.py:349)."""
    from . import modules, samplers

    def _with_precision(cls):
        class _P(cls):
            def __init__(self, *a, **k):
                k.setdefault("precision", precision)
                super().__init__(*a, **k)
        _P.__name__ = cls.__name__
        _P.__qualname__ = cls.__qualname__
        return _P

    U1, U3 = _with_precision(modules.UNet1DModel), _with_precision(modules.UNet3DModel)

    class DiffusionUNet(modules.DiffusionUNet):
        def __init__(self, unet_params, vq_conf=None, conditioning_key=None):
            super().__init__(unet_params, vq_conf=vq_conf, conditioning_key=conditioning_key, precision=precision)

    VQ = _with_precision(modules.VQVAE)
    done = {}
    targets = [
        ("model.graph", {"GraphTripleConv": modules.GraphTripleConv, "GraphTripleConvNet": modules.GraphTripleConvNet}),
        ("model.networks.diffusion_layout.denoise_net", {"UNet1DModel": U1}),
        ("model.networks.diffusion_layout.echo2layout", {"UNet1DModel": U1, "DiffusionPoint": samplers.DiffusionPoint}),
        ("model.networks.diffusion_shape.openai_model_3d", {"UNet3DModel": U3}),
        ("model.networks.diffusion_shape.network", {"UNet3DModel": U3, "DiffusionUNet": DiffusionUNet}),
        ("model.networks.diffusion_shape.echo2shape", {"DDIMSampler": samplers.DDIMSampler, "DiffusionUNet": DiffusionUNet}),
        ("model.EchoScene", {"GraphTripleConvNet": modules.GraphTripleConvNet}),
        ("model.EchoLayout", {"GraphTripleConvNet": modules.GraphTripleConvNet}),
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
    return done
