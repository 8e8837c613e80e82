"""Import the reference's hot-path modules from /root/reference with stubs for absent optional deps.

TEST INFRASTRUCTURE (used only by oracle/gen_golden.py in the build container; /root/reference does
not exist on the GPU box).  Nothing is copied: the reference is imported in place.
Recipe: SURVEY.md §8(c).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("ECHOSCENE_REF", "/root/reference")

_MOCKS = [
    "termcolor", "mcubes", "pytorch3d", "pytorch3d.io", "pytorch3d.structures", "pytorch3d.renderer",
    "pytorch3d.transforms", "pytorch3d.ops", "pytorch3d.loss", "pytorch3d.utils", "pytorch3d.renderer.cameras",
    "fvcore", "fvcore.common", "fvcore.common.param_scheduler", "trimesh", "h5py", "clip",
    "tensorboardX", "open3d", "skimage", "skimage.measure", "imageio", "matplotlib", "matplotlib.pyplot",
    "kornia", "seaborn", "pyrender",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def install_stubs():
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        class OmegaConf:  # minimal: the hot-path modules only touch ListConfig
            @staticmethod
            def load(path):
                import yaml
                with open(path) as f:
                    return yaml.safe_load(f)

        lc.ListConfig = ListConfig
        om.OmegaConf = OmegaConf
        om.listconfig = lc
        om.ListConfig = ListConfig
        sys.modules["omegaconf"] = om
        sys.modules["omegaconf.listconfig"] = lc
    for name in _MOCKS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()


def load():
    """Returns a namespace with the reference classes the oracle is pinned against."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    g = importlib.import_module("model.graph")
    ns.GraphTripleConv, ns.GraphTripleConvNet = g.GraphTripleConv, g.GraphTripleConvNet
    d = importlib.import_module("model.networks.diffusion_layout.denoise_net")
    ns.UNet1DModel = d.UNet1DModel
    dd = importlib.import_module("model.networks.diffusion_layout.diffusion_ddpm")
    ns.DiffusionPoint = dd.DiffusionPoint
    o = importlib.import_module("model.networks.diffusion_shape.openai_model_3d")
    ns.UNet3DModel = o.UNet3DModel
    u = importlib.import_module("model.networks.diffusion_shape.ldm_diffusion_util")
    ns.ldm_util = u
    s = importlib.import_module("model.networks.diffusion_shape.samplers.ddim")
    ns.DDIMSampler = s.DDIMSampler
    v = importlib.import_module("model.networks.vqvae_networks.network")
    ns.VQVAE = v.VQVAE
    return ns
