"""Sampler loops of the two branches, mirroring the reference's sampler surface.

  DDIMSampler.sample(...)            model/networks/diffusion_shape/samplers/ddim.py:60-181
  DiffusionPoint.gen_samples_sg(...) model/networks/diffusion_layout/diffusion_ddpm.py:330-345, 617-620

Each loop iteration is ONE C-ABI call (`echo_shape_step` / `echo_layout_step`): denoiser forward + sampler update,
no host synchronisation, no per-step tensor allocation except the reference-visible noise draw of the DDPM chain
(kept as a `noise_fn` call per step so the RNG stream matches the reference's, diffusion_ddpm.py:302).
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.nn as nn

from ._lib import EchoError
from .modules import UNet1DModel, UNet3DModel


def _find_unet3d(model) -> UNet3DModel:
    if isinstance(model, UNet3DModel):
        return model
    for path in ("df.diffusion_net", "df.module.diffusion_net", "diffusion_net", "df_module.diffusion_net"):
        obj = model
        try:
            for p in path.split("."):
                obj = getattr(obj, p)
        except AttributeError:
            continue
        if isinstance(obj, UNet3DModel):
            return obj
    raise EchoError("DDIMSampler needs a model whose denoiser is an echoscene_b200 UNet3DModel "
                    "(there is no PyTorch fallback on this path)")


class DDIMSampler(object):
    """DDIM sampler over the shape latents (eta = 0, classifier-free guidance disabled as in the reference's
    `elif True:` branch, samplers/ddim.py:207-217: exactly one denoiser forward per step)."""

    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.unet = _find_unet3d(model)
        self.ddpm_num_timesteps = int(getattr(model, "num_timesteps", self.unet.timesteps_total))
        self.schedule = schedule

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        if ddim_discretize != "uniform":
            raise EchoError("only the 'uniform' DDIM discretisation is implemented (the reference's default)")
        if ddim_eta != 0.0:
            raise EchoError("only eta = 0 is on the hot path (echo2shape.py:118, ddim.py:60-125)")
        ls = float(getattr(self.model, "linear_start", self.unet.linear_start))
        le = float(getattr(self.model, "linear_end", self.unet.linear_end))
        self.unet.set_schedule(int(ddim_num_steps), self.ddpm_num_timesteps, ls, le)
        c = self.ddpm_num_timesteps // int(ddim_num_steps)
        self.ddim_timesteps = np.asarray(list(range(0, self.ddpm_num_timesteps, c))) + 1

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, triplet=None,
               **kwargs):
        """Same arguments as the reference.  `unconditional_conditioning` is the `uc_s` (N,1,1280) embedding that the
        reference feeds to the denoiser as obj_embed (ddim.py:216); `conditioning` (c_s) is accepted and unused, as
        in the reference's message-passing configuration (openai_model_3d.py:843-844)."""
        if mask is not None or x0 is not None or score_corrector is not None or quantize_x0:
            raise EchoError("mask / x0 / score_corrector / quantize_x0 are not on the hot path")
        if unconditional_conditioning is None or triplet is None:
            raise EchoError("unconditional_conditioning (uc_s) and triplet are required")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        size = (batch_size,) + tuple(shape)
        device = unconditional_conditioning.device
        img = torch.randn(size, device=device) if x_T is None else x_T
        img = img.float().contiguous().clone()
        nxt = torch.empty_like(img)
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        total = len(self.ddim_timesteps)
        unet = self.unet
        unet._ensure(img.shape[0], triplet.shape[0])
        unet.frozen = True
        try:
            for i in range(total):
                index = total - i - 1
                unet.ddim_step(img, unconditional_conditioning, triplet, index, out=nxt)
                img, nxt = nxt, img
                if callback:
                    callback(i)
                if index % log_every_t == 0 or index == total - 1:
                    intermediates["x_inter"].append(img.clone())
        finally:
            unet.frozen = False
        return img, intermediates


def _find_unet1d(model) -> UNet1DModel:
    if isinstance(model, UNet1DModel):
        return model
    raise EchoError("DiffusionPoint needs an echoscene_b200 UNet1DModel as denoise_net (no PyTorch fallback)")


class _ScheduleView:
    """The GaussianDiffusion attributes other reference code reads (num_timesteps + the fp32 tables)."""

    def __init__(self, tables: torch.Tensor, time_num: int):
        self.num_timesteps = time_num
        (self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1,
         self.posterior_mean_coef2, self.posterior_log_variance_clipped) = tables


class DiffusionPoint(nn.Module):
    """Layout diffusion wrapper — diffusion_ddpm.py:552-620 (same constructor arguments).  Sampling only."""

    def __init__(self, denoise_net, config=None, conditioning_key=None, schedule_type="linear", beta_start=0.0001,
                 beta_end=0.02, time_num=1000, loss_type="mse", model_mean_type="eps", model_var_type="fixedsmall",
                 loss_separate=False, loss_iou=False, iou_type="obb", train_stats_file=None):
        super().__init__()
        if schedule_type != "linear" or model_mean_type != "eps" or model_var_type != "fixedsmall":
            raise EchoError("only schedule 'linear' / mean 'eps' / var 'fixedsmall' are on the hot path "
                            "(config/full_mp.yaml:41-47)")
        self.model = _find_unet1d(denoise_net)
        self.model.set_schedule(int(time_num), float(beta_start), float(beta_end))
        self.time_num = int(time_num)
        self._diffusion = None

    @property
    def diffusion(self):
        if self._diffusion is None:
            self._diffusion = _ScheduleView(self.model.schedule_tables(), self.time_num)
        return self._diffusion

    @torch.no_grad()
    def _denoise(self, data, obj_embed, triples, t, condition_cross=None):
        out = self.model(data, obj_embed, triples, t, context=condition_cross)
        return out.squeeze(-1)

    @torch.no_grad()
    def gen_samples_sg(self, shape, device, obj_embed, triples=None, condition=None,
                       noise_fn: Callable = torch.randn, clip_denoised=True, keep_running=False):
        """p_sample_loop_sg: x_T = noise_fn(shape); for t = T-1..0: x = step(x, t, noise_fn(shape))."""
        if clip_denoised:
            raise EchoError("clip_denoised=True is not on the hot path (generate_layout_sg passes False, "
                            "echo2layout.py:113)")
        assert isinstance(shape, (tuple, list))
        x = noise_fn(size=shape, dtype=torch.float, device=device)
        m = self.model
        m._ensure(shape[0], triples.shape[0])
        m.frozen = True
        try:
            for t in reversed(range(0, self.time_num)):
                noise = noise_fn(size=x.shape, dtype=x.dtype, device=x.device)
                x = m.ddpm_step(x, obj_embed, triples, t, noise)
        finally:
            m.frozen = False
        assert tuple(x.shape) == tuple(shape)
        return x
