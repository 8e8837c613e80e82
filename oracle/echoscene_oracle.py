"""CPU oracle for the EchoScene denoiser hot path — TEST INFRASTRUCTURE ONLY.

A functional restatement (plain torch ops on CPU tensors, fp32 or fp64) of the reference's
per-timestep denoiser forward and of the two sampler updates.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; the product package ``echoscene_b200`` never does (it fails loudly when the CUDA
library is missing instead).

Pinning status: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against *outputs of the reference itself*: ``oracle/gen_golden.py`` imports
the reference modules from /root/reference, loads the same seeded state_dict (strict=True),
checks this restatement against them (max-abs ≤ 1e-5 in fp32, see tests/golden/PINNING.json) and
commits the reference outputs as fixtures under tests/golden/.

All functions take ``sd`` (a state_dict with the reference's key names) and a key prefix.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# graph.py
# --------------------------------------------------------------------------------------


def _linear(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _bn_eval(sd: SD, p: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """BatchNorm1d in eval mode (running statistics), model/layers.py:29-30."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, eps)


def _bn_batch(sd: SD, p: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """BatchNorm1d under model.train(): the statistics of the batch (biased variance), model/layers.py:29-30.  The running
    statistics' update is a side effect the forward value does not depend on; it is restated when the caller asks for it by
    putting a dict under ``sd["__running_update__"]`` (torch/nn/modules/batchnorm.py: momentum 0.1, UNBIASED variance,
    num_batches_tracked += 1)."""
    track = sd.get("__running_update__")
    if track is not None:
        with torch.no_grad():
            m = 0.1
            track[p + ".running_mean"] = (1 - m) * sd[p + ".running_mean"] + m * x.detach().mean(0)
            track[p + ".running_var"] = (1 - m) * sd[p + ".running_var"] + m * x.detach().var(0, unbiased=True)
            if (p + ".num_batches_tracked") in sd:
                track[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, eps)


def gcn_mlp(sd: SD, p: str, x: Tensor, n_layers: int = 2, batch_stats: bool = False) -> Tensor:
    """build_mlp(..., batch_norm='batch', final_nonlinearity=True): (Linear, BN, ReLU) x n.
    model/layers.py:21-38; index layout 0,1,2 / 3,4,5."""
    has_bn = (p + ".1.running_mean") in sd
    step = 3 if has_bn else 2
    for i in range(n_layers):
        x = _linear(sd, f"{p}.{i * step}", x)
        if has_bn:
            x = (_bn_batch if batch_stats else _bn_eval)(sd, f"{p}.{i * step + 1}", x)
        x = F.relu(x)
    return x


def graph_triple_conv(sd: SD, p: str, obj_vecs: Tensor, pred_vecs: Tensor, edges: Tensor, batch_stats: bool = False
                      ) -> Tuple[Tensor, Tensor]:
    """GraphTripleConv.forward with pooling='avg', residual=True.  model/graph.py:124-211."""
    n_obj, n_tri = obj_vecs.shape[0], pred_vecs.shape[0]
    dp = pred_vecs.shape[1]
    s_idx = edges[:, 0].contiguous()
    o_idx = edges[:, 1].contiguous()
    cur_s = obj_vecs[s_idx]                                     # graph.py:146
    cur_o = obj_vecs[o_idx]                                     # graph.py:147
    cur_t = torch.cat([cur_s, pred_vecs, cur_o], dim=1)         # graph.py:151
    new_t = gcn_mlp(sd, p + "net1", cur_t, batch_stats=batch_stats)   # graph.py:152
    hid = (new_t.shape[1] - dp) // 2
    new_s, new_p, new_o = new_t[:, :hid], new_t[:, hid:hid + dp], new_t[:, hid + dp:]   # :156-158
    pooled = torch.zeros(n_obj, hid, dtype=obj_vecs.dtype, device=obj_vecs.device)
    pooled = pooled.index_add(0, s_idx, new_s)                  # scatter_add, graph.py:176
    pooled = pooled.index_add(0, o_idx, new_o)                  # graph.py:177
    counts = torch.zeros(n_obj, dtype=obj_vecs.dtype, device=obj_vecs.device)
    ones = torch.ones(n_tri, dtype=obj_vecs.dtype, device=obj_vecs.device)
    counts = counts.index_add(0, s_idx, ones).index_add(0, o_idx, ones)    # :191-192
    pooled = pooled / counts.clamp(min=1).view(-1, 1)           # :198-199
    new_obj = gcn_mlp(sd, p + "net2", pooled, batch_stats=batch_stats)   # :203
    if (p + "linear_projection.weight") in sd:                  # residual, :205-209
        new_obj = new_obj + _linear(sd, p + "linear_projection", obj_vecs)
        new_p = new_p + _linear(sd, p + "linear_projection_pred", pred_vecs)
    return new_obj, new_p


def graph_triple_conv_net(sd: SD, p: str, obj_vecs: Tensor, pred_vecs: Tensor, edges: Tensor,
                          num_layers: int = 5, batch_stats: bool = False) -> Tuple[Tensor, Tensor]:
    """GraphTripleConvNet.forward.  model/graph.py:246-250.  batch_stats: as under model.train() (BatchNorm1d on batch statistics)."""
    for i in range(num_layers):
        obj_vecs, pred_vecs = graph_triple_conv(sd, f"{p}gconvs.{i}.", obj_vecs, pred_vecs, edges, batch_stats)
    return obj_vecs, pred_vecs


def gather_rows(obj_vecs: Tensor, idx: Tensor) -> Tensor:
    """obj_vecs[idx] — the bit-exact edge-index gather, model/graph.py:146-147."""
    return obj_vecs[idx]


def edges_of(triples: Tensor) -> Tuple[Tensor, Tensor]:
    """triples (T,3) [s,p,o] -> edges (T,2) [s,o], predicate ids.  denoise_net.py:759-761."""
    return torch.stack([triples[:, 0], triples[:, 2]], dim=1), triples[:, 1]


# --------------------------------------------------------------------------------------
# ldm_diffusion_util.py / attention.py building blocks (dims = 1 or 3)
# --------------------------------------------------------------------------------------


def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """[cos(t f) | sin(t f)], f_i = exp(-ln(max_period) i / half).  ldm_diffusion_util.py:174-194."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _conv(sd: SD, p: str, x: Tensor, stride=1, padding=0) -> Tensor:
    w = sd[p + ".weight"]
    fn = F.conv1d if w.dim() == 3 else F.conv3d
    return fn(x, w, sd[p + ".bias"], stride=stride, padding=padding)


def _gn(sd: SD, p: str, x: Tensor, eps: float) -> Tensor:
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def res_block(sd: SD, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward, use_scale_shift_norm=False, no up/down.
    openai_model_3d.py:294-314 / denoise_net.py:293-313; GroupNorm32 eps=1e-5 (ldm_diffusion_util.py:222-239)."""
    h = _conv(sd, p + "in_layers.2", F.silu(_gn(sd, p + "in_layers.0", x, 1e-5)), padding=1)
    e = _linear(sd, p + "emb_layers.1", F.silu(emb))
    while e.dim() < h.dim():
        e = e[..., None]
    h = h + e
    h = _conv(sd, p + "out_layers.3", F.silu(_gn(sd, p + "out_layers.0", h, 1e-5)), padding=1)
    if (p + "skip_connection.weight") in sd:
        x = _conv(sd, p + "skip_connection", x)
    return x + h


def cross_attention(sd: SD, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """CrossAttention.forward.  attention.py:174-219."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _linear(sd, p + ".to_out.0", out)


def transformer_block(sd: SD, p: str, x: Tensor, context: Tensor, heads: int) -> Tensor:
    """BasicTransformerBlock._forward.  attention.py:237-245 (GEGLU: attention.py:39-46)."""
    x = cross_attention(sd, p + "attn1", _ln(sd, p + "norm1", x), None, heads) + x
    x = cross_attention(sd, p + "attn2", _ln(sd, p + "norm2", x), context, heads) + x
    a, g = _linear(sd, p + "ff.net.0.proj", _ln(sd, p + "norm3", x)).chunk(2, dim=-1)
    return _linear(sd, p + "ff.net.2", a * F.gelu(g)) + x


def spatial_transformer(sd: SD, p: str, x: Tensor, context: Tensor, heads: int, depth: int = 1) -> Tensor:
    """SpatialTransformer{1D,3D}.forward; Normalize = GroupNorm(32, eps=1e-6).  attention.py:78-79,334-396."""
    x_in = x
    sp = x.shape[2:]
    h = _conv(sd, p + "proj_in", _gn(sd, p + "norm", x, 1e-6))
    b, c = h.shape[:2]
    h = h.reshape(b, c, -1).permute(0, 2, 1)
    for d in range(depth):
        h = transformer_block(sd, f"{p}transformer_blocks.{d}.", h, context, heads)
    h = h.permute(0, 2, 1).reshape(b, c, *sp)
    return _conv(sd, p + "proj_out", h) + x_in


def _unet_trunk(sd: SD, cfg, dims: int, h: Tensor, emb: Tensor, context: Tensor) -> Tensor:
    """input_blocks -> middle_block -> output_blocks(with skip concat) -> out.
    openai_model_3d.py:849-863 / denoise_net.py:795-806."""
    from echoscene_b200.arch import unet_blocks   # architecture walk only (pure python, no CUDA)
    inp, out, _ = unet_blocks(cfg.model_channels, cfg.channel_mult, cfg.num_res_blocks,
                              cfg.attention_resolutions)
    heads = cfg.num_heads
    hs: List[Tensor] = []
    stride = 2 if dims == 1 else (1, 2, 2)
    for b in inp:
        p = b.name + "."
        if b.kind == "conv_in":
            h = _conv(sd, p + "0", h, padding=1)
        elif b.kind == "res":
            h = res_block(sd, p + "0.", h, emb)
            if b.attn:
                h = spatial_transformer(sd, p + "1.", h, context, heads, cfg.transformer_depth)
        else:  # Downsample with conv: openai_model_3d.py:188-192
            h = _conv(sd, p + "0.op", h, stride=stride, padding=1)
        hs.append(h)
    h = res_block(sd, "middle_block.0.", h, emb)
    h = spatial_transformer(sd, "middle_block.1.", h, context, heads, cfg.transformer_depth)
    h = res_block(sd, "middle_block.2.", h, emb)
    for b in out:
        p = b.name + "."
        h = torch.cat([h, hs.pop()], dim=1)
        h = res_block(sd, p + "0.", h, emb)
        k = 1
        if b.attn:
            h = spatial_transformer(sd, p + "1.", h, context, heads, cfg.transformer_depth)
            k = 2
        if b.up:
            if dims == 3:   # nearest x(1,2,2), openai_model_3d.py:150-153
                h = F.interpolate(h, (h.shape[2], h.shape[3] * 2, h.shape[4] * 2), mode="nearest")
            # dims == 1: scale_factor=1 (denoise_net.py:154) -> identity
            h = _conv(sd, f"{p}{k}.conv", h, padding=1)
    h = F.silu(_gn(sd, "out.0", h, 1e-5))
    return _conv(sd, "out.2", h, padding=1)


# --------------------------------------------------------------------------------------
# layout branch: UNet1DModel + DDPM
# --------------------------------------------------------------------------------------


def box_message_passing(sd: SD, cfg, obj_embed: Tensor, triples: Tensor, box_t: Tensor, emb: Tensor, batch_stats: bool = False) -> Tensor:
    """UNet1DModel.box_messsage_passing.  denoise_net.py:758-771.  batch_stats: as under model.train()."""
    edges, p = edges_of(triples)
    box_embed = _linear(sd, "box_embeddings", box_t)
    pred_embed = sd["pred_embeddings.weight"][p]
    node = torch.cat([obj_embed, box_embed], dim=1)
    if cfg.enable_t_emb:
        node = torch.cat([node, _linear(sd, "box_time_emb", emb)], dim=1)
    out, _ = graph_triple_conv_net(sd, "box_graph_cov.", node, pred_embed, edges, batch_stats=batch_stats)
    return out


def unet1d_forward(sd: SD, cfg, box_t: Tensor, obj_embed: Tensor, triples: Tensor, timesteps: Tensor,
                   context: Optional[Tensor] = None, batch_stats: bool = False) -> Tensor:
    """UNet1DModel.forward (conditioning_key='crossattn'): returns (N, 8, 1).  denoise_net.py:773-806.
    batch_stats: the forward under model.train() -- box_graph_cov's BatchNorm1d layers on the statistics of the batch (GroupNorm /
    LayerNorm do not depend on the mode; dropout is 0; the checkpoint wrapper only changes what autograd stores)."""
    emb = _linear(sd, "time_embed.2", F.silu(_linear(sd, "time_embed.0",
                                                     timestep_embedding(timesteps, cfg.model_channels))))
    latent = box_message_passing(sd, cfg, obj_embed, triples, box_t, emb, batch_stats)
    ctx = latent.unsqueeze(1)                      # overwrites `context`, denoise_net.py:791-792
    h = box_t.unsqueeze(1).permute(0, 2, 1)        # (N, 8, 1)
    return _unet_trunk(sd, cfg, 1, h, emb, ctx)


class DDPMSchedule:
    """GaussianDiffusion tables: float64 numpy -> float32 torch.  diffusion_ddpm.py:38-40,133-162."""

    def __init__(self, beta_start: float = 1e-4, beta_end: float = 0.02, time_num: int = 1000):
        betas = np.linspace(beta_start, beta_end, time_num).astype(np.float64)
        alphas = 1.0 - betas
        ac = torch.from_numpy(np.cumprod(alphas, axis=0)).float()
        ac_prev = torch.from_numpy(np.append(1.0, ac[:-1])).float()
        b32 = torch.from_numpy(betas).float()
        a32 = torch.from_numpy(alphas).float()
        self.num_timesteps = time_num
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / ac).float()
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / ac - 1).float()
        post_var = b32 * (1.0 - ac_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = torch.log(torch.max(post_var, 1e-20 * torch.ones_like(post_var)))
        self.posterior_mean_coef1 = b32 * torch.sqrt(ac_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - ac_prev) * torch.sqrt(a32) / (1.0 - ac)

    def tables(self) -> Tensor:
        """(5, T) fp32: [sqrt_recip, sqrt_recipm1, coef1, coef2, log_var] — the layout the CUDA side takes."""
        return torch.stack([self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod,
                            self.posterior_mean_coef1, self.posterior_mean_coef2,
                            self.posterior_log_variance_clipped]).contiguous()


def ddpm_update(sch: DDPMSchedule, x_t: Tensor, eps: Tensor, t: int, noise: Tensor) -> Tensor:
    """p_mean_variance('eps','fixedsmall', clip_denoised=False) + p_sample_sg.
    diffusion_ddpm.py:220-264, 266-271, 296-309."""
    x0 = sch.sqrt_recip_alphas_cumprod[t] * x_t - sch.sqrt_recipm1_alphas_cumprod[t] * eps
    mean = sch.posterior_mean_coef1[t] * x0 + sch.posterior_mean_coef2[t] * x_t
    nz = 0.0 if t == 0 else 1.0
    return mean + nz * torch.exp(0.5 * sch.posterior_log_variance_clipped[t]) * noise


def layout_chain(sd: SD, cfg, obj_embed: Tensor, triples: Tensor, x_T: Tensor, noises: Sequence[Tensor],
                 time_num: int) -> Tensor:
    """p_sample_loop_sg with injected noise (noises[i] is the draw at loop iteration i).  diffusion_ddpm.py:330-345."""
    sch = DDPMSchedule(time_num=time_num)
    x = x_T
    n = x.shape[0]
    for i, t in enumerate(reversed(range(time_num))):
        ts = torch.full((n,), t, dtype=torch.int64)
        eps = unet1d_forward(sd, cfg, x, obj_embed, triples, ts).squeeze(-1)
        x = ddpm_update(sch, x, eps, t, noises[i])
    return x


# --------------------------------------------------------------------------------------
# shape branch: UNet3DModel + DDIM
# --------------------------------------------------------------------------------------


def shape_embeddings(sd: SD, x: Tensor) -> Tensor:
    """shape_embeddings ModuleList: conv(3->32) maxpool(2,2) conv(32->64) maxpool(k2,s4) flatten linear(512->64).
    openai_model_3d.py:757-764."""
    h = _conv(sd, "shape_embeddings.0", x, padding=1)
    h = F.max_pool3d(h, 2, 2)
    h = _conv(sd, "shape_embeddings.2", h, padding=1)
    h = F.max_pool3d(h, 2, 4)
    return _linear(sd, "shape_embeddings.5", h.flatten(1))


def shape_message_passing(sd: SD, cfg, obj_embed: Tensor, triples: Tensor, x: Tensor, emb: Tensor, batch_stats: bool = False) -> Tensor:
    """UNet3DModel.shape_messsage_passing.  openai_model_3d.py:800-814.  batch_stats: as under model.train()."""
    edges, p = edges_of(triples)
    code = shape_embeddings(sd, x)
    pred_embed = sd["pred_embeddings.weight"][p]
    node = torch.cat([obj_embed.squeeze(1), code], dim=1)
    if cfg.enable_t_emb:
        node = torch.cat([node, _linear(sd, "shape_time_emb", emb)], dim=1)
    out, _ = graph_triple_conv_net(sd, "shape_code_graph_cov.", node, pred_embed, edges, batch_stats=batch_stats)
    return out


def unet3d_forward(sd: SD, cfg, x: Tensor, obj_embed: Tensor, triples: Tensor, timesteps: Tensor,
                   context: Optional[Tensor] = None, batch_stats: bool = False) -> Tensor:
    """UNet3DModel.forward (crossattn, message passing): returns (N, 3, 16, 16, 16).  openai_model_3d.py:816-863.
    batch_stats: the forward under model.train() (shape_code_graph_cov's BatchNorm1d layers on the statistics of the batch)."""
    emb = _linear(sd, "time_embed.2", F.silu(_linear(sd, "time_embed.0",
                                                     timestep_embedding(timesteps, cfg.model_channels))))
    latent = shape_message_passing(sd, cfg, obj_embed, triples, x, emb, batch_stats)
    ctx = latent.unsqueeze(1)                      # "we dont use the previous context", :843-844
    return _unet_trunk(sd, cfg, 3, x, emb, ctx)


class DDIMSchedule:
    """register_schedule + DDIMSampler.make_schedule for eta=0, 'uniform' discretisation.
    echo2shape.py:174-190, ldm_diffusion_util.py:43-47,68-96, samplers/ddim.py:28-57."""

    def __init__(self, S: int = 100, timesteps: int = 1000, linear_start: float = 0.00085,
                 linear_end: float = 0.012):
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
        ac = torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)   # model.alphas_cumprod (fp32)
        c = timesteps // S
        self.ddim_timesteps = np.asarray(list(range(0, timesteps, c))) + 1
        acn = ac.numpy()
        self.alphas = acn[self.ddim_timesteps]
        self.alphas_prev = np.asarray([acn[0]] + acn[self.ddim_timesteps[:-1]].tolist())
        self.sqrt_one_minus_alphas = np.sqrt(1.0 - self.alphas)
        self.sigmas = 0.0 * self.alphas

    def coeffs(self, index: int) -> Tuple[float, float, float, float]:
        """Scalars the way torch.full(..., numpy_scalar) rounds them to fp32 (samplers/ddim.py:246-249)."""
        a_t = torch.full((1,), float(self.alphas[index]))
        a_prev = torch.full((1,), float(self.alphas_prev[index]))
        sigma = torch.full((1,), float(self.sigmas[index]))
        s1m = torch.full((1,), float(self.sqrt_one_minus_alphas[index]))
        return a_t, a_prev, sigma, s1m

    def table(self) -> Tensor:
        """(S, 4) fp32 [sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2)] for the CUDA update."""
        rows = []
        for i in range(len(self.alphas)):
            a_t, a_prev, sigma, s1m = self.coeffs(i)
            rows.append(torch.cat([a_t.sqrt(), s1m, a_prev.sqrt(), (1.0 - a_prev - sigma ** 2).sqrt()]))
        return torch.stack(rows).contiguous()


def ddim_update(sch: DDIMSchedule, x: Tensor, e_t: Tensor, index: int) -> Tuple[Tensor, Tensor]:
    """p_sample_ddim tail with sigma=0.  samplers/ddim.py:246-262."""
    a_t, a_prev, sigma, s1m = sch.coeffs(index)
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


def shape_chain(sd: SD, cfg, uc: Tensor, triples: Tensor, x_T: Tensor, S: int, n_steps: Optional[int] = None) -> Tensor:
    """DDIMSampler.ddim_sampling (first n_steps iterations of an S-step schedule).  samplers/ddim.py:128-181."""
    sch = DDIMSchedule(S)
    x = x_T
    n = x.shape[0]
    total = len(sch.ddim_timesteps)
    for i, step in enumerate(np.flip(sch.ddim_timesteps)):
        if n_steps is not None and i >= n_steps:
            break
        ts = torch.full((n,), int(step), dtype=torch.int64)
        e_t = unet3d_forward(sd, cfg, x, uc, triples, ts)
        x, _ = ddim_update(sch, x, e_t, total - i - 1)
    return x


# --------------------------------------------------------------------------------------
# VQ-VAE decode: the step right after the shape chain, EchoToShape.rel2shape -> decode_no_quant
# (echo2shape.py:522; vqvae_networks/network.py:95-103)
# --------------------------------------------------------------------------------------


def vq_quantize(sd: SD, z: Tensor) -> Tuple[Tensor, Tensor]:
    """VectorQuantizer.forward(z, is_voxel=True) in eval: nearest codebook entry per voxel.  quantizer.py:68-99.
    z (B, C, D, H, W) -> (z_q (B, C, D, H, W), indices (B*D*H*W,)).  The straight-through form z + (z_q - z) is kept:
    it is what the reference returns, rounding included."""
    e = sd["quantize.embedding.weight"]
    zp = z.permute(0, 2, 3, 4, 1).contiguous()                      # b d h w c
    zf = zp.view(-1, e.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", zf, e.t())
    idx = torch.argmin(d, dim=1)
    z_q = F.embedding(idx, e).view(zp.shape)
    z_q = zp + (z_q - zp)
    return z_q.permute(0, 4, 1, 2, 3).contiguous(), idx


def _vq_norm(sd: SD, p: str, x: Tensor) -> Tensor:
    """Normalize(): GroupNorm(32 groups, eps 1e-6, affine) for C in {64, 128, 256}.  vqvae_modules.py:13-22."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)                                      # vqvae_modules.py:9-11


def _conv3(sd: SD, p: str, x: Tensor, pad: int) -> Tensor:
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def vq_resnet_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with temb = None, dropout 0.  vqvae_modules.py:107-127."""
    h = _conv3(sd, p + ".conv1", _swish(_vq_norm(sd, p + ".norm1", x)), 1)
    h = _conv3(sd, p + ".conv2", _swish(_vq_norm(sd, p + ".norm2", h)), 1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv3(sd, p + ".nin_shortcut", x, 0)
    return x + h


def vq_attn_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """AttnBlock.forward: single-head attention over all voxels, scale C^-0.5.  vqvae_modules.py:158-189."""
    h_ = _vq_norm(sd, p + ".norm", x)
    q, k, v = (_conv3(sd, p + "." + n, h_, 0) for n in ("q", "k", "v"))
    b, c = q.shape[:2]
    q = q.reshape(b, c, -1).permute(0, 2, 1)
    k = k.reshape(b, c, -1)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, -1)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(x.shape)
    return x + _conv3(sd, p + ".proj_out", h_, 0)


def vq_decoder(sd: SD, cfg, z: Tensor, p: str = "decoder") -> Tensor:
    """Decoder3D.forward.  vqvae_modules.py:377-409."""
    h = _conv3(sd, p + ".conv_in", z, 1)
    h = vq_resnet_block(sd, p + ".mid.block_1", h)
    h = vq_attn_block(sd, p + ".mid.attn_1", h)
    h = vq_resnet_block(sd, p + ".mid.block_2", h)
    for lvl in reversed(range(len(cfg.ch_mult))):
        for i in range(cfg.num_res_blocks):
            h = vq_resnet_block(sd, f"{p}.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # Upsample, vqvae_modules.py:35-39
            h = _conv3(sd, f"{p}.up.{lvl}.upsample.conv", h, 1)
    h = F.gelu(_vq_norm(sd, p + ".norm_out", h))                    # activ = 'gelu' (exact erf), :300-305,404-407
    return _conv3(sd, p + ".conv_out", h, 1)


def vqvae_decode_no_quant(sd: SD, cfg, h: Tensor) -> Tensor:
    """VQVAE.decode_no_quant(h): quantize -> post_quant_conv -> decoder.  network.py:95-103."""
    quant, _ = vq_quantize(sd, h)
    quant = _conv3(sd, "post_quant_conv", quant, 0)
    return vq_decoder(sd, cfg, quant)


# --------------------------------------------------------------------------------------
# once-per-scene encoders of Sg2ScDiffModel.sample, model/EchoScene.py:143-157, 181-195, 388-413
# --------------------------------------------------------------------------------------


def _scene_embed(sd: SD, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor,
                 pred_table: str = "pred_embeddings_ec.weight"):
    edges, p = edges_of(triples)
    obj_embed = torch.cat([text_feat, F.embedding(objs, sd["obj_embeddings_ec.weight"])], dim=1)       # :149-153
    pred_embed = torch.cat([rel_feat, F.embedding(p, sd[pred_table])], dim=1)
    return edges, obj_embed, pred_embed


def scene_init_encoder(sd: SD, cfg, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor):
    """Sg2ScDiffModel.init_encoder, EchoScene.py:143-157 (use_clip) -> obj_embed, pred_embed, latent_obj, latent_pred."""
    edges, obj_embed, pred_embed = _scene_embed(sd, objs, triples, text_feat, rel_feat)
    latent_obj, latent_pred = graph_triple_conv_net(sd, "gconv_net_ec.", obj_embed, pred_embed, edges, cfg.num_layers)   # :155
    return obj_embed, pred_embed, latent_obj, latent_pred


def scene_manipulate(sd: SD, cfg, latent_f: Tensor, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor):
    """Sg2ScDiffModel.manipulate, EchoScene.py:181-195: latent_f (N, feat + gconv_dim) -> obj_vecs, pred_vecs, obj_embed,
    pred_embed."""
    # the layout-only model looks the predicates up in pred_embeddings_man_dc here (EchoLayout.py:154)
    table = "pred_embeddings_man_dc.weight" if getattr(cfg, "man_dc_preds", False) else "pred_embeddings_ec.weight"
    edges, obj_embed, pred_embed = _scene_embed(sd, objs, triples, text_feat, rel_feat, table)
    obj_vecs = torch.cat([latent_f, obj_embed], dim=1)                                                 # :192
    latent, pred_vecs = graph_triple_conv_net(sd, "gconv_net_manipulation.", obj_vecs, pred_embed, edges, min(cfg.num_layers, 5))
    return latent, pred_vecs, obj_embed, pred_embed


def scene_rel_s(sd: SD, x: Tensor) -> Tensor:
    """rel_s_mlp = make_mlp([640, 960, 1280], batch_norm='batch', norelu=True): Linear, BN, ReLU, Linear.  EchoScene.py:97-100"""
    h = F.relu(_bn_eval(sd, "rel_s_mlp.1", _linear(sd, "rel_s_mlp.0", x)))
    return _linear(sd, "rel_s_mlp.3", h)


def scene_encode(sd: SD, cfg, objs: Tensor, triples: Tensor, text_feat: Tensor, rel_feat: Tensor,
                 change: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """What `sample` computes before the layout and shape chains: init_encoder (embeddings + CLIP -> gconv_net_ec), the zero
    change vector, manipulate (gconv_net_manipulation) and the two rel_s_mlp conditionings.
      obj_embed   (N, 640)    -> layout branch `obj_embed`                     (prepare_boxes(..., obj_embed_, ...))
      latent      (N, 640)    -> layout branch relation condition
      uc_s, c_s   (N, 1, 1280) -> shape branch unconditional / conditional context (EchoScene.py:405-410)"""
    obj_embed, pred_embed, latent_obj, _ = scene_init_encoder(sd, cfg, objs, triples, text_feat, rel_feat)
    if change is None:
        change = torch.zeros(latent_obj.shape[0], cfg.gconv_dim, dtype=latent_obj.dtype, device=latent_obj.device)  # :393-397
    latent, _, obj_embed, pred_embed = scene_manipulate(sd, cfg, torch.cat([latent_obj, change], dim=1), objs, triples, text_feat,
                                                        rel_feat)
    out = {"obj_embed": obj_embed, "pred_embed": pred_embed, "latent": latent, "latent_obj": latent_obj}
    if "rel_s_mlp.0.weight" in sd:   # the layout-only model has no rel_s_mlp (and never samples shapes)
        out["uc_s"] = scene_rel_s(sd, obj_embed).unsqueeze(1)
        out["c_s"] = scene_rel_s(sd, latent).unsqueeze(1)
    return out


def vq_encoder(sd: SD, cfg, x: Tensor, p: str = "encoder") -> Tensor:
    """Encoder3D.forward.  vqvae_modules.py:256-289.  Downsample = zero-pad (0,1) on every spatial axis + Conv3d k3 stride 2
    pad 0 (:42-58)."""
    h = _conv3(sd, p + ".conv_in", x, 1)
    nres = len(cfg.ch_mult)
    for lvl in range(nres):
        for i in range(cfg.num_res_blocks):
            h = vq_resnet_block(sd, f"{p}.down.{lvl}.block.{i}", h)
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1, 0, 1), mode="constant", value=0)
            h = F.conv3d(h, sd[f"{p}.down.{lvl}.downsample.conv.weight"], sd[f"{p}.down.{lvl}.downsample.conv.bias"], stride=2, padding=0)
    h = vq_resnet_block(sd, p + ".mid.block_1", h)
    h = vq_attn_block(sd, p + ".mid.attn_1", h)
    h = vq_resnet_block(sd, p + ".mid.block_2", h)
    h = F.gelu(_vq_norm(sd, p + ".norm_out", h))
    return _conv3(sd, p + ".conv_out", h, 1)


def vqvae_encode_no_quant(sd: SD, cfg, x: Tensor) -> Tensor:
    """VQVAE.encode_no_quant(x): encoder -> quant_conv (no quantisation).  network.py:84-88; x (N, 1, 64, 64, 64) SDF."""
    return _conv3(sd, "quant_conv", vq_encoder(sd, cfg, x), 0)
