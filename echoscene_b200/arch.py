"""Architecture description of the EchoScene denoiser hot path.

This file is the single Python-side statement of *what tensors exist* on the hot
path: for each network it yields an ordered ``{state_dict key: ParamSpec}`` whose
keys and shapes are those of the reference modules, so checkpoints written by the
reference load unchanged (SURVEY.md §5 "Checkpoint / resume").  The C++ side
(``csrc/plan_*.cu``) looks weights up *by these same names*.

Reference constructors this mirrors (names/shapes only, no code shared):
  GraphTripleConv / GraphTripleConvNet   model/graph.py:89-244
  build_mlp                              model/layers.py:21-38
  UNet1DModel                            model/networks/diffusion_layout/denoise_net.py:451-756
  UNet3DModel                            model/networks/diffusion_shape/openai_model_3d.py:452-782
  BasicTransformerBlock/SpatialTransformer{1D,3D}  model/networks/diffusion_shape/attention.py:222-396
  VQVAE (decode path), Decoder3D, VectorQuantizer  model/networks/vqvae_networks/network.py:56-103,
                                                   vqvae_modules.py:61-195,292-409, quantizer.py:10-43
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

# --------------------------------------------------------------------------------------
# configs (config/full_mp.yaml:18-39, config/sdfusion-txt2shape_mp.yaml:16-41)
# --------------------------------------------------------------------------------------


@dataclass
class GCNConfig:
    """GraphTripleConvNet(input_dim_obj, input_dim_pred, num_layers, hidden_dim, output_dim)."""
    input_dim_obj: int
    input_dim_pred: int
    num_layers: int = 5
    hidden_dim: int = 256
    output_dim: int | None = None
    residual: bool = True
    pooling: str = "avg"
    mlp_normalization: str = "batch"


@dataclass
class UNet1DConfig:
    """denoiser_kwargs of config/full_mp.yaml:24-39."""
    in_channels: int = 8
    out_channels: int = 8
    model_channels: int = 512
    channel_mult: Tuple[int, ...] = (1, 1, 1, 1)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (4, 2)
    num_heads: int = 8
    transformer_depth: int = 1
    concat_dim: int = 1280
    crossattn_dim: int = 1280
    using_clip: bool = True
    enable_t_emb: bool = True
    gconv_dim: int = 64

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.model_channels

    @property
    def obj_embed_dim(self) -> int:
        # denoise_net.py:725: gconv_dim*2 + add_dim(512 when using_clip)
        return 2 * self.gconv_dim + (512 if self.using_clip else 0)

    def gcn(self) -> GCNConfig:
        d = self.obj_embed_dim + self.gconv_dim + (self.gconv_dim if self.enable_t_emb else 0)
        return GCNConfig(d, 2 * self.gconv_dim, 5, 4 * self.gconv_dim, self.concat_dim)


@dataclass
class UNet3DConfig:
    """unet.params of config/sdfusion-txt2shape_mp.yaml:16-41 (conditioning_key='crossattn')."""
    in_channels: int = 3
    out_channels: int = 3
    model_channels: int = 224
    channel_mult: Tuple[int, ...] = (1, 2, 3)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (4, 2)
    num_heads: int = 8
    transformer_depth: int = 1
    context_dim: int = 1280
    image_size: int = 16          # latent is (C, 16, 16, 16)
    enable_t_emb: bool = True
    gconv_dim: int = 64

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.model_channels

    def gcn(self) -> GCNConfig:
        d = self.gconv_dim + self.context_dim + (self.gconv_dim if self.enable_t_emb else 0)
        return GCNConfig(d, 2 * self.gconv_dim, 5, 4 * self.gconv_dim, self.context_dim)


# --------------------------------------------------------------------------------------
# parameter specs
# --------------------------------------------------------------------------------------


@dataclass
class ParamSpec:
    shape: Tuple[int, ...]
    init: str            # 'linear_w','linear_b','conv_w','conv_b','kaiming_normal','xavier_normal',
                         # 'ones','zeros','normal','zero_w','zero_b','bn_mean','bn_var','bn_count'
    fan_in: int = 0
    buffer: bool = False  # BatchNorm running stats
    dtype: torch.dtype = torch.float32


Specs = "OrderedDict[str, ParamSpec]"


def _lin(sp: Dict, name: str, cin: int, cout: int, bias: bool = True, w_init: str = "linear_w"):
    sp[name + ".weight"] = ParamSpec((cout, cin), w_init, cin)
    if bias:
        sp[name + ".bias"] = ParamSpec((cout,), "linear_b", cin)


def _conv(sp: Dict, name: str, cin: int, cout: int, k: int, dims: int, w_init: str = "conv_w",
          zero: bool = False):
    ks = (k,) * dims
    fan = cin * k ** dims
    sp[name + ".weight"] = ParamSpec((cout, cin) + ks, "zero_w" if zero else w_init, fan)
    sp[name + ".bias"] = ParamSpec((cout,), "zero_b" if zero else "conv_b", fan)


def _norm(sp: Dict, name: str, c: int):
    sp[name + ".weight"] = ParamSpec((c,), "ones")
    sp[name + ".bias"] = ParamSpec((c,), "zeros")


def _bn(sp: Dict, name: str, c: int):
    _norm(sp, name, c)
    sp[name + ".running_mean"] = ParamSpec((c,), "bn_mean", buffer=True)
    sp[name + ".running_var"] = ParamSpec((c,), "bn_var", buffer=True)
    sp[name + ".num_batches_tracked"] = ParamSpec((), "bn_count", buffer=True, dtype=torch.int64)


def gcn_layer_specs(sp: Dict, prefix: str, din: int, dp: int, hidden: int, dout: int,
                    residual: bool = True, batch_norm: bool = True):
    """One GraphTripleConv (graph.py:89-122): net1 = MLP[2Din+Dp, H, 2H+Dp], net2 = MLP[H, H, Dout]."""
    def mlp(name, dims):
        idx = 0
        for i in range(len(dims) - 1):
            _lin(sp, f"{prefix}{name}.{idx}", dims[i], dims[i + 1], w_init="kaiming_normal")
            idx += 1
            if batch_norm:
                _bn(sp, f"{prefix}{name}.{idx}", dims[i + 1])
                idx += 1
            idx += 1  # ReLU
    mlp("net1", [2 * din + dp, hidden, 2 * hidden + dp])
    mlp("net2", [hidden, hidden, dout])
    if residual:
        _lin(sp, prefix + "linear_projection", din, dout)
        _lin(sp, prefix + "linear_projection_pred", dp, dp)


def gcn_specs(cfg: GCNConfig, prefix: str = "") -> Specs:
    sp: Dict[str, ParamSpec] = OrderedDict()
    for i in range(cfg.num_layers):
        last = cfg.output_dim is not None and i >= cfg.num_layers - 1
        dout = cfg.output_dim if last else cfg.input_dim_obj
        gcn_layer_specs(sp, f"{prefix}gconvs.{i}.", cfg.input_dim_obj, cfg.input_dim_pred,
                        cfg.hidden_dim, dout, cfg.residual, cfg.mlp_normalization == "batch")
    return sp


def _resblock(sp, p, cin, cout, emb, dims):
    _norm(sp, p + "in_layers.0", cin)
    _conv(sp, p + "in_layers.2", cin, cout, 3, dims)
    _lin(sp, p + "emb_layers.1", emb, cout)
    _norm(sp, p + "out_layers.0", cout)
    _conv(sp, p + "out_layers.3", cout, cout, 3, dims, zero=True)
    if cin != cout:
        _conv(sp, p + "skip_connection", cin, cout, 1, dims)


def _transformer(sp, p, ch, heads, ctx_dim, dims, depth):
    inner = heads * (ch // heads)
    xav = "xavier_normal" if dims == 3 else "conv_w"   # attention.py:294-296,331 (3-D only)
    _norm(sp, p + "norm", ch)
    _conv(sp, p + "proj_in", ch, inner, 1, dims, w_init=xav)
    for d in range(depth):
        b = f"{p}transformer_blocks.{d}."
        for attn, cd in (("attn1", inner), ("attn2", ctx_dim)):
            _lin(sp, b + attn + ".to_q", inner, inner, bias=False)
            _lin(sp, b + attn + ".to_k", cd, inner, bias=False)
            _lin(sp, b + attn + ".to_v", cd, inner, bias=False)
            _lin(sp, b + attn + ".to_out.0", inner, inner)
            if attn == "attn1":
                _lin(sp, b + "ff.net.0.proj", inner, inner * 8)
                _lin(sp, b + "ff.net.2", inner * 4, inner)
        for n in ("norm1", "norm2", "norm3"):
            _norm(sp, b + n, inner)
    # proj_out is zero_module'd, then (3-D only) re-initialised by self.apply(init_weights):
    # weight xavier-normal, bias stays zero (attention.py:324-331).
    if dims == 3:
        sp[p + "proj_out.weight"] = ParamSpec((ch, inner, 1, 1, 1), "xavier_normal", inner)
        sp[p + "proj_out.bias"] = ParamSpec((ch,), "zero_b", inner)
    else:
        _conv(sp, p + "proj_out", inner, ch, 1, dims, zero=True)


@dataclass
class BlockDesc:
    """One TimestepEmbedSequential of the UNet, in execution order."""
    name: str                 # e.g. 'input_blocks.4'
    kind: str                 # 'conv_in' | 'res' | 'down' | 'mid'
    cin: int = 0
    cout: int = 0
    attn: bool = False
    up: bool = False
    ds: int = 1               # downsample factor at which the block runs


def unet_blocks(model_channels: int, channel_mult: Sequence[int], num_res_blocks: int,
                attention_resolutions: Sequence[int]) -> Tuple[List[BlockDesc], List[BlockDesc], int]:
    """Walk the openai-UNet constructor order (openai_model_3d.py:563-728 / denoise_net.py:553-713)."""
    inp: List[BlockDesc] = [BlockDesc("input_blocks.0", "conv_in", 0, model_channels)]
    chans = [model_channels]
    ch, ds = model_channels, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            co = mult * model_channels
            inp.append(BlockDesc(f"input_blocks.{len(inp)}", "res", ch, co, ds in attention_resolutions, ds=ds))
            ch = co
            chans.append(ch)
        if level != len(channel_mult) - 1:
            inp.append(BlockDesc(f"input_blocks.{len(inp)}", "down", ch, ch, ds=ds))
            chans.append(ch)
            ds *= 2
    mid_ch = ch
    out: List[BlockDesc] = []
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            co = model_channels * mult
            b = BlockDesc(f"output_blocks.{len(out)}", "res", ch + ich, co, ds in attention_resolutions, ds=ds)
            ch = co
            if level and i == num_res_blocks:
                b.up = True
                ds //= 2
            out.append(b)
    return inp, out, mid_ch


def _unet_trunk_specs(sp, cfg, dims: int, ctx_dim: int):
    emb = cfg.time_embed_dim
    _lin(sp, "time_embed.0", cfg.model_channels, emb)
    _lin(sp, "time_embed.2", emb, emb)
    inp, out, mid_ch = unet_blocks(cfg.model_channels, cfg.channel_mult, cfg.num_res_blocks,
                                   cfg.attention_resolutions)
    for b in inp:
        p = b.name + "."
        if b.kind == "conv_in":
            _conv(sp, p + "0", cfg.in_channels, cfg.model_channels, 3, dims)
        elif b.kind == "res":
            _resblock(sp, p + "0.", b.cin, b.cout, emb, dims)
            if b.attn:
                _transformer(sp, p + "1.", b.cout, cfg.num_heads, ctx_dim, dims, cfg.transformer_depth)
        else:
            _conv(sp, p + "0.op", b.cin, b.cout, 3, dims)
    _resblock(sp, "middle_block.0.", mid_ch, mid_ch, emb, dims)
    _transformer(sp, "middle_block.1.", mid_ch, cfg.num_heads, ctx_dim, dims, cfg.transformer_depth)
    _resblock(sp, "middle_block.2.", mid_ch, mid_ch, emb, dims)
    for b in out:
        p = b.name + "."
        _resblock(sp, p + "0.", b.cin, b.cout, emb, dims)
        k = 1
        if b.attn:
            _transformer(sp, p + "1.", b.cout, cfg.num_heads, ctx_dim, dims, cfg.transformer_depth)
            k = 2
        if b.up:
            _conv(sp, f"{p}{k}.conv", b.cout, b.cout, 3, dims)
    _norm(sp, "out.0", cfg.model_channels)
    _conv(sp, "out.2", cfg.model_channels, cfg.out_channels, 3, dims, zero=True)


def unet1d_specs(cfg: UNet1DConfig) -> Specs:
    """UNet1DModel state_dict (denoise_net.py:451-756)."""
    sp: Dict[str, ParamSpec] = OrderedDict()
    _unet_trunk_specs(sp, cfg, 1, cfg.crossattn_dim)
    sp["pred_embeddings.weight"] = ParamSpec((16, 2 * cfg.gconv_dim), "normal")
    _lin(sp, "box_embeddings", cfg.in_channels, cfg.gconv_dim, w_init="kaiming_normal")
    if cfg.enable_t_emb:
        _lin(sp, "box_time_emb", cfg.time_embed_dim, cfg.gconv_dim)
    sp.update(gcn_specs(cfg.gcn(), "box_graph_cov."))
    return sp


def unet3d_specs(cfg: UNet3DConfig) -> Specs:
    """UNet3DModel state_dict (openai_model_3d.py:452-782), crossattn + message passing."""
    sp: Dict[str, ParamSpec] = OrderedDict()
    _unet_trunk_specs(sp, cfg, 3, cfg.context_dim)
    g = cfg.gconv_dim
    sp["pred_embeddings.weight"] = ParamSpec((16, 2 * g), "normal")
    _conv(sp, "shape_embeddings.0", cfg.in_channels, 32, 3, 3)
    _conv(sp, "shape_embeddings.2", 32, 64, 3, 3)
    _lin(sp, "shape_embeddings.5", 64 * 2 * 2 * 2, g)
    if cfg.enable_t_emb:
        _lin(sp, "shape_time_emb", cfg.time_embed_dim, g)
    sp.update(gcn_specs(cfg.gcn(), "shape_code_graph_cov."))
    return sp


# --------------------------------------------------------------------------------------
# initialisation
# --------------------------------------------------------------------------------------


@dataclass
class SceneEncoderConfig:
    """The once-per-scene encoders of Sg2ScDiffModel (model/EchoScene.py:15-128, as built by SGDiff.py:21-22:
    embedding_dim = 64, mlp_normalization = 'batch', gconv_num_layers = 5, CLIP features on)."""
    gconv_dim: int = 64
    add_dim: int = 512            # CLIP ViT-B/32 feature width (use_clip)
    num_objs: int = 36            # len(vocab['object_idx_to_name']) incl. '_scene_'
    num_preds: int = 16
    num_layers: int = 5
    residual: bool = True
    rel_s_hidden: int = 960
    context_dim: int = 1280
    # the layout-only model embeds the predicates of `manipulate` with pred_embeddings_man_dc (EchoLayout.py:154);
    # Sg2ScDiffModel uses pred_embeddings_ec there too (EchoScene.py:187)
    man_dc_preds: bool = False

    @property
    def feat_dim(self) -> int:    # out_dim_ini_encoder == out_dim_manipulator
        return 2 * self.gconv_dim + self.add_dim

    def gcn_ec(self) -> GCNConfig:
        return GCNConfig(self.feat_dim, self.feat_dim, self.num_layers, 4 * self.gconv_dim, self.feat_dim, self.residual)

    def gcn_manipulation(self) -> GCNConfig:
        din = self.feat_dim + self.gconv_dim + self.feat_dim   # latent_f | change flag | obj embedding + CLIP
        return GCNConfig(din, self.feat_dim, min(self.num_layers, 5), 4 * self.gconv_dim, self.feat_dim, self.residual)


def scene_encoder_specs(cfg: SceneEncoderConfig) -> Specs:
    """The tensors Sg2ScDiffModel.sample touches before the two chains start (EchoScene.py:143-157, 181-195, 388-413):
    embeddings, gconv_net_ec, gconv_net_manipulation, rel_s_mlp (make_mlp([640, 960, 1280], 'batch', norelu=True))."""
    sp: Dict[str, ParamSpec] = OrderedDict()
    sp["obj_embeddings_ec.weight"] = ParamSpec((cfg.num_objs + 1, 2 * cfg.gconv_dim), "normal")
    sp["pred_embeddings_ec.weight"] = ParamSpec((cfg.num_preds, 2 * cfg.gconv_dim), "normal")
    if cfg.man_dc_preds:
        sp["pred_embeddings_man_dc.weight"] = ParamSpec((cfg.num_preds, 2 * cfg.gconv_dim), "normal")
    sp.update(gcn_specs(cfg.gcn_ec(), "gconv_net_ec."))
    sp.update(gcn_specs(cfg.gcn_manipulation(), "gconv_net_manipulation."))
    _lin(sp, "rel_s_mlp.0", cfg.feat_dim, cfg.rel_s_hidden, w_init="kaiming_normal")
    _bn(sp, "rel_s_mlp.1", cfg.rel_s_hidden)
    _lin(sp, "rel_s_mlp.3", cfg.rel_s_hidden, cfg.context_dim, w_init="kaiming_normal")
    return sp


@dataclass
class VQVAEConfig:
    """model.params of config/vqvae_snet.yaml:5-19 (the decode half)."""
    embed_dim: int = 3
    n_embed: int = 8192
    z_channels: int = 3
    resolution: int = 64
    out_ch: int = 1
    ch: int = 64
    ch_mult: Tuple[int, ...] = (1, 2, 4)
    num_res_blocks: int = 1

    @property
    def latent_size(self) -> int:
        return self.resolution // 2 ** (len(self.ch_mult) - 1)


def _vq_resnet(sp: Dict, name: str, cin: int, cout: int):
    """ResnetBlock with temb_channels = 0 (vqvae_modules.py:61-127): norm1, conv1, norm2, conv2 (+ 1x1 nin_shortcut)."""
    _norm(sp, name + ".norm1", cin)
    _conv(sp, name + ".conv1", cin, cout, 3, 3)
    _norm(sp, name + ".norm2", cout)
    _conv(sp, name + ".conv2", cout, cout, 3, 3)
    if cin != cout:
        _conv(sp, name + ".nin_shortcut", cin, cout, 1, 3)


def vqvae_decode_specs(cfg: VQVAEConfig) -> Specs:
    """The part of the VQVAE state_dict that `VQVAE.decode_no_quant` touches (network.py:95-103): codebook,
    post_quant_conv, Decoder3D (vqvae_modules.py:292-409; attn_resolutions = [] -> only mid.attn_1)."""
    sp: Dict[str, ParamSpec] = OrderedDict()
    sp["quantize.embedding.weight"] = ParamSpec((cfg.n_embed, cfg.embed_dim), "normal")
    _conv(sp, "post_quant_conv", cfg.embed_dim, cfg.z_channels, 1, 3)
    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[-1]
    _conv(sp, "decoder.conv_in", cfg.z_channels, block_in, 3, 3)
    _vq_resnet(sp, "decoder.mid.block_1", block_in, block_in)
    _norm(sp, "decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(sp, "decoder.mid.attn_1." + n, block_in, block_in, 1, 3)
    _vq_resnet(sp, "decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for i in range(cfg.num_res_blocks):
            _vq_resnet(sp, f"decoder.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            _conv(sp, f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3, 3)
    _norm(sp, "decoder.norm_out", block_in)
    _conv(sp, "decoder.conv_out", block_in, cfg.out_ch, 3, 3)
    return sp


def vqvae_encode_specs(cfg: VQVAEConfig, in_channels: int = 1) -> Specs:
    """The part of the VQVAE state_dict that `VQVAE.encode_no_quant` touches (network.py:84-88): Encoder3D
    (vqvae_modules.py:178-289; attn_resolutions = [] -> only mid.attn_1) and quant_conv."""
    sp: Dict[str, ParamSpec] = OrderedDict()
    nres = len(cfg.ch_mult)
    in_mult = (1,) + tuple(cfg.ch_mult)
    _conv(sp, "encoder.conv_in", in_channels, cfg.ch, 3, 3)
    block_in = cfg.ch
    for lvl in range(nres):
        block_in = cfg.ch * in_mult[lvl]
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for i in range(cfg.num_res_blocks):
            _vq_resnet(sp, f"encoder.down.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != nres - 1:
            _conv(sp, f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3, 3)
    _vq_resnet(sp, "encoder.mid.block_1", block_in, block_in)
    _norm(sp, "encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(sp, "encoder.mid.attn_1." + n, block_in, block_in, 1, 3)
    _vq_resnet(sp, "encoder.mid.block_2", block_in, block_in)
    _norm(sp, "encoder.norm_out", block_in)
    _conv(sp, "encoder.conv_out", block_in, cfg.z_channels, 3, 3)
    _conv(sp, "quant_conv", cfg.z_channels, cfg.embed_dim, 1, 3)
    return sp


def init_tensor(spec: ParamSpec, gen: torch.Generator, rerandomize_zero: bool = False,
                randomize_bn: bool = False) -> torch.Tensor:
    """Draw one tensor following the reference's initialisers.

    ``rerandomize_zero`` replaces ``zero_module`` tensors by N(0, 0.02) and ``randomize_bn`` draws
    non-trivial BatchNorm running statistics / affine terms: both are required for a non-vacuous
    parity test on untrained weights.
    """
    shp = spec.shape
    k = spec.init
    if k in ("linear_w", "conv_w"):
        bound = 1.0 / math.sqrt(spec.fan_in)       # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
        return (torch.rand(shp, generator=gen) * 2 - 1) * bound
    if k in ("linear_b", "conv_b"):
        bound = 1.0 / math.sqrt(spec.fan_in)
        return (torch.rand(shp, generator=gen) * 2 - 1) * bound
    if k == "kaiming_normal":
        return torch.randn(shp, generator=gen) * math.sqrt(2.0 / spec.fan_in)
    if k == "xavier_normal":
        rf = 1
        for s in shp[2:]:
            rf *= s
        return torch.randn(shp, generator=gen) * math.sqrt(2.0 / (shp[0] * rf + shp[1] * rf))
    if k == "normal":
        return torch.randn(shp, generator=gen)
    if k in ("zero_w", "zero_b"):
        if rerandomize_zero:
            return torch.randn(shp, generator=gen) * 0.02
        return torch.zeros(shp)
    if k == "ones":
        if randomize_bn:
            return 1.0 + 0.1 * torch.randn(shp, generator=gen)
        return torch.ones(shp)
    if k == "zeros":
        if randomize_bn:
            return 0.1 * torch.randn(shp, generator=gen)
        return torch.zeros(shp)
    if k == "bn_mean":
        return 0.1 * torch.randn(shp, generator=gen) if randomize_bn else torch.zeros(shp)
    if k == "bn_var":
        return 0.5 + torch.rand(shp, generator=gen) if randomize_bn else torch.ones(shp)
    if k == "bn_count":
        return torch.zeros(shp, dtype=torch.int64)
    raise ValueError(k)


def make_state_dict(specs: Specs, seed: int, parity: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Seeded state_dict. ``parity=True`` ⇒ zero-init tensors re-drawn, BN stats/affines randomised."""
    gen = torch.Generator().manual_seed(seed)
    return OrderedDict((k, init_tensor(s, gen, parity, parity)) for k, s in specs.items())


def count_params(specs: Specs) -> int:
    n = 0
    for s in specs.values():
        if not s.buffer:
            m = 1
            for d in s.shape:
                m *= d
            n += m
    return n
