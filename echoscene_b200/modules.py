"""Host-side mirror of the reference's operator surface for the denoiser hot path.

Each class keeps the reference constructor signature, forward signature and ``state_dict`` key names (so the
reference's checkpoints load with ``strict=True``), holds its parameters as ordinary ``nn.Parameter``s, and runs its
forward as CUDA kernels of ``libechoscene_b200.so`` through the C ABI (``_lib``).  There is no PyTorch arithmetic
and no CPU path in any forward: weights are repacked by the library when the handle is (re)built, which happens
lazily on the first call and whenever a parameter was modified or the graph outgrew the handle's capacity.

  GraphTripleConv / GraphTripleConvNet   model/graph.py:89-250
  UNet1DModel                            model/networks/diffusion_layout/denoise_net.py:451-806
  UNet3DModel, DiffusionUNet             model/networks/diffusion_shape/openai_model_3d.py:452-863, network.py:11-43
  SceneEncoder                           the encoder sub-modules / methods of model/EchoScene.py:46-100, 143-195
  VQVAE                                  model/networks/vqvae_networks/network.py:56-103 (decode 8f-1, encode_no_quant 8f-3)

Inference only (eval-mode BatchNorm, no autograd): the training backward is outside this round's scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, arch
from ._lib import EchoError


def _next_pow2(n: int, lo: int) -> int:
    c = lo
    while c < n:
        c *= 2
    return c


class _SpecModule(nn.Module):
    """nn.Module whose (nested) parameters/buffers are created from an ``arch`` spec table, so state_dict keys and
    shapes are the reference's."""

    def _build_from_specs(self, specs: "arch.Specs"):
        self._spec_keys = list(specs.keys())
        # load_state_dict(assign=True) swaps Parameter objects without touching their versions: drop the cached list so
        # that the next call sees the new storage (the hook also fires when a parent module loads a checkpoint)
        self._register_load_state_dict_pre_hook(self._forget_weights)
        for key, spec in specs.items():
            parts = key.split(".")
            mod: nn.Module = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, nn.Module())
                mod = mod._modules[p]
            t = arch.init_tensor(spec, _SpecModule._gen)
            if spec.buffer:
                mod.register_buffer(parts[-1], t)
            else:
                mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))

    _gen = torch.Generator().manual_seed(0)

    # ---- handle management ----
    _handle = None
    _handle_key = None

    _wlist = None
    frozen = False   # samplers set this inside a chain: weights cannot change between two steps

    def _weights_version(self):
        if self.frozen and self._handle_key is not None:
            return self._handle_key[0]
        if self._wlist is None or len(self._wlist) != len(self._spec_keys):
            self._wlist = list(self.state_dict(keep_vars=True).values())
        return tuple((t.data_ptr(), t._version) for t in self._wlist)

    def _forget_weights(self, *a, **k):
        self._wlist = None

    def _apply(self, fn, *a, **k):   # .cuda()/.to() replace parameter storage
        self._wlist = None
        return super()._apply(fn, *a, **k)

    def _check_eval(self):
        if self.training:
            raise EchoError(f"{type(self).__name__}: this call is a sampler step and runs under eval() only "
                            "(the reference samples under model.eval(), scripts/eval_3dfront.py:395)")

    # train(): forward computes VALUES on batch statistics, without an autograd tape.  integrate.patch_reference() turns this off on the
    # classes it hands to the reference: inside the reference's own training loop a loss without a tape must not appear at all
    train_forward_values = True

    def _check_train_values(self):
        if self.training and not self.train_forward_values:
            raise EchoError(f"{type(self).__name__}: train() mode inside the patched reference -- the B200 path computes forward values "
                            "only (no autograd tape, no backward pass for this module yet); keep the reference's own module for "
                            "training, or call .eval() for sampling (scripts/eval_3dfront.py:395)")

    _set_batch_stats_fn = None   # name of the library's echo_*_set_batch_stats for this handle type

    def _apply_mode(self):
        """model.train(): the BatchNorm1d layers of the GCN MLPs (and rel_s_mlp) normalise with the statistics of the batch
        (model/layers.py:29-30) -- forward VALUES of the reference's training forward; nothing is recorded for a backward pass and
        running statistics are not updated.  Called after _ensure()."""
        fn = getattr(_lib.lib(), self._set_batch_stats_fn)
        _lib.check(fn(self._handle, int(bool(self.training))))

    def _destroy_handle(self):
        raise NotImplementedError

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------------------------
# graph.py
# ----------------------------------------------------------------------------------------------------------------------
class GraphTripleConvNet(_SpecModule):
    """A sequence of scene graph convolution layers — model/graph.py:214-250 (same arguments)."""

    def __init__(self, input_dim_obj, input_dim_pred, num_layers=2, hidden_dim=512, residual=False, pooling="avg",
                 mlp_normalization="none", output_dim=None):
        super().__init__()
        if pooling != "avg":
            raise EchoError(f"pooling='{pooling}' is not on the hot path (every denoiser construction passes 'avg', "
                            "denoise_net.py:729, openai_model_3d.py:770)")
        if mlp_normalization not in ("batch", "none"):
            raise EchoError(f"unsupported mlp_normalization {mlp_normalization}")
        self.num_layers = num_layers
        self.cfg = arch.GCNConfig(input_dim_obj, input_dim_pred, num_layers, hidden_dim, output_dim, residual,
                                  pooling, mlp_normalization)
        self._build_from_specs(arch.gcn_specs(self.cfg))
        self.eval()

    def _ensure(self, n_nodes: int, n_triples: int, train_weights: bool = False):
        ver = self._weights_version()
        cap = self._handle_key[1] if self._handle_key else (0, 0, False)
        if (self._handle is not None and self._handle_key[0] == ver and n_nodes <= cap[0] and n_triples <= cap[1]
                and (cap[2] or not train_weights)):
            return
        self._destroy_handle()
        cap = (_next_pow2(n_nodes, 32), _next_pow2(n_triples, 128), bool(train_weights or cap[2]))
        d = _lib.GcnDesc(self.cfg.input_dim_obj, self.cfg.input_dim_pred, self.cfg.num_layers, self.cfg.hidden_dim,
                         self.cfg.output_dim or 0, cap[0], cap[1], 1e-5, int(cap[2]))
        arr, n, keep = _lib.weights_table(self.state_dict_for_lib())
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_gcn_create(C.byref(h), C.byref(d), arr, n))
        self._handle, self._handle_key = h, (ver, cap)

    def state_dict_for_lib(self):
        """the state_dict under the key names the library expects (gconvs.<i>.*); the single-layer subclass re-prefixes its keys"""
        return self.state_dict()

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().echo_gcn_destroy(self._handle)
            self._handle = None

    @torch.no_grad()
    def forward(self, obj_vecs, pred_vecs, edges):
        """eval(): running statistics (folded into the Linears).  train(): ``forward_batch_stats``."""
        if self.training:
            self._check_train_values()
            return self.forward_batch_stats(obj_vecs, pred_vecs, edges)
        return self._run(obj_vecs, pred_vecs, edges, batch_stats=False)

    @torch.no_grad()
    def forward_batch_stats(self, obj_vecs, pred_vecs, edges):
        """The forward the reference computes under ``model.train()`` (scripts/train_3dfront.py:237): every BatchNorm1d of the MLPs
        normalises with the statistics of the batch (model/layers.py:29-30) -- triples for net1, nodes for net2.  Forward values only:
        no autograd tape is recorded and the running statistics are not updated."""
        if self.cfg.mlp_normalization != "batch":
            return self._run(obj_vecs, pred_vecs, edges, batch_stats=False)
        return self._run(obj_vecs, pred_vecs, edges, batch_stats=True)

    def _run(self, obj_vecs, pred_vecs, edges, batch_stats: bool):
        _lib.require_cuda(obj_vecs, pred_vecs, edges)
        obj_vecs = obj_vecs.float().contiguous()
        pred_vecs = pred_vecs.float().contiguous()
        n, t = obj_vecs.shape[0], pred_vecs.shape[0]
        assert obj_vecs.shape[1] == self.cfg.input_dim_obj and pred_vecs.shape[1] == self.cfg.input_dim_pred
        assert edges.shape == (t, 2)
        self._ensure(n, t, train_weights=batch_stats)
        g = _lib.graph_for_edges(edges, n) if not hasattr(edges, "_echo_graph") else edges._echo_graph
        dout = self.cfg.output_dim or self.cfg.input_dim_obj
        obj_out = torch.empty(n, dout, device=obj_vecs.device)
        pred_out = torch.empty(t, self.cfg.input_dim_pred, device=obj_vecs.device)
        fn = _lib.lib().echo_gcn_forward_train if batch_stats else _lib.lib().echo_gcn_forward
        _lib.check(fn(self._handle, g.h, _lib.ptr(obj_vecs), _lib.ptr(pred_vecs), _lib.ptr(obj_out), _lib.ptr(pred_out), _lib.stream_ptr()))
        return obj_out, pred_out


class GraphTripleConv(GraphTripleConvNet):
    """A single layer of scene graph convolution — model/graph.py:89-211 (same arguments)."""

    def __init__(self, input_dim_obj, input_dim_pred, output_dim=None, hidden_dim=512, pooling="avg",
                 mlp_normalization="none", residual=True):
        nn.Module.__init__(self)
        if pooling != "avg":
            raise EchoError(f"pooling='{pooling}' is not on the hot path")
        self.num_layers = 1
        self.input_dim_obj, self.input_dim_pred = input_dim_obj, input_dim_pred
        self.output_dim = output_dim or input_dim_obj
        self.hidden_dim, self.residual, self.pooling = hidden_dim, residual, pooling
        self.cfg = arch.GCNConfig(input_dim_obj, input_dim_pred, 1, hidden_dim, self.output_dim, residual, pooling,
                                  mlp_normalization)
        specs = arch.OrderedDict()
        arch.gcn_layer_specs(specs, "", input_dim_obj, input_dim_pred, hidden_dim, self.output_dim, residual,
                             mlp_normalization == "batch")
        self._build_from_specs(specs)
        self.eval()

    def state_dict_for_lib(self):
        return {"gconvs.0." + k: v for k, v in self.state_dict().items()}



# ----------------------------------------------------------------------------------------------------------------------
# once-per-scene encoders
# ----------------------------------------------------------------------------------------------------------------------
class SceneEncoder(_SpecModule):
    """The encoder sub-modules of ``Sg2ScDiffModel`` and the methods of it that run before the two chains start
    (model/EchoScene.py): ``obj_embeddings_ec``, ``pred_embeddings_ec``, ``gconv_net_ec``, ``gconv_net_manipulation``,
    ``rel_s_mlp`` -- same state_dict keys, so the matching slice of a reference checkpoint loads with strict=True
    (``load_reference_state_dict`` drops the keys of other sub-modules).

    (``man_dc_preds=True``, ``with_rel_s=False``: the layout-only ``Sg2BoxDiffModel``, model/EchoLayout.py, whose ``manipulate``
    embeds predicates with ``pred_embeddings_man_dc`` and which has no ``rel_s_mlp``.)

      init_encoder(objs, triples, text_feat, rel_feat)            EchoScene.py:143-157
      manipulate(latent_f, objs, triples, text_feat, rel_feat)    EchoScene.py:181-195
      rel_s(x)   [the reference's self.rel_s_mlp(x)]              EchoScene.py:97-100
      encode(objs, triples, text_feat, rel_feat)                  the encoder stage of sample(), EchoScene.py:388-410

    Each method is ONE asynchronous C-ABI call (echo_scene_*); the reference's per-node host->device loop (:393-397)
    does not exist here."""

    PREFIXES = ("obj_embeddings_ec.", "pred_embeddings_ec.", "gconv_net_ec.", "gconv_net_manipulation.", "rel_s_mlp.")
    # + "pred_embeddings_man_dc." when man_dc_preds: the layout-only model's `manipulate` looks predicates up there
    # (EchoLayout.py:154); Sg2ScDiffModel owns that table too but never reads it when sampling

    def __init__(self, num_objs: int = 36, num_preds: int = 16, embedding_dim: int = 64, gconv_num_layers: int = 5,
                 residual: bool = True, use_clip: bool = True, gconv_pooling: str = "avg",
                 mlp_normalization: str = "batch", rel_s_hidden: int = 960, context_dim: int = 1280,
                 with_rel_s: bool = True, man_dc_preds: bool = False):
        super().__init__()
        if gconv_pooling != "avg":
            raise EchoError(f"gconv_pooling='{gconv_pooling}' is not on the hot path (SGDiff.py:21-22 passes 'avg')")
        if mlp_normalization != "batch":
            raise EchoError("SceneEncoder mirrors the SGDiff construction (mlp_normalization='batch', SGDiff.py:21-22)")
        self.cfg = arch.SceneEncoderConfig(gconv_dim=embedding_dim, add_dim=512 if use_clip else 0, num_objs=num_objs,
                                           num_preds=num_preds, num_layers=gconv_num_layers, residual=residual,
                                           rel_s_hidden=rel_s_hidden, context_dim=context_dim, man_dc_preds=bool(man_dc_preds))
        self.embedding_dim = embedding_dim
        self.clip = use_clip
        self.out_dim_ini_encoder = self.out_dim_manipulator = self.cfg.feat_dim
        specs = arch.scene_encoder_specs(self.cfg)
        self.with_rel_s = bool(with_rel_s)
        if not self.with_rel_s:   # the layout-only model (model/EchoLayout.py) has no rel_s_mlp
            specs = arch.OrderedDict((k, v) for k, v in specs.items() if not k.startswith("rel_s_mlp."))
        self._build_from_specs(specs)
        self.eval()

    @classmethod
    def from_vocab(cls, vocab: dict, **kw) -> "SceneEncoder":
        """num_objs / num_preds as Sg2ScDiffModel derives them (EchoScene.py:37-43)."""
        return cls(num_objs=len(set(vocab["object_idx_to_name"])), num_preds=len(set(vocab["pred_idx_to_name"])), **kw)

    def load_reference_state_dict(self, state_dict, strict: bool = True):
        """Loads the encoder slice of a ``Sg2ScDiffModel`` state_dict (its other sub-modules -- the *_dc embeddings,
        LayoutDiff, ShapeDiff -- are dropped)."""
        pre = self.reference_prefixes()
        sub = {k: v for k, v in state_dict.items() if k.startswith(pre)}
        return self.load_state_dict(sub, strict=strict)

    def reference_prefixes(self) -> tuple:
        """key prefixes of the reference model's state_dict that belong to this encoder"""
        pre = self.PREFIXES + (("pred_embeddings_man_dc.",) if self.cfg.man_dc_preds else ())
        return pre if self.with_rel_s else tuple(p for p in pre if p != "rel_s_mlp.")

    # ---- handle ----
    _set_batch_stats_fn = "echo_scene_set_batch_stats"

    def _ensure(self, n_nodes: int, n_triples: int):
        ver = self._weights_version()
        cap = self._handle_key[1] if self._handle_key else (0, 0, False)
        if (self._handle is not None and self._handle_key[0] == ver and n_nodes <= cap[0] and n_triples <= cap[1]
                and (cap[2] or not self.training)):
            return self._apply_mode()
        self._destroy_handle()
        cap = (_next_pow2(n_nodes, 32), _next_pow2(n_triples, 128), bool(self.training or cap[2]))
        c = self.cfg
        d = _lib.SceneDesc(c.gconv_dim, c.add_dim, c.num_objs + 1, c.num_preds, c.num_layers, c.rel_s_hidden,
                           c.context_dim, cap[0], cap[1], 1e-5, int(c.man_dc_preds), int(cap[2]))
        arr, n, keep = _lib.weights_table(self.state_dict())
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_scene_create(C.byref(h), C.byref(d), arr, n))
        self._handle, self._handle_key = h, (ver, cap)
        self._apply_mode()

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().echo_scene_destroy(self._handle)
            self._handle = None

    def _prep(self, objs, triples, text_feat, rel_feat):
        _lib.require_cuda(objs, triples, text_feat, rel_feat)
        n, t = int(objs.shape[0]), int(triples.shape[0])
        if objs.dtype != torch.int64 or objs.dim() != 1:
            raise EchoError(f"objs must be (N,) int64, got {tuple(objs.shape)} {objs.dtype}")
        if n == 0:
            raise EchoError("empty scene")
        # nn.Embedding raises on these (EchoScene.py:149); one host read per scene (the reference does N + 44)
        lo, hi = int(objs.min()), int(objs.max())
        if lo < 0 or hi > self.cfg.num_objs:
            raise IndexError(f"object class ids [{lo}, {hi}] outside obj_embeddings_ec ({self.cfg.num_objs + 1} rows)")
        if self.clip:
            if text_feat is None or rel_feat is None:
                raise EchoError("use_clip=True needs text_feat (N,512) and rel_feat (T,512)")
            if tuple(text_feat.shape) != (n, self.cfg.add_dim) or tuple(rel_feat.shape) != (t, self.cfg.add_dim):
                raise EchoError(f"text_feat / rel_feat must be ({n},{self.cfg.add_dim}) / ({t},{self.cfg.add_dim}), got "
                                f"{tuple(text_feat.shape)} / {tuple(rel_feat.shape)}")
            text_feat, rel_feat = text_feat.float().contiguous(), rel_feat.float().contiguous()
        else:
            text_feat = rel_feat = None
        self._ensure(n, t)
        return objs.contiguous(), _lib.graph_for(triples, n), text_feat, rel_feat, n, t

    @torch.no_grad()
    def init_encoder(self, objs, triples, enc_text_feat=None, enc_rel_feat=None):
        """-> obj_embed (N,feat), pred_embed (T,feat), latent_obj_f (N,feat), latent_pred_f (None: never read by the
        sampling path, EchoScene.py:390-400)."""
        objs, g, tf, rf, n, t = self._prep(objs, triples, enc_text_feat, enc_rel_feat)
        f, dev = self.cfg.feat_dim, objs.device
        obj_embed, pred_embed = torch.empty(n, f, device=dev), torch.empty(t, f, device=dev)
        latent = torch.empty(n, f, device=dev)
        _lib.check(_lib.lib().echo_scene_init_encoder(self._handle, g.h, _lib.ptr(objs), _lib.ptr(tf), _lib.ptr(rf),
                                                      _lib.ptr(obj_embed), _lib.ptr(pred_embed), _lib.ptr(latent),
                                                      _lib.stream_ptr()))
        return obj_embed, pred_embed, latent, None

    @torch.no_grad()
    def manipulate(self, latent_f, objs, triples, dec_text_feat=None, dec_rel_feat=None):
        """latent_f (N, feat + embedding_dim) = [latent | change flag] -> obj_vecs (N,feat), pred_vecs (None: unused by
        the sampling path), obj_embed (N,feat), pred_embed (T,feat)."""
        objs, g, tf, rf, n, t = self._prep(objs, triples, dec_text_feat, dec_rel_feat)
        _lib.require_cuda(latent_f)
        f, dev = self.cfg.feat_dim, objs.device
        if tuple(latent_f.shape) != (n, f + self.cfg.gconv_dim):
            raise EchoError(f"latent_f must be ({n},{f + self.cfg.gconv_dim}), got {tuple(latent_f.shape)}")
        latent_f = latent_f.float().contiguous()
        obj_embed, pred_embed = torch.empty(n, f, device=dev), torch.empty(t, f, device=dev)
        latent = torch.empty(n, f, device=dev)
        _lib.check(_lib.lib().echo_scene_manipulate(self._handle, g.h, _lib.ptr(latent_f), _lib.ptr(objs), _lib.ptr(tf),
                                                    _lib.ptr(rf), _lib.ptr(latent), _lib.ptr(obj_embed),
                                                    _lib.ptr(pred_embed), _lib.stream_ptr()))
        return latent, None, obj_embed, pred_embed

    @torch.no_grad()
    def rel_s(self, x):
        """self.rel_s_mlp(x): (M, feat) -> (M, context_dim).  Under train() its BatchNorm1d uses the statistics of the M rows."""
        if not self.with_rel_s:
            raise EchoError("this SceneEncoder was built without rel_s_mlp (layout-only model)")
        _lib.require_cuda(x)
        if x.dim() != 2 or x.shape[1] != self.cfg.feat_dim:
            raise EchoError(f"rel_s input must be (M,{self.cfg.feat_dim}), got {tuple(x.shape)}")
        x = x.float().contiguous()
        m = int(x.shape[0])
        self._ensure(max(m, 1), 0)
        out = torch.empty(m, self.cfg.context_dim, device=x.device)
        _lib.check(_lib.lib().echo_scene_rel_s(self._handle, _lib.ptr(x), m, _lib.ptr(out), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def encode(self, objs, triples, text_feat=None, rel_feat=None, change=None, shape_cond: bool = True) -> Dict[str, torch.Tensor]:
        """The encoder stage of ``Sg2ScDiffModel.sample`` (EchoScene.py:388-410) in one call.  ``change`` (N, embedding_dim)
        defaults to the zero flag of ``sample``.  -> obj_embed, latent (N,feat); uc_s, c_s (N,1,context_dim) when
        ``shape_cond`` (gen_shape=True)."""
        self._check_eval()
        objs, g, tf, rf, n, t = self._prep(objs, triples, text_feat, rel_feat)
        c, dev = self.cfg, objs.device
        if change is not None:
            _lib.require_cuda(change)
            if tuple(change.shape) != (n, c.gconv_dim):
                raise EchoError(f"change must be ({n},{c.gconv_dim}), got {tuple(change.shape)}")
            change = change.float().contiguous()
        out = {"obj_embed": torch.empty(n, c.feat_dim, device=dev), "latent": torch.empty(n, c.feat_dim, device=dev)}
        if shape_cond and not self.with_rel_s:
            raise EchoError("shape_cond=True needs rel_s_mlp; this SceneEncoder was built without it (layout-only model)")
        if shape_cond:
            out["uc_s"] = torch.empty(n, 1, c.context_dim, device=dev)
            out["c_s"] = torch.empty(n, 1, c.context_dim, device=dev)
        _lib.check(_lib.lib().echo_scene_encode(self._handle, g.h, _lib.ptr(objs), _lib.ptr(tf), _lib.ptr(rf),
                                                _lib.ptr(change), _lib.ptr(out["obj_embed"]), _lib.ptr(out["latent"]),
                                                _lib.ptr(out.get("uc_s")), _lib.ptr(out.get("c_s")), _lib.stream_ptr()))
        return out


# ----------------------------------------------------------------------------------------------------------------------
# layout denoiser
# ----------------------------------------------------------------------------------------------------------------------
def _fill_levels(desc, channel_mult, attention_resolutions):
    desc.num_levels = len(channel_mult)
    for i, m in enumerate(channel_mult):
        desc.channel_mult[i] = int(m)
    desc.num_attention_resolutions = len(attention_resolutions)
    for i, a in enumerate(attention_resolutions):
        desc.attention_resolutions[i] = int(a)


class UNet1DModel(_SpecModule):
    """The layout denoiser — denoise_net.py:451-806.  Same keyword arguments as the reference; the settings the
    default ``*_mp`` configs use (crossattn + message passing + spatial transformer) are the implemented ones."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=1, use_checkpoint=False, use_fp16=False,
                 num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=False,
                 transformer_depth=1, concat_dim=None, crossattn_dim=None, conditioning_key="crossattn",
                 using_clip=True, enable_t_emb=False, precision: str = "fp32", time_num: int = 1000,
                 beta_start: float = 1e-4, beta_end: float = 0.02):
        super().__init__()
        unsupported = dict(dims=(dims, 1), conditioning_key=(conditioning_key, "crossattn"),
                           use_spatial_transformer=(use_spatial_transformer, True),
                           transformer_depth=(transformer_depth, 1), use_scale_shift_norm=(use_scale_shift_norm, False),
                           resblock_updown=(resblock_updown, False), conv_resample=(conv_resample, True),
                           num_head_channels=(num_head_channels, -1), dropout=(dropout, 0))
        for k, (got, want) in unsupported.items():
            if got != want:
                raise EchoError(f"UNet1DModel: {k}={got!r} is outside the hot path (implemented: {want!r})")
        self.conditioning_key = conditioning_key
        self.cfg = arch.UNet1DConfig(in_channels=in_channels, out_channels=out_channels, model_channels=model_channels,
                                     channel_mult=tuple(channel_mult), num_res_blocks=num_res_blocks,
                                     attention_resolutions=tuple(attention_resolutions), num_heads=num_heads,
                                     concat_dim=concat_dim, crossattn_dim=crossattn_dim, using_clip=using_clip,
                                     enable_t_emb=enable_t_emb)
        self.precision = precision
        self.time_num, self.beta_start, self.beta_end = time_num, beta_start, beta_end
        self._build_from_specs(arch.unet1d_specs(self.cfg))
        self.eval()

    def set_schedule(self, time_num, beta_start=1e-4, beta_end=0.02):
        if (time_num, beta_start, beta_end) != (self.time_num, self.beta_start, self.beta_end):
            self.time_num, self.beta_start, self.beta_end = time_num, beta_start, beta_end
            self._destroy_handle()

    _set_batch_stats_fn = "echo_layout_set_batch_stats"

    def _ensure(self, n_nodes, n_triples):
        ver = self._weights_version()
        cap = self._handle_key[1] if self._handle_key else (0, 0, False)
        if (self._handle is not None and self._handle_key[0] == ver and n_nodes <= cap[0] and n_triples <= cap[1]
                and (cap[2] or not self.training)):
            return self._apply_mode()
        self._destroy_handle()
        cap = (_next_pow2(n_nodes, 32), _next_pow2(n_triples, 128), bool(self.training or cap[2]))
        c = self.cfg
        d = _lib.LayoutDesc()
        d.keep_train_weights = int(cap[2])
        d.in_channels, d.out_channels, d.model_channels = c.in_channels, c.out_channels, c.model_channels
        _fill_levels(d, c.channel_mult, c.attention_resolutions)
        d.num_res_blocks, d.num_heads, d.context_dim = c.num_res_blocks, c.num_heads, c.crossattn_dim
        d.obj_embed_dim, d.gconv_dim, d.enable_t_emb = c.obj_embed_dim, c.gconv_dim, int(c.enable_t_emb)
        d.max_nodes, d.max_triples = cap[:2]
        d.precision = _lib.precision_code(self.precision)
        d.time_num, d.beta_start, d.beta_end = self.time_num, self.beta_start, self.beta_end
        dev = next(self.parameters()).device
        arr, n, keep = _lib.weights_table(self.state_dict(),
                                          {"__timestep_freqs": _lib.timestep_freqs(c.model_channels, dev)})
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_layout_create(C.byref(h), C.byref(d), arr, n))
        self._handle, self._handle_key = h, (ver, cap)
        self._apply_mode()

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().echo_layout_destroy(self._handle)
            self._handle = None

    @torch.no_grad()
    def forward(self, box_t, obj_embed, triples, timesteps=None, context=None, y=None, **kwargs):
        """box_t (N,8), obj_embed (N,640), triples (T,3) i64, timesteps (N,) i64 -> (N,8,1).  ``context`` is accepted
        and ignored exactly as the reference ignores it in crossattn mode (denoise_net.py:791-792).  Under train() box_graph_cov's
        BatchNorm1d layers use the statistics of the batch (forward values of get_loss_iter's denoiser call)."""
        self._check_train_values()
        _lib.require_cuda(box_t, obj_embed, triples, timesteps)
        n = box_t.shape[0]
        box_t = box_t.float().contiguous()
        obj_embed = obj_embed.float().contiguous()
        timesteps = timesteps.to(torch.int64).contiguous()
        assert box_t.shape == (n, self.cfg.in_channels) and obj_embed.shape == (n, self.cfg.obj_embed_dim)
        assert timesteps.shape == (n,)
        self._ensure(n, triples.shape[0])
        g = _lib.graph_for(triples, n)
        out = torch.empty(n, self.cfg.out_channels, device=box_t.device)
        _lib.check(_lib.lib().echo_layout_forward(self._handle, g.h, _lib.ptr(box_t), _lib.ptr(obj_embed),
                                                  _lib.ptr(timesteps), _lib.ptr(out), _lib.stream_ptr()))
        return out.unsqueeze(-1)

    @torch.no_grad()
    def ddpm_step(self, x_t, obj_embed, triples, t: int, noise):
        """One iteration of p_sample_loop_sg (forward + posterior update) — diffusion_ddpm.py:296-345."""
        self._check_eval()
        _lib.require_cuda(x_t, obj_embed, triples, noise)
        n = x_t.shape[0]
        if (tuple(x_t.shape) != (n, self.cfg.in_channels) or tuple(obj_embed.shape) != (n, self.cfg.obj_embed_dim)
                or tuple(noise.shape) != tuple(x_t.shape)):
            raise EchoError(f"ddpm_step: x_t / noise must be ({n},{self.cfg.in_channels}) and obj_embed ({n},"
                            f"{self.cfg.obj_embed_dim}); got {tuple(x_t.shape)}, {tuple(noise.shape)}, {tuple(obj_embed.shape)}")
        if not 0 <= int(t) < self.time_num:
            raise EchoError(f"ddpm_step: t = {t} outside the {self.time_num}-step schedule")
        self._ensure(n, triples.shape[0])
        g = _lib.graph_for(triples, n)
        x_t, obj_embed, noise = x_t.float().contiguous(), obj_embed.float().contiguous(), noise.float().contiguous()
        out = torch.empty_like(x_t)
        _lib.check(_lib.lib().echo_layout_step(self._handle, g.h, _lib.ptr(x_t), _lib.ptr(obj_embed), int(t),
                                               _lib.ptr(noise), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def schedule_tables(self) -> torch.Tensor:
        self._ensure(1, 1)
        out = torch.empty(5, self.time_num, dtype=torch.float32)
        _lib.check(_lib.lib().echo_layout_schedule(self._handle, out.data_ptr()))
        return out


# ----------------------------------------------------------------------------------------------------------------------
# shape denoiser
# ----------------------------------------------------------------------------------------------------------------------
class UNet3DModel(_SpecModule):
    """The shape denoiser — openai_model_3d.py:452-863, same keyword arguments."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 messsage_passing=True, conditioning_key=None, enable_t_emb=False, precision: str = "fp32",
                 ddim_steps: int = 100, timesteps: int = 1000, linear_start: float = 0.00085,
                 linear_end: float = 0.012):
        super().__init__()
        unsupported = dict(dims=(dims, 3), conditioning_key=(conditioning_key, "crossattn"),
                           use_spatial_transformer=(use_spatial_transformer, True), messsage_passing=(messsage_passing, True),
                           transformer_depth=(transformer_depth, 1), use_scale_shift_norm=(use_scale_shift_norm, False),
                           resblock_updown=(resblock_updown, False), conv_resample=(conv_resample, True),
                           num_head_channels=(num_head_channels, -1), num_classes=(num_classes, None),
                           n_embed=(n_embed, None), legacy=(legacy, False), dropout=(dropout, 0))
        for k, (got, want) in unsupported.items():
            if got != want:
                raise EchoError(f"UNet3DModel: {k}={got!r} is outside the hot path (implemented: {want!r})")
        self.conditioning_key = conditioning_key
        self.messsage_passing = messsage_passing
        self.cfg = arch.UNet3DConfig(in_channels=in_channels, out_channels=out_channels, model_channels=model_channels,
                                     channel_mult=tuple(channel_mult), num_res_blocks=num_res_blocks,
                                     attention_resolutions=tuple(attention_resolutions), num_heads=num_heads,
                                     context_dim=context_dim, image_size=image_size, enable_t_emb=enable_t_emb)
        self.precision = precision
        self.ddim_steps, self.timesteps_total = ddim_steps, timesteps
        self.linear_start, self.linear_end = linear_start, linear_end
        self._build_from_specs(arch.unet3d_specs(self.cfg))
        self.eval()

    def set_schedule(self, ddim_steps, timesteps=1000, linear_start=0.00085, linear_end=0.012):
        new = (ddim_steps, timesteps, linear_start, linear_end)
        if new != (self.ddim_steps, self.timesteps_total, self.linear_start, self.linear_end):
            self.ddim_steps, self.timesteps_total, self.linear_start, self.linear_end = new
            self._destroy_handle()

    _set_batch_stats_fn = "echo_shape_set_batch_stats"

    def _ensure(self, n_nodes, n_triples, n_local=None, sharded_call=False):
        """sharded_call: the caller is embed_local / trunk_local (codes come from outside): a handle built for a shard also serves
        a call whose local range happens to be the whole (sub)graph, so the handle is not rebuilt back and forth."""
        n_local = n_nodes if n_local is None else n_local
        ver = self._weights_version()
        cap = self._handle_key[1] if self._handle_key else (0, 0, 0, False)
        if (self._handle is not None and self._handle_key[0] == ver and n_nodes <= cap[0] and n_triples <= cap[1]
                and n_local <= cap[2] and ((cap[2] == cap[0]) == (n_local == n_nodes) or (sharded_call and cap[2] != cap[0]))
                and (cap[3] or not self.training)):
            return self._apply_mode()
        self._destroy_handle()
        cap = (max(n_nodes, 1), _next_pow2(n_triples, 128), max(n_local, 1), bool(self.training or cap[3]))
        c = self.cfg
        d = _lib.ShapeDesc()
        d.keep_train_weights = int(cap[3])
        d.in_channels, d.out_channels, d.model_channels = c.in_channels, c.out_channels, c.model_channels
        _fill_levels(d, c.channel_mult, c.attention_resolutions)
        d.num_res_blocks, d.num_heads, d.context_dim = c.num_res_blocks, c.num_heads, c.context_dim
        d.gconv_dim, d.enable_t_emb, d.latent_size = c.gconv_dim, int(c.enable_t_emb), c.image_size
        d.max_nodes, d.max_triples, d.max_local_nodes = cap[:3]
        d.precision = _lib.precision_code(self.precision)
        d.timesteps, d.ddim_steps = self.timesteps_total, self.ddim_steps
        d.linear_start, d.linear_end = self.linear_start, self.linear_end
        dev = next(self.parameters()).device
        arr, n, keep = _lib.weights_table(self.state_dict(),
                                          {"__timestep_freqs": _lib.timestep_freqs(c.model_channels, dev)})
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_shape_create(C.byref(h), C.byref(d), arr, n))
        self._handle, self._handle_key = h, (ver, cap)
        self._apply_mode()

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().echo_shape_destroy(self._handle)
            self._handle = None

    def _prep(self, x, obj_embed, triples):
        _lib.require_cuda(x, obj_embed, triples)
        n = x.shape[0]
        c = self.cfg
        assert x.shape == (n, c.in_channels, c.image_size, c.image_size, c.image_size), tuple(x.shape)
        obj_embed = obj_embed.reshape(obj_embed.shape[0], -1).float().contiguous()
        assert obj_embed.shape[1] == c.context_dim
        if obj_embed.shape[0] != n:
            raise EchoError(f"obj_embed has {obj_embed.shape[0]} rows for {n} objects")
        return n, x.float().contiguous(), obj_embed

    @torch.no_grad()
    def forward(self, x, obj_embed, triples, timesteps=None, context=None, y=None, **kwargs):
        """x (N,3,16,16,16), obj_embed (N,1,1280), triples (T,3) i64, timesteps (N,) i64 -> e_t like x.  ``context``
        is accepted and ignored as in the reference ("we dont use the previous context", openai_model_3d.py:843-844).  Under train()
        shape_code_graph_cov's BatchNorm1d layers use the statistics of the batch (forward values of p_losses' denoiser call)."""
        self._check_train_values()
        n, x, obj_embed = self._prep(x, obj_embed, triples)
        timesteps = timesteps.to(torch.int64).contiguous()
        self._ensure(n, triples.shape[0])
        g = _lib.graph_for(triples, n)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().echo_shape_forward(self._handle, g.h, _lib.ptr(x), _lib.ptr(obj_embed),
                                                 _lib.ptr(timesteps), _lib.ptr(out), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def ddim_step(self, x_t, obj_embed, triples, index: int, out: Optional[torch.Tensor] = None):
        """One iteration of DDIMSampler.ddim_sampling: forward at ddim_timesteps[index] + x_prev update (eta = 0)."""
        self._check_eval()
        n, x_t, obj_embed = self._prep(x_t, obj_embed, triples)
        self._ensure(n, triples.shape[0])
        g = _lib.graph_for(triples, n)
        if out is None:
            out = torch.empty_like(x_t)
        elif (out.shape != x_t.shape or out.dtype != torch.float32 or out.device != x_t.device or not out.is_contiguous()
              or out.data_ptr() == x_t.data_ptr()):
            raise EchoError("ddim_step: `out` must be a distinct contiguous float32 tensor shaped and placed like x_t")
        _lib.check(_lib.lib().echo_shape_step(self._handle, g.h, _lib.ptr(x_t), _lib.ptr(obj_embed), int(index),
                                              _lib.ptr(out), _lib.stream_ptr()))
        return out

    def set_step_index(self, index: int):
        """Stream-ordered write of the DDIM index that ``ddim_step`` / ``trunk_local`` read when called with
        ``index=_lib.INDEX_FROM_DEVICE``: such a step launches the same kernels for every index, so a chain can capture it once
        in a CUDA graph (with the NCCL all-gather of a sharded step) and replay it -- set the index, replay."""
        if self._handle is None:
            raise EchoError("set_step_index: no handle yet (run a step, or _ensure(), first)")
        _lib.check(_lib.lib().echo_shape_set_index(self._handle, int(index), _lib.stream_ptr()))

    # ---- per-object sharding: embed local objects, all-gather the 64-d codes, run the trunk locally ----
    @torch.no_grad()
    def embed_local(self, x_local, n_nodes, n_triples):
        _lib.require_cuda(x_local)
        nl = x_local.shape[0]
        self._ensure(n_nodes, n_triples, nl, sharded_call=True)
        codes = torch.empty(nl, self.cfg.gconv_dim, device=x_local.device)
        x_local = x_local.float().contiguous()
        _lib.check(_lib.lib().echo_shape_embed(self._handle, _lib.ptr(x_local), nl, _lib.ptr(codes), _lib.stream_ptr()))
        return codes

    @torch.no_grad()
    def trunk_local(self, x_local, obj_begin, codes_all, obj_embed_all, triples, index: int = -1, timesteps_all=None,
                    out: Optional[torch.Tensor] = None, codes_stream: Optional[torch.cuda.Stream] = None,
                    restrict_to_components: bool = False):
        """codes_stream: the stream the embed + all-gather of `codes_all` were queued on; when given, only the echo chain
        waits for it and the first trunk blocks overlap the exchange.
        restrict_to_components: run the echo GCN only on the connected components of the scene graph that contain the local
        objects (`shard.echo_components`; exact -- message passing never leaves a component).  In a collated batch of scenes that
        is the rank's own scenes instead of the whole batch: the redundant part of the sharded step no longer grows with the
        number of ranks."""
        _lib.require_cuda(x_local, codes_all, obj_embed_all, triples)
        if restrict_to_components:
            key = (triples.data_ptr(), tuple(triples.shape), triples._version, int(codes_all.shape[0]), int(obj_begin), int(x_local.shape[0]),
                   obj_embed_all.data_ptr())
            hit = self.__dict__.setdefault("_component_cache", {}).get(key)
            if hit is None:
                from . import shard
                res = shard.echo_components(triples, codes_all.shape[0], int(obj_begin), int(x_local.shape[0]))
                if res is not None:
                    nodes, tri_sub, begin_sub = res
                    res = (nodes, tri_sub, begin_sub, obj_embed_all.reshape(codes_all.shape[0], -1).index_select(0, nodes).contiguous())
                if len(self._component_cache) > 16:
                    self._component_cache.clear()
                hit = self._component_cache[key] = (res, triples, obj_embed_all)    # the keyed tensors stay alive with the entry
            if hit[0] is not None:
                nodes, tri_sub, begin_sub, uc_sub = hit[0]
                if codes_stream is not None:   # the gather of the exchanged codes belongs behind the all-gather, on its stream
                    with torch.cuda.stream(codes_stream):
                        codes_sub = codes_all.index_select(0, nodes)
                else:
                    codes_sub = codes_all.index_select(0, nodes)
                t_sub = None if timesteps_all is None else timesteps_all.index_select(0, nodes)
                return self.trunk_local(x_local, begin_sub, codes_sub, uc_sub, tri_sub, index=index, timesteps_all=t_sub, out=out,
                                        codes_stream=codes_stream)
        n = codes_all.shape[0]
        nl = x_local.shape[0]
        self._ensure(n, triples.shape[0], nl, sharded_call=True)
        g = _lib.graph_for(triples, n)
        x_local = x_local.float().contiguous()
        obj_embed_all = obj_embed_all.reshape(n, -1).float().contiguous()
        codes_all = codes_all.float().contiguous()
        out = torch.empty_like(x_local) if out is None else out
        t = None if timesteps_all is None else timesteps_all.to(torch.int64).contiguous()
        if codes_stream is not None:
            _lib.check(_lib.lib().echo_shape_trunk_async(self._handle, g.h, _lib.ptr(x_local), int(obj_begin), nl,
                                                         _lib.ptr(codes_all), _lib.ptr(obj_embed_all), _lib.ptr(t), int(index),
                                                         _lib.ptr(out), codes_stream.cuda_stream, _lib.stream_ptr()))
            return out
        _lib.check(_lib.lib().echo_shape_trunk(self._handle, g.h, _lib.ptr(x_local), int(obj_begin), nl,
                                               _lib.ptr(codes_all), _lib.ptr(obj_embed_all), _lib.ptr(t), int(index),
                                               _lib.ptr(out), _lib.stream_ptr()))
        return out

    def last_latent(self, n_nodes) -> torch.Tensor:
        """latent_shape_rel (N, context_dim) of the last call (a copy)."""
        out = torch.empty(n_nodes, self.cfg.context_dim, device=next(self.parameters()).device)
        _lib.check(_lib.lib().echo_shape_latent(self._handle, int(n_nodes), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def schedule_tables(self):
        self._ensure(1, 1)
        coef = torch.empty(self.ddim_steps_effective(), 4, dtype=torch.float32)
        ts = torch.empty(self.ddim_steps_effective(), dtype=torch.int32)
        _lib.check(_lib.lib().echo_shape_schedule(self._handle, coef.data_ptr(), ts.data_ptr()))
        return coef, ts

    def ddim_steps_effective(self) -> int:
        c = self.timesteps_total // self.ddim_steps
        return len(range(0, self.timesteps_total, c))


class DiffusionUNet(nn.Module):
    """Conditioning router around UNet3DModel — diffusion_shape/network.py:11-43."""

    def __init__(self, unet_params, vq_conf=None, conditioning_key=None, **extra):
        super().__init__()
        self.conditioning_key = conditioning_key
        params = dict(unet_params)
        params["conditioning_key"] = conditioning_key
        params.update(extra)
        self.diffusion_net = UNet3DModel(**params)

    def forward(self, x, obj_embed, triples, t, c_concat: list = None, c_crossattn: list = None):
        if self.conditioning_key != "crossattn":
            raise EchoError("only conditioning_key='crossattn' is on the hot path")
        cc = torch.cat(c_crossattn, 1)
        return self.diffusion_net(x, obj_embed, triples, t, context=cc)


# ----------------------------------------------------------------------------------------------------------------------
# VQ-VAE decode
# ----------------------------------------------------------------------------------------------------------------------
class VQVAE(_SpecModule):
    """model/networks/vqvae_networks/network.py:56-103 (same constructor).  ``decode_no_quant(h)`` = quantize ->
    post_quant_conv -> Decoder3D, what EchoToShape.rel2shape calls on the sampled latents (echo2shape.py:522).  state_dict
    keys are the reference's (quantize.embedding.weight, post_quant_conv.*, decoder.*).  With ``with_encoder=True`` the
    module also owns encoder.* and quant_conv.* and ``encode_no_quant(x)`` (network.py:84-88, the call of the training
    step on the ground-truth SDFs, echo2shape.py:334-364) runs through echo_vqvae_encode in fp32; without it those entries
    of a reference checkpoint are accepted and ignored by ``load_state_dict`` and ``encode*`` raise."""

    def __init__(self, ddconfig, n_embed, embed_dim, remap=None, sane_index_shape=False, precision: str = "fp32",
                 with_encoder: bool = False):
        super().__init__()
        if remap is not None:
            raise EchoError("VQVAE: remap is outside the hot path")
        dd = dict(ddconfig)
        if dd.get("attn_resolutions"):
            raise EchoError("VQVAE: attn_resolutions != [] is outside the hot path (config/vqvae_snet.yaml uses [])")
        self.cfg = arch.VQVAEConfig(embed_dim=embed_dim, n_embed=n_embed, z_channels=dd["z_channels"], resolution=dd["resolution"],
                                    out_ch=dd["out_ch"], ch=dd["ch"], ch_mult=tuple(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"])
        self.ddconfig, self.n_embed, self.embed_dim = dd, n_embed, embed_dim
        self.precision = precision
        self.with_encoder = bool(with_encoder)
        specs = arch.vqvae_decode_specs(self.cfg)
        if self.with_encoder:
            if dd.get("double_z", False):
                raise EchoError("VQVAE: double_z=True is outside the hot path (config/vqvae_snet.yaml uses False)")
            if dd.get("in_channels", 1) != 1:
                raise EchoError("VQVAE: the encoder takes 1-channel SDF volumes (config/vqvae_snet.yaml)")
            enc = arch.vqvae_encode_specs(self.cfg, 1)
            enc.update(specs)
            specs = enc
        self._build_from_specs(specs)
        self._enc_handle, self._enc_key = None, None
        self.eval()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        own = state_dict
        if not self.with_encoder:
            own = {k: v for k, v in state_dict.items() if not (k.startswith("encoder.") or k.startswith("quant_conv."))}
        return super().load_state_dict(own, strict=strict, **kw)

    def _desc(self, cap: int, precision: str):
        c = self.cfg
        d = _lib.VqvaeDesc()
        d.embed_dim, d.n_embed, d.z_channels, d.latent_size = c.embed_dim, c.n_embed, c.z_channels, c.latent_size
        d.ch, d.num_levels, d.num_res_blocks, d.out_ch = c.ch, len(c.ch_mult), c.num_res_blocks, c.out_ch
        for i, m in enumerate(c.ch_mult):
            d.ch_mult[i] = m
        d.max_objects = cap
        d.precision = _lib.precision_code(precision)
        return d

    def _ensure_encoder(self, n):
        ver = self._weights_version()
        if self._enc_handle is not None and self._enc_key[0] == ver and n <= self._enc_key[1]:
            return
        self._destroy_encoder()
        cap = max(n, 1)
        d = self._desc(cap, "fp32")                      # encode_no_quant is fp32 only (echo_vqvae_encoder_create)
        arr, nw, keep = _lib.weights_table(self.state_dict())
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_vqvae_encoder_create(C.byref(h), C.byref(d), arr, nw))
        self._enc_handle, self._enc_key = h, (ver, cap)

    def _destroy_encoder(self):
        if getattr(self, "_enc_handle", None) is not None:
            _lib.lib().echo_vqvae_destroy(self._enc_handle)
            self._enc_handle = None

    def _ensure(self, n):
        ver = self._weights_version()
        if self._handle is not None and self._handle_key[0] == ver and n <= self._handle_key[1]:
            return
        self._destroy_handle()
        cap = max(n, 1)
        d = self._desc(cap, self.precision)
        arr, nw, keep = _lib.weights_table(self.state_dict())
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_vqvae_create(C.byref(h), C.byref(d), arr, nw))
        self._handle, self._handle_key = h, (ver, cap)

    def _destroy_handle(self):
        if self._handle is not None:
            _lib.lib().echo_vqvae_destroy(self._handle)
            self._handle = None
        self._destroy_encoder()

    max_chunk = 32   # objects decoded per library call: the decoder keeps ~0.5 GB of activations per object (bf16)

    @torch.no_grad()
    def decode_no_quant(self, h, force_not_quantize=False, return_indices: bool = False):
        """h (N, 3, 16, 16, 16) -> SDF (N, 1, 64, 64, 64), network.py:95-103.  Objects are independent (per-sample
        GroupNorm and attention), so large batches are decoded in chunks of ``max_chunk`` with bit-identical results."""
        self._check_eval()
        if force_not_quantize:
            raise EchoError("VQVAE.decode_no_quant(force_not_quantize=True) is outside the hot path")
        _lib.require_cuda(h)
        c = self.cfg
        n, L = h.shape[0], c.latent_size
        assert h.shape == (n, c.z_channels, L, L, L), tuple(h.shape)
        h = h.float().contiguous()
        chunk = max(1, int(self.max_chunk))
        self._ensure(min(n, chunk))
        out = torch.empty(n, c.out_ch, c.resolution, c.resolution, c.resolution, device=h.device)
        idx = torch.empty(n * L * L * L, dtype=torch.int32, device=h.device) if return_indices else None
        for b in range(0, n, chunk):
            e = min(n, b + chunk)
            _lib.check(_lib.lib().echo_vqvae_decode(self._handle, _lib.ptr(h[b:e]), e - b, _lib.ptr(out[b:e]),
                                                    _lib.ptr(idx[b * L * L * L:e * L * L * L]) if idx is not None else None,
                                                    _lib.stream_ptr()))
        return (out, idx) if return_indices else out

    max_encode_chunk = 8   # objects encoded per library call: the fp32 encoder keeps ~0.6 GB of activations per object

    @torch.no_grad()
    def encode_no_quant(self, x):
        """x (N, 1, 64, 64, 64) SDF -> latents (N, 3, 16, 16, 16): encoder -> quant_conv, no quantisation (network.py:84-88).
        Objects are independent, so large batches run in chunks of ``max_encode_chunk`` with bit-identical results."""
        if not self.with_encoder:
            raise EchoError("VQVAE.encode_no_quant needs VQVAE(..., with_encoder=True) (the decode-only module owns no encoder "
                            "weights)")
        self._check_eval()
        _lib.require_cuda(x)
        c = self.cfg
        n, R, L = x.shape[0], c.resolution, c.latent_size
        if tuple(x.shape) != (n, 1, R, R, R):
            raise EchoError(f"encode_no_quant input must be (N,1,{R},{R},{R}), got {tuple(x.shape)}")
        x = x.float().contiguous()
        chunk = max(1, int(self.max_encode_chunk))
        self._ensure_encoder(min(n, chunk))
        out = torch.empty(n, c.embed_dim, L, L, L, device=x.device)
        for b in range(0, n, chunk):
            e = min(n, b + chunk)
            _lib.check(_lib.lib().echo_vqvae_encode(self._enc_handle, _lib.ptr(x[b:e]), e - b, _lib.ptr(out[b:e]),
                                                    _lib.stream_ptr()))
        return out

    def forward(self, *a, **k):
        raise EchoError("VQVAE.forward (encode + quantise + decode, the VQ-VAE's own training step) is outside the hot path; "
                        "use encode_no_quant / decode_no_quant")

    def encode(self, *a, **k):
        raise EchoError("VQVAE.encode (with quantisation losses) is outside the hot path; use encode_no_quant")
