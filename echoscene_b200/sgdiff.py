"""``SGDiff`` -- the facade ``scripts/eval_3dfront.py`` talks to (model/SGDiff.py:6-129), built on the B200 components
from the reference's own YAML configuration files, without importing the reference.

    opt  = "config/full_mp.yaml"                       # or the loaded dict / OmegaConf object
    model = SGDiff("echoscene", opt, vocab, replace_latent=False, with_changes=True, residual=True, clip=True)
    model.load_networks(exp="../full_mp", epoch=2050)  # model{epoch}.pth as SGDiff.save writes it
    model = model.cuda().eval()
    out = model.sample_box_and_shape(dec_objs, dec_triples, text_feat, rel_feat, gen_shape=True)

Construction follows the reference's constructors: ``Sg2ScDiffModel`` / ``Sg2BoxDiffModel`` with ``embedding_dim=64,
mlp_normalization='batch', gconv_num_layers=5`` (SGDiff.py:21-26); ``EchoToLayout`` = ``UNet1DModel(**layout_branch.
denoiser_kwargs)`` under ``DiffusionPoint(**layout_branch.diffusion_kwargs)`` (echo2layout.py:15-24); ``EchoToShape`` =
``DiffusionUNet(unet.params)`` with the DDPM schedule of ``model.params`` and the VQ-VAE of ``vq_cfg`` / ``vq_ckpt``
(echo2shape.py:62-85, 174-190; model_utils.py:7-32), ``ddim_steps = 100`` (7 when ``misc.debug == 1``, echo2shape.py:116-120).

``save`` / ``state_dict(epoch, counter)`` write the reference's checkpoint format (a file the unmodified reference loads with
strict=True after a ``load_networks`` round trip).  ``train()`` + ``forward_mani`` compute the VALUES of the training forward (losses on batch statistics; ``with_vq_encoder=True``
for the shape branch's VQ-VAE encode); the backward pass is outside this path.
"""
from __future__ import annotations

import os
from typing import Any, Optional

import torch

from . import modules, samplers, scene
from ._lib import EchoError


def _load_yaml(path: str) -> dict:
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def _get(cfg: Any, path: str, default: Any = None) -> Any:
    """cfg['a']['b'] / cfg.a.b for dicts, OmegaConf objects and attribute namespaces alike."""
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key, None)
        else:
            try:
                cur = cur[key]
            except (KeyError, TypeError, IndexError):
                cur = getattr(cur, key, None)
    return default if cur is None else cur


def _plain(x: Any) -> Any:
    """dict / list copies of config nodes (OmegaConf containers included) so they can be splatted as keyword arguments"""
    if isinstance(x, dict) or hasattr(x, "keys"):
        return {k: _plain(x[k]) for k in x.keys()}
    if isinstance(x, (list, tuple)) or type(x).__name__ == "ListConfig":
        return [_plain(v) for v in x]
    return x


class SGDiff:
    """Same constructor arguments as model/SGDiff.py:8-9, plus ``precision`` ('fp32' parity mode | 'bf16' tensor cores) and
    ``config_dir`` (where relative ``df_cfg`` / ``vq_cfg`` paths of the YAML are looked up; the reference resolves them against
    the working directory of its scripts)."""

    def __init__(self, type, diff_opt, vocab, replace_latent=False, with_changes=True, residual=False, gconv_pooling="avg",
                 with_angles=False, clip=True, separated=False, precision: str = "fp32", config_dir: Optional[str] = None,
                 with_vq_encoder: bool = False):
        assert type in ["echoscene", "echolayout"], "{} is not included".format(type)
        assert replace_latent is not None and with_changes is not None
        if separated:
            raise EchoError("separated=True (gconv_net_ec_rel_s/_l, EchoScene.py:66-80) is outside the hot path")
        if isinstance(diff_opt, (str, os.PathLike)):
            config_dir = config_dir or os.path.dirname(os.path.abspath(diff_opt))
            diff_opt = _load_yaml(diff_opt)
        self.type_, self.vocab, self.with_angles, self.diff_opt = type, vocab, with_angles, diff_opt
        self.epoch, self.counter = 0, 0
        self.precision, self._config_dir = precision, config_dir
        self._foreign, self._foreign_opt, self._foreign_vqvae = {}, None, None

        layout_only = type == "echolayout"
        self.encoder = modules.SceneEncoder.from_vocab(vocab, embedding_dim=64, gconv_num_layers=5, residual=residual,
                                                       use_clip=clip, gconv_pooling=gconv_pooling, mlp_normalization="batch",
                                                       with_rel_s=not layout_only, man_dc_preds=layout_only)
        # ---- layout branch (echo2layout.py:9-30)
        if _get(diff_opt, "layout_branch.denoiser") != "unet1d":
            raise EchoError("layout_branch.denoiser must be 'unet1d' (echo2layout.py:15-18)")
        if not _get(diff_opt, "layout_branch.relation_condition", True):
            raise EchoError("layout_branch.relation_condition=false raises in the reference too (echo2layout.py:94-97)")
        self.unet1d = modules.UNet1DModel(**_plain(_get(diff_opt, "layout_branch.denoiser_kwargs")), precision=precision)
        self.layout = samplers.DiffusionPoint(self.unet1d, config=_get(diff_opt, "layout_branch"),
                                              **_plain(_get(diff_opt, "layout_branch.diffusion_kwargs", {})))
        t_dim = int(_get(diff_opt, "layout_branch.translation_dim", 3))
        s_dim = int(_get(diff_opt, "layout_branch.size_dim", 3))
        box_dim = t_dim + s_dim + int(_get(diff_opt, "layout_branch.angle_dim"))
        if box_dim != self.unet1d.cfg.in_channels:
            raise EchoError(f"size_dim + translation_dim + angle_dim = {box_dim} != denoiser in_channels {self.unet1d.cfg.in_channels}")
        # ---- shape branch (echo2shape.py:62-120)
        self.unet3d = self.vqvae = None
        ddim_steps, z_shape = 100, (3, 16, 16, 16)
        if not layout_only:
            df = _load_yaml(self._resolve(_get(diff_opt, "shape_branch.df_cfg")))
            vq = _load_yaml(self._resolve(_get(diff_opt, "shape_branch.vq_cfg")))
            dd = _plain(_get(vq, "model.params.ddconfig"))
            L = int(dd["resolution"]) // 2 ** (len(dd["ch_mult"]) - 1)
            z_shape = (int(dd["z_channels"]), L, L, L)
            ddim_steps = 7 if int(_get(diff_opt, "misc.debug", 0)) == 1 else 100
            mp = _get(df, "model.params")
            self.unet3d = modules.UNet3DModel(**_plain(_get(df, "unet.params")), conditioning_key=_get(mp, "conditioning_key"),
                                              precision=precision, ddim_steps=ddim_steps, timesteps=int(_get(mp, "timesteps", 1000)),
                                              linear_start=float(_get(mp, "linear_start")), linear_end=float(_get(mp, "linear_end")))
            self.vqvae = modules.VQVAE(dd, int(_get(vq, "model.params.n_embed")), int(_get(vq, "model.params.embed_dim")),
                                       precision=precision, with_encoder=with_vq_encoder)
            vq_ckpt = _get(diff_opt, "shape_branch.vq_ckpt")
            if isinstance(vq_ckpt, str) and os.path.isfile(self._resolve(vq_ckpt, must_exist=False)):
                self.load_vqvae(self._resolve(vq_ckpt))                      # load_vqvae, model_utils.py:20-25
        cls = scene.Sg2BoxDiffModel if layout_only else scene.Sg2ScDiffModel
        bs = _get(diff_opt, "hyper.batch_size")                                  # diffusion_bs, EchoScene.py:76 / SGDiff.py:21
        kw = dict(replace_latent=replace_latent, box_dim=box_dim, size_dim=s_dim, translation_dim=t_dim,
                  diffusion_bs=16 if bs is None else int(bs))
        if layout_only:
            self.diff = cls(self.encoder, self.layout, **kw)
        else:
            self.diff = cls(self.encoder, self.layout, shape=self.unet3d, vqvae=self.vqvae, ddim_steps=ddim_steps,
                            uc_scale=float(_get(diff_opt, "shape_branch.uc_scale", 3.0)), z_shape=z_shape, **kw)

    def _resolve(self, path: str, must_exist: bool = True) -> str:
        if path is None:
            raise EchoError("shape_branch.df_cfg / vq_cfg missing from the configuration (echo2shape.py:59-60 asserts them)")
        cands = [path]
        if self._config_dir and not os.path.isabs(path):
            cands += [os.path.join(self._config_dir, path), os.path.join(self._config_dir, os.path.basename(path))]
        for c in cands:
            if os.path.exists(c):
                return c
        if must_exist:
            raise EchoError(f"cannot find {path!r} (looked in {cands}); pass config_dir=")
        return path

    # ---- weights -------------------------------------------------------------------------------------------------------
    def components(self):
        return [m for m in (self.encoder, self.unet1d, self.unet3d, self.vqvae) if m is not None]

    def load_vqvae(self, path_or_state_dict) -> None:
        """model_utils.load_vqvae: a file holding the VQVAE state_dict, or {'vqvae': state_dict, ...}."""
        sd = path_or_state_dict
        if isinstance(sd, (str, os.PathLike)):
            sd = torch.load(sd, map_location="cpu", weights_only=False)
        self.vqvae.load_state_dict(sd["vqvae"] if "vqvae" in sd else sd)

    def load_networks(self, exp, epoch, strict=True, restart_optim=False, load_shape_branch=True):
        """SGDiff.load_networks (SGDiff.py:49-84): <exp>/checkpoint/model<epoch>.pth."""
        path = os.path.join(exp, "checkpoint", "model{}.pth".format(epoch))
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        shape = load_shape_branch and self.unet3d is not None and "shape_df" in ckpt and "vqvae" in ckpt
        info = scene.load_reference_checkpoint(ckpt, encoder=self.encoder, unet1d=self.unet1d,
                                               unet3d=self.unet3d if shape else None, vqvae=self.vqvae if shape else None,
                                               strict=strict)
        if info["epoch"] is not None:
            self.epoch = info["epoch"]
        if info["counter"] is not None:
            self.counter = info["counter"]
        # what the checkpoint holds besides this facade's components -- parameters of reference sub-modules that never run on this
        # path (obj_embeddings_dc, pred_embeddings_dc, ...), the VQ-VAE encoder of a decode-only facade, the optimizer
        # state: kept so that save() writes a checkpoint the reference loads with strict=True
        owned = tuple(self.encoder.reference_prefixes()) + ("LayoutDiff.df.model.", "LayoutDiff.df.module.model.")
        self._foreign = {k: v for k, v in ckpt.items() if torch.is_tensor(v) and not k.startswith(owned)}
        self._foreign_opt = ckpt.get("opt")
        self._foreign_vqvae = {k: v for k, v in ckpt.get("vqvae", {}).items()} if "vqvae" in ckpt else None
        return info

    def to(self, device):
        for m in self.components():
            m.to(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        for m in self.components():
            m.eval()
        return self

    def train(self, mode: bool = True):
        """model.train() of scripts/train_3dfront.py:237: the BatchNorm1d layers switch to batch statistics (forward values only;
        the VQ-VAE stays in eval as in the reference)."""
        self.diff.train(mode)
        return self

    # ---- sampling (SGDiff.py:87-121) ---------------------------------------------------------------------------------------
    def sample_box_and_shape(self, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, gen_shape=False):
        return self.diff.sample_box_and_shape(dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, gen_shape=gen_shape)

    def sample_boxes_and_shape_with_changes(self, *args, gen_shape=False):
        return self.diff.sample_boxes_and_shape_with_changes(*args, gen_shape=gen_shape)

    def sample_boxes_and_shape_with_additions(self, *args, gen_shape=False):
        return self.diff.sample_boxes_and_shape_with_additions(*args, gen_shape=gen_shape)

    def sample_scenes(self, objs, triples, text_feat, rel_feat, obj_to_scene, gen_shape=False, x_T_per_scene=None):
        """A collated batch of scenes in one call (no counterpart in the reference, which samples scene by scene):
        scene.Sg2ScDiffModel.sample_scenes."""
        shape_dict, layout_dict, o2s = self.diff.sample_scenes(objs, triples, text_feat, rel_feat, obj_to_scene, gen_shape=gen_shape,
                                                               x_T_per_scene=x_T_per_scene)
        return {**shape_dict, **layout_dict, "obj_to_scene": o2s}

    def forward_mani(self, enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs, dec_objs_grained,
                     dec_triples, dec_boxes, dec_angles, dec_sdfs, encoded_dec_text_feat, encoded_dec_rel_feat, dec_objs_to_scene,
                     missing_nodes, manipulated_nodes):
        """SGDiff.forward_mani (SGDiff.py:32-47), same argument order -> (obj_selected, shape_loss, layout_loss, loss_dict) as
        forward VALUES under train(): nothing is recorded for loss.backward()."""
        if self.type_ == "echoscene":
            return self.diff.forward(enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs, dec_objs_grained,
                                     dec_triples, dec_boxes, encoded_dec_text_feat, encoded_dec_rel_feat, dec_objs_to_scene,
                                     missing_nodes, manipulated_nodes, dec_sdfs, dec_angles)
        return self.diff.forward(enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs, dec_triples, dec_boxes,
                                 encoded_dec_text_feat, encoded_dec_rel_feat, dec_objs_to_scene, missing_nodes, manipulated_nodes,
                                 dec_angles)

    def state_dict(self, epoch, counter, optimizer_state=None) -> dict:
        """Sg2ScDiffModel.state_dict(epoch, counter) / Sg2BoxDiffModel's (model/EchoScene.py:534-544): the checkpoint dictionary the
        reference's ``SGDiff.load_networks`` (SGDiff.py:49-84) reads -- flat keys of the scene model (this facade's encoder under its
        reference names, the layout denoiser under ``LayoutDiff.df.model.``), ``epoch``, ``counter``, ``opt``, and for the full model
        ``vqvae`` / ``shape_df`` (keys under ``diffusion_net.``, DiffusionUNet, network.py:17).  Tensors of reference sub-modules this
        path never runs, and the optimizer state, are passed through from the checkpoint this facade was loaded from
        (``load_networks``); without one they are absent and the reference needs ``strict=False`` / ``restart_optim=True``."""
        out = {k: v.detach().cpu() for k, v in getattr(self, "_foreign", {}).items()}
        out.update({k: v.detach().cpu() for k, v in self.encoder.state_dict().items()})
        out.update({"LayoutDiff.df.model." + k: v.detach().cpu() for k, v in self.unet1d.state_dict().items()})
        opt = optimizer_state if optimizer_state is not None else getattr(self, "_foreign_opt", None)
        out.update({"epoch": epoch, "counter": counter, "opt": opt if opt is not None else {}})
        if self.unet3d is not None:
            vq = dict(getattr(self, "_foreign_vqvae", None) or {})
            vq.update({k: v.detach().cpu() for k, v in self.vqvae.state_dict().items()})
            out["vqvae"] = vq
            out["shape_df"] = {"diffusion_net." + k: v.detach().cpu() for k, v in self.unet3d.state_dict().items()}
        return out

    def save(self, exp, outf, epoch, counter=None, optimizer_state=None):
        """SGDiff.save (SGDiff.py:123-129): <exp>/<outf>/model<epoch>.pth in the reference's checkpoint format."""
        os.makedirs(os.path.join(exp, outf), exist_ok=True)
        torch.save(self.state_dict(epoch, counter, optimizer_state), os.path.join(exp, outf, "model{}.pth".format(epoch)))
