"""Scene-level glue: the host side of ``Sg2ScDiffModel.sample`` / ``sample_with_changes`` / ``sample_with_additions``
(model/EchoScene.py:388-532), of the training forward ``Sg2ScDiffModel.forward`` as loss values (:328-386, with ``select_sdfs``
:289-319) and of the ``SGDiff`` facade over them (model/SGDiff.py:32-47, 87-121).

The arithmetic lives behind the C ABI (``SceneEncoder`` -> echo_scene_*, ``DiffusionPoint`` -> echo_layout_step,
``DDIMSampler`` -> echo_shape_step, ``VQVAE`` -> echo_vqvae_decode); this file is the order of calls and the row
bookkeeping the reference does around them (change flags, inserted nodes, which latent rows are replaced, ``keep``),
restated without the per-node host->device copies of the reference (one upload of the change flags per scene instead of
N, EchoScene.py:432-439).  It holds no parameters of its own: each component loads its slice of the reference's
checkpoints (``SceneEncoder.load_reference_state_dict``, ``UNet1DModel`` / ``UNet3DModel`` / ``VQVAE`` state_dicts).

Random draws follow the reference's streams: change flags from ``np.random.normal`` in ascending node order
(EchoScene.py:437, 494), the layout chain from ``torch.randn`` (diffusion_ddpm.py:302, 336), ONE shape noise draw repeated
for every object (echo2shape.py:507-510).  The reference also reseeds torch's global generator from the wall clock before
that draw (echo2shape.py:502); that side effect is reproduced only with ``reference_rng=True``.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from ._lib import EchoError


def change_flags(n_nodes: int, marked: Sequence[int], dim: int, device) -> torch.Tensor:
    """(N, dim) f32: zeros, and one ``np.random.normal(0, 1, dim)`` row per node of ``marked``, drawn in ascending node
    order as the reference's loop does (EchoScene.py:432-439, 489-496).  One host->device copy."""
    marked = set(int(i) for i in marked)
    host = np.zeros((n_nodes, dim), dtype=np.float64)
    for i in range(n_nodes):
        if i in marked:
            host[i] = np.random.normal(0, 1, dim)
    return torch.from_numpy(host).float().to(device)


def insert_zero_rows(latent: torch.Tensor, missing_nodes: Sequence[int]):
    """The "append zero nodes" loop of sample_with_additions (EchoScene.py:478-486): for the i-th missing node a zero row is
    inserted at ``missing_nodes[i] + i``, sequentially.  -> (latent with the rows inserted, nodes_added)."""
    order = list(range(latent.shape[0]))
    nodes_added = []
    for i, m in enumerate(missing_nodes):
        ad_id = int(m) + i
        nodes_added.append(ad_id)
        order.insert(ad_id, -1)          # same clamping as latent[:ad_id] ++ zeros ++ latent[ad_id:]
    idx = torch.tensor(order, dtype=torch.int64, device=latent.device)
    out = torch.zeros(len(order), latent.shape[1], dtype=latent.dtype, device=latent.device)
    old = idx >= 0
    out[old] = latent[idx[old]]
    return out, nodes_added


def replace_rows(base: torch.Tensor, new: torch.Tensor, touched: Sequence[int]) -> torch.Tensor:
    """"take original nodes when untouched" (EchoScene.py:444-450, 501-507): rows ``touched`` of ``base`` come from ``new``."""
    out = base.clone()
    t = sorted(set(int(i) for i in touched))
    if t:
        if t[0] < 0 or t[-1] >= base.shape[0]:
            # the reference's slice-and-cat would silently change the row count here; refuse instead of mis-shaping
            raise IndexError(f"touched nodes {t} outside [0, {base.shape[0]})")
        ti = torch.tensor(t, dtype=torch.int64, device=base.device)
        out[ti] = new[ti]
    return out


def keep_mask(n: int, dropped: Sequence[int], device) -> torch.Tensor:
    """(n, 1) f32: 1 for nodes kept from the input scene, 0 for manipulated / added ones (EchoScene.py:466-472)."""
    dropped = set(int(i) for i in dropped)
    k = np.asarray([0 if i in dropped else 1 for i in range(n)]).reshape(-1, 1)
    return torch.from_numpy(k).float().to(device)


def _strip(sd: dict, prefixes: Sequence[str]) -> dict:
    """Entries of ``sd`` under the first prefix of ``prefixes`` that occurs, with that prefix removed."""
    for p in prefixes:
        sub = {k[len(p):]: v for k, v in sd.items() if isinstance(k, str) and k.startswith(p)}
        if sub:
            return sub
    return {}


def load_reference_checkpoint(ckpt, encoder=None, unet1d=None, unet3d=None, vqvae=None, strict: bool = True) -> dict:
    """Distributes a checkpoint written by the reference (``SGDiff.save`` -> ``Sg2ScDiffModel.state_dict(epoch, counter)``,
    model/SGDiff.py:123-129, model/EchoScene.py:534-544) over the B200 components, the way ``SGDiff.load_networks``
    (model/SGDiff.py:49-84) distributes it over the reference's modules:

      flat keys obj_embeddings_ec.* ... rel_s_mlp.*      -> encoder  (modules.SceneEncoder)
      flat keys LayoutDiff.df.model.*                     -> unet1d   (modules.UNet1DModel; DiffusionPoint.model, diffusion_ddpm.py:562)
      ckpt['shape_df']  keys diffusion_net.*              -> unet3d   (modules.UNet3DModel; DiffusionUNet.diffusion_net, network.py:17)
      ckpt['vqvae']                                       -> vqvae    (modules.VQVAE)
      ckpt['epoch'], ckpt['counter']                      -> returned; ckpt['opt'] (AdamW state) is not used when sampling

    ``ckpt`` is a path or the loaded dict.  A ``module.`` prefix left by DistributedDataParallel (echo2shape.py:128-139) is
    accepted.  Components passed as None are skipped; a component whose slice is absent raises (KeyError) unless it is the
    shape branch of a layout-only checkpoint and ``strict`` is False (load_networks prints and carries on there, :66-67)."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    info = {"epoch": ckpt.get("epoch"), "counter": ckpt.get("counter"), "loaded": {}}
    flat = {k: v for k, v in ckpt.items() if torch.is_tensor(v)}
    if encoder is not None:
        encoder.load_reference_state_dict(flat, strict=strict)
        info["loaded"]["encoder"] = sum(1 for k in flat if k.startswith(encoder.reference_prefixes()))
    if unet1d is not None:
        sub = _strip(flat, ("LayoutDiff.df.model.", "LayoutDiff.df.module.model."))
        if not sub:
            raise KeyError("checkpoint has no LayoutDiff.df.model.* entries (not written by Sg2ScDiffModel / Sg2BoxDiffModel?)")
        unet1d.load_state_dict(sub, strict=strict)
        info["loaded"]["unet1d"] = len(sub)
    for name, comp, key, prefixes in (("unet3d", unet3d, "shape_df", ("diffusion_net.", "module.diffusion_net.")),
                                      ("vqvae", vqvae, "vqvae", ("module.",))):
        if comp is None:
            continue
        if key not in ckpt:
            if strict:
                raise KeyError(f"checkpoint has no '{key}' entry (layout-only checkpoint?)")
            continue
        sd = ckpt[key]
        sub = _strip(sd, prefixes) or dict(sd)
        comp.load_state_dict(sub, strict=strict)
        info["loaded"][name] = len(sub)
    return info


def split_by_scene(out: Dict[str, Optional[torch.Tensor]], obj_to_scene: torch.Tensor):
    """Per-scene views of the row-wise results of ``sample_scenes``: a list (one dict per scene) of the rows of every tensor
    whose first dimension is the node count."""
    o2s = obj_to_scene.detach().cpu().to(torch.int64)
    n_scenes = int(o2s.max()) + 1 if o2s.numel() else 0
    idx = [torch.nonzero(o2s == k).flatten() for k in range(n_scenes)]
    return [{k: (None if v is None else v[i.to(v.device)]) for k, v in out.items()} for i in idx]


class Sg2ScDiffModel:
    """Sampling surface of the reference's scene model on the B200 components.

    encoder : modules.SceneEncoder           (init_encoder / manipulate / rel_s)
    layout  : samplers.DiffusionPoint        (gen_samples_sg)
    shape   : modules.UNet3DModel or None    (iterated by samplers.DDIMSampler); None = layout only
    vqvae   : modules.VQVAE or None          (decode_no_quant); None = ``shapes`` are the (N,3,16,16,16) latents
    """

    def __init__(self, encoder, layout, shape=None, vqvae=None, ddim_steps: int = 100, uc_scale: float = 3.0,
                 z_shape=(3, 16, 16, 16), replace_latent: bool = False, box_dim: int = 8, size_dim: int = 3,
                 translation_dim: int = 3, reference_rng: bool = False, ddim_sampler_cls: Optional[Callable] = None,
                 diffusion_bs: int = 16):
        self.encoder, self.layout, self.shape, self.vqvae = encoder, layout, shape, vqvae
        self.ddim_steps, self.uc_scale, self.z_shape = int(ddim_steps), uc_scale, tuple(z_shape)
        self.replace_all_latent = replace_latent                      # EchoScene.py:27
        self.box_dim, self.size_dim, self.translation_dim = box_dim, size_dim, translation_dim
        self.embedding_dim = encoder.embedding_dim
        self.out_dim_ini_encoder = encoder.out_dim_ini_encoder
        self.reference_rng = reference_rng
        self._ddim_cls = ddim_sampler_cls
        self.diffusion_bs = int(diffusion_bs)                          # EchoScene.py:76 (SGDiff.py:21 passes 16)
        self._shape_tables = self._layout_tables = None

    # ---- the two chains -------------------------------------------------------------------------------------------
    def generate_layout(self, triples, obj_embed, relation_cond) -> Dict[str, torch.Tensor]:
        """prepare_boxes + EchoToLayout.generate_layout_sg (EchoScene.py:321-326, echo2layout.py:100-126)."""
        n = obj_embed.shape[0]
        samples = self.layout.gen_samples_sg((n, self.box_dim), obj_embed.device, obj_embed, triples,
                                             condition=relation_cond, clip_denoised=False)
        s, t = self.size_dim, self.size_dim + self.translation_dim
        return {"sizes": samples[:, 0:s].contiguous(), "translations": samples[:, s:t].contiguous(),
                "angles": samples[:, t:self.box_dim].contiguous()}

    def rel2shape(self, triples, c_s, uc_s, x_T: Optional[torch.Tensor] = None):
        """EchoToShape.rel2shape (echo2shape.py:484-525): one noise draw repeated per object, DDIM chain, VQ-VAE decode."""
        if self.shape is None:
            raise EchoError("gen_shape=True needs the shape branch (a UNet3DModel)")
        B = c_s.shape[0]
        if x_T is None:
            if self.reference_rng:
                torch.manual_seed(int(time.time()))                   # echo2shape.py:502
            single = torch.randn((1,) + self.z_shape, device=c_s.device)
            x_T = single.repeat(B, 1, 1, 1, 1)
        cls = self._ddim_cls
        if cls is None:
            from .samplers import DDIMSampler as cls
        samples, _ = cls(self.shape).sample(S=self.ddim_steps, batch_size=B, shape=self.z_shape, conditioning=c_s, x_T=x_T,
                                            verbose=False, unconditional_guidance_scale=self.uc_scale,
                                            unconditional_conditioning=uc_s, triplet=triples, eta=0.0)
        return samples if self.vqvae is None else self.vqvae.decode_no_quant(samples)

    def _chains(self, dec_triples, obj_embed_, latent, gen_shape, x_T):
        layout_dict = self.generate_layout(dec_triples, obj_embed_, latent)
        gen_sdf = None
        if gen_shape:
            uc_s = self.encoder.rel_s(obj_embed_).unsqueeze(1)        # embedding + CLIP, EchoScene.py:405-406
            c_s = self.encoder.rel_s(latent).unsqueeze(1)
            gen_sdf = self.rel2shape(dec_triples, c_s, uc_s, x_T)
        return {"shapes": gen_sdf}, layout_dict

    # ---- Sg2ScDiffModel ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, dec_objs, dec_triplets, dec_text_feat, dec_rel_feat, gen_shape=False, x_T=None):
        """EchoScene.py:388-420 -> ({'shapes': sdf or None}, {'sizes', 'translations', 'angles'})."""
        enc = self.encoder.encode(dec_objs, dec_triplets, dec_text_feat, dec_rel_feat, shape_cond=gen_shape)
        layout_dict = self.generate_layout(dec_triplets, enc["obj_embed"], enc["latent"])
        gen_sdf = self.rel2shape(dec_triplets, enc["c_s"], enc["uc_s"], x_T) if gen_shape else None
        return {"shapes": gen_sdf}, layout_dict

    @torch.no_grad()
    def sample_scenes(self, objs, triples, text_feat, rel_feat, obj_to_scene, gen_shape=False, x_T_per_scene=None):
        """Many scenes in one call (BASELINE config 4): the inputs are a collated batch as dataset collate_fn builds it
        (threedfront_dataset.py:618-743: node / triple tensors concatenated, triple indices offset per scene, ``obj_to_scene``
        (N,) int64).  The batched graph is block-diagonal, so encoders and both chains run over all nodes at once -- objects
        are the batch dimension of both denoisers and the echo GCN never crosses a scene.  The reference samples scene by
        scene with ONE shape-noise draw per call (echo2shape.py:507-510); here every scene gets its own draw
        (``x_T_per_scene`` (S, 3, 16, 16, 16) or fresh ``randn``), repeated over that scene's objects.
        -> ({'shapes'}, layout_dict, obj_to_scene): rows in node order; ``split_by_scene`` cuts them per scene."""
        o2s = obj_to_scene.to(torch.int64)
        n_scenes = int(o2s.max()) + 1 if o2s.numel() else 0
        enc = self.encoder.encode(objs, triples, text_feat, rel_feat, shape_cond=gen_shape)
        layout_dict = self.generate_layout(triples, enc["obj_embed"], enc["latent"])
        gen_sdf = None
        if gen_shape:
            if x_T_per_scene is None:
                if self.reference_rng:
                    torch.manual_seed(int(time.time()))
                x_T_per_scene = torch.randn((n_scenes,) + self.z_shape, device=enc["c_s"].device)
            if tuple(x_T_per_scene.shape) != (n_scenes,) + self.z_shape:
                raise EchoError(f"x_T_per_scene must be {(n_scenes,) + self.z_shape}, got {tuple(x_T_per_scene.shape)}")
            x_T = x_T_per_scene.to(enc["c_s"].device)[o2s.to(enc["c_s"].device)].contiguous()
            gen_sdf = self.rel2shape(triples, enc["c_s"], enc["uc_s"], x_T)
        return {"shapes": gen_sdf}, layout_dict, o2s

    @torch.no_grad()
    def sample_with_changes(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triplets, dec_text_feat,
                            dec_rel_feat, manipulated_nodes, gen_shape=False, x_T=None):
        """EchoScene.py:422-472 -> (keep (N,1), {'shapes'}, layout_dict)."""
        e = self.encoder
        _, _, latent_obj, _ = e.init_encoder(enc_objs, enc_triples, enc_text_feat, enc_rel_feat)
        change = change_flags(latent_obj.shape[0], manipulated_nodes, self.embedding_dim, latent_obj.device)
        latent_, _, obj_embed_, _ = e.manipulate(torch.cat([latent_obj, change], dim=1), dec_objs, dec_triplets, dec_text_feat,
                                                 dec_rel_feat)
        latent = latent_ if self.replace_all_latent else replace_rows(latent_obj, latent_, manipulated_nodes)
        shape_dict, layout_dict = self._chains(dec_triplets, obj_embed_, latent, gen_shape, x_T)
        keep = keep_mask(len(layout_dict["translations"]), manipulated_nodes, latent.device)
        return keep, shape_dict, layout_dict

    @torch.no_grad()
    def sample_with_additions(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triplets, dec_text_feat,
                              dec_rel_feat, missing_nodes, gen_shape=False, x_T=None):
        """EchoScene.py:474-532 -> (keep (N,1), {'shapes'}, layout_dict).  As in the reference, the change flags are drawn
        for the indices in ``missing_nodes`` (:492) while rows are replaced / dropped from ``keep`` at ``nodes_added`` =
        missing_nodes[i] + i (:480, :503, :528)."""
        e = self.encoder
        _, _, latent_obj, _ = e.init_encoder(enc_objs, enc_triples, enc_text_feat, enc_rel_feat)
        latent_obj, nodes_added = insert_zero_rows(latent_obj, missing_nodes)
        change = change_flags(latent_obj.shape[0], missing_nodes, self.embedding_dim, latent_obj.device)
        latent_, _, obj_embed_, _ = e.manipulate(torch.cat([latent_obj, change], dim=1), dec_objs, dec_triplets, dec_text_feat,
                                                 dec_rel_feat)
        latent = latent_ if self.replace_all_latent else replace_rows(latent_obj, latent_, nodes_added)
        shape_dict, layout_dict = self._chains(dec_triplets, obj_embed_, latent, gen_shape, x_T)
        keep = keep_mask(len(layout_dict["translations"]), nodes_added, latent.device)
        return keep, shape_dict, layout_dict

    # ---- the training forward (forward VALUES) --------------------------------------------------------------
    def train(self, mode: bool = True):
        """model.train() / model.eval() of the reference (scripts/train_3dfront.py:237): BatchNorm1d of the GCN MLPs and rel_s_mlp
        switch to batch statistics.  The VQ-VAE always stays in eval (echo2shape.py:334 switches only ``df``)."""
        self.encoder.train(mode)
        self.layout.train(mode)
        if self.shape is not None:
            self.shape.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def select_sdfs(self, dec_objs_to_scene, obj_cats, triples, sdfs, s_feat_ucon, s_feat_con, sample_type: str = "greedy"):
        """Sg2ScDiffModel.select_sdfs, sample_type 'greedy' (EchoScene.py:289-319; the only option with message passing in the shape
        branch, :104): whole scenes in order until the next one no longer fits ``diffusion_bs`` objects, and the triples among the
        selected nodes.  -> (obj_cat_selected, {'sdf', 'uc_s', 'c_s', 'scene_ids', 'triples'})."""
        if sample_type != "greedy":
            raise EchoError(f"select_sdfs: sampling='{sample_type}' is outside the hot path (greedy is asserted with message passing, "
                            "EchoScene.py:104)")
        o2s = np.asarray(dec_objs_to_scene.detach().cpu() if torch.is_tensor(dec_objs_to_scene) else dec_objs_to_scene)
        num = 0
        for i in np.unique(o2s):
            ids = np.where(o2s == i)[0]
            if self.diffusion_bs - num < len(ids):
                break
            num += len(ids)
        if num == 0:
            raise EchoError(f"select_sdfs: the first scene has more than diffusion_bs = {self.diffusion_bs} objects (the reference "
                            "fails in torch.cat of an empty list here)")
        # scenes are contiguous node ranges in a collated batch (threedfront_dataset.py collate_fn), which the reference's
        # `triples[:, 0] < num` mask relies on as well
        sel = np.concatenate([np.where(o2s == i)[0] for i in np.unique(o2s)])[:num]
        if not np.array_equal(sel, np.arange(num)):
            raise EchoError("select_sdfs: obj_to_scene is not sorted by scene (not a collated batch)")
        mask = (triples[:, 0] < num) & (triples[:, 2] < num)
        triples_selected = triples[mask]
        if len(triples_selected) == 0:
            raise EchoError("select_sdfs: no triple among the selected objects (the reference passes triples=None to the denoiser here "
                            "and fails inside it)")
        bs = self.diffusion_bs
        return obj_cats[:num][:bs], {"sdf": sdfs[:num][:bs], "uc_s": s_feat_ucon[:num][:bs], "c_s": s_feat_con[:num][:bs],
                                      "scene_ids": o2s[:num][:bs], "triples": triples_selected}

    @torch.no_grad()
    def shape_loss(self, diff_dict):
        """EchoToShape.set_input + forward (echo2shape.py:229-241, 334-366): VQ-VAE encode without quantisation, one timestep per
        object, q_sample, the denoiser under train(), the eps losses.  Draw order: torch.randint (t), torch.randn_like (noise)."""
        from . import train as T
        if self.shape is None or self.vqvae is None:
            raise EchoError("shape_loss needs the shape branch and a VQVAE built with with_encoder=True")
        if self._shape_tables is None:
            u = self.shape
            self._shape_tables = {k: v.to(diff_dict["sdf"].device) for k, v in
                                  T.shape_train_tables(u.timesteps_total, u.linear_start, u.linear_end).items()}
        tb = self._shape_tables
        self.shape.train(True)                                                        # switch_train, echo2shape.py:336
        z = self.vqvae.encode_no_quant(diff_dict["sdf"])
        t = torch.randint(0, self.shape.timesteps_total, (z.shape[0],), device=z.device).long()
        noise = torch.randn_like(z)
        x_noisy = T.q_sample(z, t, noise, tb["sqrt_alphas_cumprod"], tb["sqrt_one_minus_alphas_cumprod"])
        out = self.shape(x_noisy, diff_dict["uc_s"], diff_dict["triples"], t, context=diff_dict["c_s"])
        loss, loss_dict = T.shape_diffusion_loss(out, noise, t, tb["logvar"], tb["lvlb_weights"])
        return loss, loss_dict

    @torch.no_grad()
    def layout_loss(self, diff_dict):
        """EchoToLayout.set_input + forward -> DiffusionPoint.get_loss_iter -> GaussianDiffusion.p_losses (echo2layout.py:57-99,
        diffusion_ddpm.py:597-608, 479-507): one timestep per SCENE, angle -> (sin, cos), q_sample, the denoiser under train(),
        diffusion_loss with loss_iou = False.  Draw order: torch.randint (t per scene), torch.randn (noise)."""
        from . import train as T
        box, o2s = diff_dict["box"], diff_dict["obj_id_to_scene"]
        if self._layout_tables is None:
            m = self.layout.model
            self._layout_tables = tuple(v.to(box.device) for v in T.layout_train_tables(m.time_num, m.beta_start, m.beta_end))
        self.layout.train(True)                                                       # echo2layout.py:80
        o2s = np.asarray(o2s.detach().cpu() if torch.is_tensor(o2s) else o2s)
        unique_scenes, inv_idx = np.unique(o2s, return_inverse=True)
        t = torch.randint(0, self.layout.time_num, size=unique_scenes.shape, device=box.device)
        t = t[torch.from_numpy(inv_idx).to(box.device)]
        D = box.shape[1]
        data_start = torch.cat((box[:, :D - 1], torch.sin(box[:, D - 1:D]), torch.cos(box[:, D - 1:D])), dim=-1).float()
        noise = torch.randn(data_start.shape, dtype=data_start.dtype, device=data_start.device)
        data_t = T.q_sample(data_start, t, noise, *self._layout_tables)
        out = self.layout._denoise(data_t, diff_dict["uc_b"], diff_dict["preds"], t, diff_dict["c_b"])
        return T.layout_diffusion_loss(out, noise, self.size_dim, self.translation_dim, self.box_dim - self.size_dim - self.translation_dim)

    @torch.no_grad()
    def forward(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_objs_grained, dec_triples, dec_boxes,
                dec_text_feat, dec_rel_feat, dec_objs_to_scene, missing_nodes, manipulated_nodes, dec_sdfs, dec_angles):
        """Sg2ScDiffModel.forward (EchoScene.py:328-386), the forward of one training iteration (scripts/train_3dfront.py:239-241,
        model.forward_mani) -> (obj_selected, Shape_loss, Layout_loss, loss_dict) as VALUES: no autograd tape (the backward pass is
        not part of this round).  Same argument order, same random draws in the same order (np.random.normal change
        flags, then the shape branch's randint / randn_like, then the layout branch's randint / randn), so a seeded run of the
        reference produces the same losses."""
        if not (self.encoder.training and self.layout.training and (self.shape is None or self.shape.training)):
            raise EchoError("forward is the training forward: call .train() first (sampling entry points are sample*/eval())")
        e = self.encoder
        _, _, latent_obj, _ = e.init_encoder(enc_objs, enc_triples, enc_text_feat, enc_rel_feat)
        latent_obj, nodes_added = insert_zero_rows(latent_obj, missing_nodes)
        marked = list(nodes_added) + [int(i) for i in manipulated_nodes]
        change = change_flags(latent_obj.shape[0], marked, self.embedding_dim, latent_obj.device)
        latent_, _, obj_embed_, _ = e.manipulate(torch.cat([latent_obj, change], dim=1), dec_objs, dec_triples, dec_text_feat,
                                                 dec_rel_feat)
        latent = latent_ if self.replace_all_latent else replace_rows(latent_obj, latent_, marked)
        obj_selected, shape_loss, loss_dict = None, None, {}
        if self.shape is not None:
            uc_s = e.rel_s(obj_embed_).unsqueeze(1)
            c_s = e.rel_s(latent).unsqueeze(1)
            obj_selected, shape_dict = self.select_sdfs(dec_objs_to_scene, dec_objs, dec_triples, dec_sdfs, uc_s, c_s)
            shape_loss, d = self.shape_loss(shape_dict)
            loss_dict.update(d)
        boxes = torch.cat((dec_boxes, dec_angles.reshape(-1, 1)), dim=-1)           # prepare_boxes, EchoScene.py:321-326
        layout_loss, d = self.layout_loss({"preds": dec_triples, "box": boxes, "uc_b": obj_embed_, "c_b": latent,
                                           "obj_id_to_scene": dec_objs_to_scene})
        loss_dict.update(d)
        return obj_selected, shape_loss, layout_loss, loss_dict

    # ---- SGDiff facade (model/SGDiff.py:87-121, type_ == 'echoscene') ----------------------------------------------
    def sample_box_and_shape(self, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, gen_shape=False):
        shape_dict, layout_dict = self.sample(dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat,
                                              gen_shape=gen_shape)
        return {**shape_dict, **layout_dict}

    def sample_boxes_and_shape_with_changes(self, enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                            dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, manipulated_nodes,
                                            gen_shape=False):
        keep, shape_dict, layout_dict = self.sample_with_changes(enc_objs, enc_triples, encoded_enc_text_feat,
                                                                 encoded_enc_rel_feat, dec_objs, dec_triples,
                                                                 encoded_dec_text_feat, encoded_dec_rel_feat,
                                                                 manipulated_nodes, gen_shape=gen_shape)
        return keep, {**shape_dict, **layout_dict}

    def sample_boxes_and_shape_with_additions(self, enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                              dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, missing_nodes,
                                              gen_shape=False):
        keep, shape_dict, layout_dict = self.sample_with_additions(enc_objs, enc_triples, encoded_enc_text_feat,
                                                                   encoded_enc_rel_feat, dec_objs, dec_triples,
                                                                   encoded_dec_text_feat, encoded_dec_rel_feat,
                                                                   missing_nodes, gen_shape=gen_shape)
        return keep, {**shape_dict, **layout_dict}


class Sg2BoxDiffModel(Sg2ScDiffModel):
    """Sampling surface of the layout-only model (model/EchoLayout.py:291-401; SGDiff type_ == 'echolayout').  Same encoder
    stage and layout chain; two behaviours of the reference differ from Sg2ScDiffModel and are kept: the additions draw
    their change flags at ``nodes_added`` (EchoLayout.py:369-374, not at ``missing_nodes``), and ``sampleBoxes_with_additions``
    returns ``keep`` as a plain Python list (:393-401)."""

    def __init__(self, encoder, layout, **kw):
        super().__init__(encoder, layout, shape=None, vqvae=None, **kw)

    @torch.no_grad()
    def forward(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triples, dec_boxes, dec_text_feat,
                dec_rel_feat, dec_objs_to_scene, missing_nodes, manipulated_nodes, dec_angles):
        """Sg2BoxDiffModel.forward (EchoLayout.py:247-289) -> (None, 0, Layout_loss, loss_dict): forward values, see
        Sg2ScDiffModel.forward."""
        _, _, layout_loss, loss_dict = super().forward(enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, None, dec_triples,
                                                       dec_boxes, dec_text_feat, dec_rel_feat, dec_objs_to_scene, missing_nodes,
                                                       manipulated_nodes, None, dec_angles)
        return None, 0, layout_loss, loss_dict

    @torch.no_grad()
    def sampleBoxes(self, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat):
        return self.sample(dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, gen_shape=False)[1]

    @torch.no_grad()
    def sampleBoxes_with_changes(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triples, dec_text_feat,
                                 dec_rel_feat, manipulated_nodes):
        keep, _, layout_dict = self.sample_with_changes(enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triples,
                                                        dec_text_feat, dec_rel_feat, manipulated_nodes, gen_shape=False)
        return keep, layout_dict

    @torch.no_grad()
    def sampleBoxes_with_additions(self, enc_objs, enc_triples, enc_text_feat, enc_rel_feat, dec_objs, dec_triples, dec_text_feat,
                                   dec_rel_feat, missing_nodes):
        e = self.encoder
        _, _, latent_obj, _ = e.init_encoder(enc_objs, enc_triples, enc_text_feat, enc_rel_feat)
        latent_obj, nodes_added = insert_zero_rows(latent_obj, missing_nodes)
        change = change_flags(latent_obj.shape[0], nodes_added, self.embedding_dim, latent_obj.device)
        latent_, _, obj_embed_, _ = e.manipulate(torch.cat([latent_obj, change], dim=1), dec_objs, dec_triples, dec_text_feat,
                                                 dec_rel_feat)
        latent = latent_ if self.replace_all_latent else replace_rows(latent_obj, latent_, nodes_added)
        layout_dict = self.generate_layout(dec_triples, obj_embed_, latent)
        added = set(nodes_added)
        keep = [0 if i in added else 1 for i in range(len(layout_dict["translations"]))]
        return keep, layout_dict

    # ---- SGDiff facade, type_ == 'echolayout' (model/SGDiff.py:87-121) ------------------------------------------------
    def sample_box_and_shape(self, dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat, gen_shape=False):
        return self.sampleBoxes(dec_objs, dec_triplets, encoded_dec_text_feat, encoded_dec_rel_feat)

    def sample_boxes_and_shape_with_changes(self, enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                            dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, manipulated_nodes,
                                            gen_shape=False):
        return self.sampleBoxes_with_changes(enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                             dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, manipulated_nodes)

    def sample_boxes_and_shape_with_additions(self, enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                              dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, missing_nodes,
                                              gen_shape=False):
        # the reference's facade drops `keep` on this branch (SGDiff.py:114-116)
        return self.sampleBoxes_with_additions(enc_objs, enc_triples, encoded_enc_text_feat, encoded_enc_rel_feat, dec_objs,
                                               dec_triples, encoded_dec_text_feat, encoded_dec_rel_feat, missing_nodes)[1]
