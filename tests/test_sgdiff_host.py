"""CPU: echoscene_b200.sgdiff.SGDiff -- the facade of model/SGDiff.py built from YAML files with the reference's key structure
(config/full_mp.yaml, config/sdfusion-txt2shape_mp.yaml, config/vqvae_snet.yaml), its checkpoint loading, and that every
sampling call ends in the CUDA library (no CPU fallback)."""

import pytest
import torch
import yaml

from echoscene_b200 import _lib, arch, modules, scene, sgdiff
from oracle import cases

VOCAB = {"object_idx_to_name": ["_scene_"] + [f"c{i}" for i in range(35)], "pred_idx_to_name": ["in"] + [f"p{i}" for i in range(15)],
         "object_idx_to_name_grained": ["x"]}

MAIN = {
    "hyper": {"batch_size": 64, "device": "cuda", "lr_init": 1e-4},
    "layout_branch": {
        "model": "diffusion_scene_layout_ddpm", "angle_dim": 2, "denoiser": "unet1d", "relation_condition": True,
        "denoiser_kwargs": {"dims": 1, "in_channels": 8, "out_channels": 8, "model_channels": 512, "channel_mult": [1, 1, 1, 1],
                            "num_res_blocks": 2, "attention_resolutions": [4, 2], "num_heads": 8, "use_spatial_transformer": True,
                            "transformer_depth": 1, "conditioning_key": "crossattn", "concat_dim": 1280, "crossattn_dim": 1280,
                            "use_checkpoint": True, "enable_t_emb": True},
        "diffusion_kwargs": {"schedule_type": "linear", "beta_start": 0.0001, "beta_end": 0.02, "time_num": 1000,
                             "model_mean_type": "eps", "model_var_type": "fixedsmall", "loss_separate": True, "loss_iou": False,
                             "iou_type": "obb", "train_stats_file": None}},
    "shape_branch": {"model": "sdfusion-txt2shape_mp", "sampling": "greedy", "df_cfg": "../config/df.yaml", "ddim_steps": 100,
                     "ddim_eta": 0.0, "uc_scale": 3.0, "vq_model": "vqvae", "vq_cfg": "../config/vq.yaml",
                     "vq_ckpt": "./checkpoint/vqvae_threedfront_best.pth"},
    "misc": {"debug": 0, "seed": 111},
}
# the shape denoiser at a reduced width so that the CPU test stays light (the structure of the file is the reference's)
DF = {"model": {"params": {"linear_start": 0.00085, "linear_end": 0.012, "conditioning_key": "crossattn", "timesteps": 1000}},
      "unet": {"params": {"image_size": 16, "in_channels": 3, "out_channels": 3, "model_channels": 32, "num_res_blocks": 1,
                          "attention_resolutions": [2], "channel_mult": [1, 2], "num_heads": 8, "dims": 3,
                          "use_spatial_transformer": True, "transformer_depth": 1, "context_dim": 1280, "use_checkpoint": True,
                          "legacy": False, "messsage_passing": True, "enable_t_emb": True}}}
VQ = {"model": {"params": {"embed_dim": 3, "n_embed": 8192, "ddconfig": {
    "double_z": False, "z_channels": 3, "resolution": 64, "in_channels": 1, "out_ch": 1, "ch": 64, "ch_mult": [1, 2, 4],
    "num_res_blocks": 1, "attn_resolutions": [], "dropout": 0.0}}}}


@pytest.fixture()
def cfg_dir(tmp_path):
    d = tmp_path / "config"
    d.mkdir()
    for name, body in (("full_mp.yaml", MAIN), ("df.yaml", DF), ("vq.yaml", VQ)):
        (d / name).write_text(yaml.safe_dump(body))
    return d


def test_echolayout_from_yaml_and_checkpoint(cfg_dir, tmp_path):
    m = sgdiff.SGDiff("echolayout", str(cfg_dir / "full_mp.yaml"), VOCAB, residual=True)
    assert isinstance(m.diff, scene.Sg2BoxDiffModel) and m.unet3d is None and m.vqvae is None
    assert m.encoder.cfg.man_dc_preds and not m.encoder.with_rel_s and m.encoder.cfg.num_objs == 36
    assert m.layout.time_num == 1000 and m.diff.box_dim == 8
    assert list(m.unet1d.state_dict().keys()) == list(arch.unet1d_specs(cases.layout_cfg()).keys())
    # a checkpoint with the key structure Sg2BoxDiffModel.state_dict(epoch, counter) writes (EchoLayout.py:403-407)
    esd = cases.scene_box_state_dict()
    lsd = arch.make_state_dict(arch.unet1d_specs(cases.layout_cfg()), 9)
    ckpt = dict(esd)
    ckpt.update({"LayoutDiff.df.model." + k: v for k, v in lsd.items()})
    ckpt.update({"epoch": 3, "counter": 77, "opt": {}})
    (tmp_path / "exp" / "checkpoint").mkdir(parents=True)
    torch.save(ckpt, tmp_path / "exp" / "checkpoint" / "model3.pth")
    info = m.load_networks(str(tmp_path / "exp"), 3)
    assert (m.epoch, m.counter) == (3, 77) and info["loaded"] == {"encoder": len(esd), "unet1d": len(lsd)}
    assert torch.equal(m.unet1d.state_dict()["out.2.weight"], lsd["out.2.weight"])
    assert torch.equal(m.encoder.state_dict()["pred_embeddings_man_dc.weight"], esd["pred_embeddings_man_dc.weight"])
    # every sampling call ends in the CUDA library: CPU tensors are refused, nothing is computed on the host
    g, objs, text, rel = cases.scene_inputs()
    m.eval()
    with pytest.raises(_lib.EchoError, match="CUDA"):
        m.sample_box_and_shape(objs, g.triples, text, rel)
    # train(): batch-statistics forward values (the VQ-VAE stays in eval); the training forward ends in the CUDA library as well
    assert m.train() is m and m.encoder.training and m.unet1d.training
    boxes, angles = torch.zeros(len(objs), 6), torch.zeros(len(objs))
    with pytest.raises(_lib.EchoError, match="CUDA"):
        m.forward_mani(objs, g.triples, text, rel, objs, objs, g.triples, boxes, angles, None, text, rel,
                       torch.zeros(len(objs), dtype=torch.int64), [], [])
    with pytest.raises(_lib.EchoError, match="eval"):
        m.sample_box_and_shape(objs, g.triples, text, rel)
    assert not m.eval().encoder.training


def test_echoscene_from_yaml(cfg_dir, tmp_path):
    m = sgdiff.SGDiff("echoscene", yaml.safe_load((cfg_dir / "full_mp.yaml").read_text()), VOCAB, residual=True,
                      config_dir=str(cfg_dir))
    assert isinstance(m.diff, scene.Sg2ScDiffModel) and not isinstance(m.diff, scene.Sg2BoxDiffModel)
    assert m.encoder.with_rel_s and not m.encoder.cfg.man_dc_preds
    assert isinstance(m.unet3d, modules.UNet3DModel) and m.unet3d.ddim_steps == 100 and m.unet3d.linear_end == 0.012
    assert m.diff.z_shape == (3, 16, 16, 16) and m.diff.ddim_steps == 100 and m.diff.uc_scale == 3.0
    assert isinstance(m.vqvae, modules.VQVAE) and m.vqvae.cfg.n_embed == 8192
    vsd = arch.make_state_dict(arch.vqvae_decode_specs(cases.vqvae_cfg()), 4)
    torch.save({"vqvae": dict(vsd), "df": {}}, tmp_path / "vq.pth")       # the file format of EchoToShape.save (echo2shape.py:679-694)
    m.load_vqvae(str(tmp_path / "vq.pth"))
    assert torch.equal(m.vqvae.state_dict()["decoder.conv_in.bias"], vsd["decoder.conv_in.bias"])
    debug = dict(MAIN, misc={"debug": 1})
    assert sgdiff.SGDiff("echoscene", debug, VOCAB, config_dir=str(cfg_dir)).diff.ddim_steps == 7     # echo2shape.py:116-120
    with pytest.raises(_lib.EchoError, match="separated"):
        sgdiff.SGDiff("echoscene", MAIN, VOCAB, separated=True, config_dir=str(cfg_dir))
    with pytest.raises(_lib.EchoError, match="cannot find"):
        sgdiff.SGDiff("echoscene", MAIN, VOCAB, config_dir=str(tmp_path))
    with pytest.raises(AssertionError):
        sgdiff.SGDiff("other", MAIN, VOCAB)


def test_config_access_helpers():
    class NS:
        def __init__(self, **k):
            self.__dict__.update(k)
    cfg = {"a": {"b": NS(c=[1, 2], d=None)}}
    assert sgdiff._get(cfg, "a.b.c") == [1, 2] and sgdiff._get(cfg, "a.b.d", 5) == 5 and sgdiff._get(cfg, "a.x.y", "z") == "z"
    assert sgdiff._plain({"k": (1, 2), "n": {"m": 3}}) == {"k": [1, 2], "n": {"m": 3}}


import os  # noqa: E402

_REF_INSTALLED = os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "model"))


@pytest.mark.skipif(not _REF_INSTALLED, reason="baseline/_ref (the reference copy) is not installed")
def test_checkpoint_round_trip_through_the_unmodified_reference(tmp_path):
    """SGDiff.save / load_networks against the reference's own: the UNMODIFIED reference SGDiff('echolayout') (reduced layout width,
    CPU) writes a checkpoint with its own save(); the B200 facade loads it, saves it again; the two files hold the same dictionary
    (keys, tensors, epoch, counter, optimizer state) and the reference restores itself from the facade's file with strict=True."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import refbind_check as rb
    rb.setup_reference()
    cfg = rb.load_yaml(os.path.join(rb.REF, "config", "full_mp.yaml"))
    cfg.hyper.device, cfg.hyper.batch_size = "cpu", 16
    cfg.hyper.logs_dir = cfg.hyper.results_dir = str(tmp_path / "logs")
    cfg.layout_branch.denoiser_kwargs.model_channels = 64                      # keeps the files small; the format does not depend on it
    vocab = {"object_idx_to_name": ["_scene_"] + [f"c{i}" for i in range(35)], "pred_idx_to_name": ["in"] + [f"p{i}" for i in range(15)],
             "object_idx_to_name_grained": ["x"]}
    RefSGDiff = importlib.import_module("model.SGDiff").SGDiff
    torch.manual_seed(3)
    ref = RefSGDiff("echolayout", cfg, vocab, replace_latent=True, with_changes=True, residual=True, gconv_pooling="avg", with_angles=True,
                    clip=True, separated=False)
    ref.counter = 123
    (tmp_path / "exp" / "checkpoint").mkdir(parents=True)
    ref.save(str(tmp_path / "exp"), "checkpoint", 7, counter=123)              # the reference's own writer
    mine = sgdiff.SGDiff("echolayout", cfg, vocab, replace_latent=True, residual=True, clip=True, config_dir=os.path.join(rb.REF, "config"))
    info = mine.load_networks(str(tmp_path / "exp"), 7)
    assert (mine.epoch, mine.counter) == (7, 123) and info["loaded"]["unet1d"] > 0
    mine.save(str(tmp_path / "exp2"), "checkpoint", 7, counter=123)
    a = torch.load(tmp_path / "exp" / "checkpoint" / "model7.pth", map_location="cpu", weights_only=False)
    b = torch.load(tmp_path / "exp2" / "checkpoint" / "model7.pth", map_location="cpu", weights_only=False)
    assert set(a) == set(b)
    for k, v in a.items():
        if torch.is_tensor(v):
            assert torch.equal(v, b[k]), k
    assert (b["epoch"], b["counter"]) == (7, 123) and b["opt"]["param_groups"] == a["opt"]["param_groups"]
    want = {k: v.clone() for k, v in torch.nn.Module.state_dict(ref.diff).items()}
    with torch.no_grad():
        for p in ref.diff.parameters():
            p.zero_()
    ref.load_networks(str(tmp_path / "exp2"), 7, strict=True, restart_optim=False)     # the reference's own reader, strict
    for k, v in torch.nn.Module.state_dict(ref.diff).items():
        assert torch.equal(v, want[k]), k
    # a facade that was never loaded writes its own components only (the reference then needs strict=False / restart_optim=True)
    fresh = sgdiff.SGDiff("echolayout", cfg, vocab, replace_latent=True, residual=True, clip=True, config_dir=os.path.join(rb.REF, "config"))
    sd = fresh.state_dict(1, 2)
    assert sd["opt"] == {} and sd["epoch"] == 1 and "LayoutDiff.df.model.out.2.weight" in sd and "vqvae" not in sd
