"""CPU: host glue of the scene-level sampling surface (echoscene_b200/scene.py) against what the reference's
own Sg2ScDiffModel.sample / sample_with_changes / sample_with_additions hand to the two diffusion branches
(tests/golden/scene_glue.pt, recorded by oracle/gen_golden_scene_glue.py), and the SceneEncoder module surface.

The encoder arithmetic is supplied here by the ORACLE (test infrastructure) behind the SceneEncoder method names, so that
what is tested is the glue: change flags and their np.random stream, inserted zero rows, replaced latent rows, the
rel_s_mlp inputs, `keep`.  The CUDA encoder itself is tested in tests/test_zz_scene_gpu.py."""
import os

import numpy as np
import pytest
import torch

from echoscene_b200 import _lib, arch, modules, scene
from oracle import cases, echoscene_oracle as orc
from oracle.scene_encoder import OracleSceneEncoder

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


OracleEncoder = OracleSceneEncoder   # SceneEncoder's method surface on the CPU oracle


class RecLayout:
    def gen_samples_sg(self, shape, device, obj_embed, triples=None, condition=None, clip_denoised=True, **kw):
        assert clip_denoised is False
        self.seen = {"uc_b": obj_embed, "c_b": condition, "preds": triples}
        return torch.arange(shape[0] * shape[1], dtype=torch.float32).reshape(shape)


class RecDDIM:
    seen = None

    def __init__(self, model):
        self.model = model

    def sample(self, S, batch_size, shape, conditioning=None, x_T=None, unconditional_conditioning=None, triplet=None, eta=0.,
               **kw):
        RecDDIM.seen = {"c_s": conditioning, "uc_s": unconditional_conditioning, "triples": triplet, "x_T": x_T, "S": S}
        assert x_T.shape == (batch_size,) + tuple(shape) and eta == 0.0
        return x_T * 0.5, {}


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "scene_glue.pt"))


@pytest.mark.parametrize("name,fn,replace", [c for c in cases.SCENE_GLUE_CASES if c[0].startswith("box_")])
def test_layout_only_glue_matches_reference_methods(gold, name, fn, replace):
    """Sg2BoxDiffModel.sampleBoxes* (model/EchoLayout.py:291-401)."""
    G = gold[name]
    lay = RecLayout()
    m = scene.Sg2BoxDiffModel(OracleEncoder(box=True), lay, replace_latent=replace)
    args, marked = cases.scene_glue_inputs(name)
    np.random.seed(cases.SCENE_GLUE_NP_SEED)
    if fn == "sampleBoxes":
        layout_dict = m.sampleBoxes(*args)
        assert m.sample_box_and_shape(*args).keys() == layout_dict.keys() == {"sizes", "translations", "angles"}
    else:
        keep, layout_dict = getattr(m, fn)(*args, marked)
        if torch.is_tensor(G["keep"]):
            assert torch.equal(keep, G["keep"])
        else:
            assert keep == G["keep"] and isinstance(keep, list)
    for k in ("uc_b", "c_b", "preds"):
        assert torch.equal(lay.seen[k], G[k]), (name, k)
    assert layout_dict["translations"].shape == (8, 3)


@pytest.mark.parametrize("name,fn,replace", [c for c in cases.SCENE_GLUE_CASES if not c[0].startswith("box_")])
def test_glue_matches_reference_methods(gold, name, fn, replace):
    G = gold[name]
    lay = RecLayout()
    m = scene.Sg2ScDiffModel(OracleEncoder(), lay, shape=object(), vqvae=None, replace_latent=replace, ddim_sampler_cls=RecDDIM)
    args, marked = cases.scene_glue_inputs(name)
    np.random.seed(cases.SCENE_GLUE_NP_SEED)
    torch.manual_seed(5)
    if fn == "sample":
        shape_dict, layout_dict = m.sample(*args, gen_shape=True)
        keep = None
    else:
        keep, shape_dict, layout_dict = getattr(m, fn)(*args, marked, gen_shape=True)
    # the oracle is pinned to the reference at max-abs 0, so the glue must reproduce the recorded tensors exactly
    for k in ("uc_b", "c_b", "preds"):
        assert torch.equal(lay.seen[k], G[k]), (name, k)
    for k in ("uc_s", "c_s"):
        assert torch.equal(RecDDIM.seen[k], G[k]), (name, k)
    if keep is not None:
        assert torch.equal(keep, G["keep"]) and keep.shape == (8, 1)
    # one noise draw repeated for every object (echo2shape.py:507-510); S = ddim_steps
    x_T = RecDDIM.seen["x_T"]
    assert x_T.shape == (8, 3, 16, 16, 16) and torch.equal(x_T[0], x_T[7]) and RecDDIM.seen["S"] == 100
    assert torch.equal(shape_dict["shapes"], x_T * 0.5)            # vqvae=None: the latents are returned
    # generate_layout_sg's split (echo2layout.py:118-122)
    s = torch.arange(64, dtype=torch.float32).reshape(8, 8)
    assert torch.equal(layout_dict["sizes"], s[:, 0:3]) and torch.equal(layout_dict["translations"], s[:, 3:6])
    assert torch.equal(layout_dict["angles"], s[:, 6:8])


def test_facade_and_layout_only(gold):
    lay = RecLayout()
    m = scene.Sg2ScDiffModel(OracleEncoder(), lay)
    args, _ = cases.scene_glue_inputs("sample")
    out = m.sample_box_and_shape(*args, gen_shape=False)
    assert out["shapes"] is None and set(out) == {"shapes", "sizes", "translations", "angles"}
    assert torch.equal(lay.seen["c_b"], gold["sample"]["c_b"])
    with pytest.raises(_lib.EchoError):
        m.sample(*args, gen_shape=True)                           # no shape branch attached
    args, marked = cases.scene_glue_inputs("changes")
    np.random.seed(cases.SCENE_GLUE_NP_SEED)
    keep, out = m.sample_boxes_and_shape_with_changes(*args, marked)
    assert torch.equal(keep, gold["changes"]["keep"]) and torch.equal(lay.seen["c_b"], gold["changes"]["c_b"])


def test_row_helpers():
    x = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    ref = x.clone()
    added = []
    for i, mnode in enumerate([1, 3, 9]):                          # the reference's loop, EchoScene.py:478-486
        ad = mnode + i
        added.append(ad)
        ref = torch.cat([ref[:ad], torch.zeros(1, 3), ref[ad:]], dim=0)
    got, nodes = scene.insert_zero_rows(x, [1, 3, 9])
    assert nodes == added and torch.equal(got, ref)
    got, nodes = scene.insert_zero_rows(x, [])
    assert nodes == [] and torch.equal(got, x)
    new = -x
    base = x.clone()
    for t in sorted([2, 0]):                                       # EchoScene.py:446-448
        base = torch.cat([base[:t], new[t:t + 1], base[t + 1:]], dim=0)
    assert torch.equal(scene.replace_rows(x, new, [2, 0]), base)
    with pytest.raises(IndexError):
        scene.replace_rows(x, new, [4])
    np.random.seed(3)
    want = np.zeros((5, 4))
    want[1] = np.random.normal(0, 1, 4)
    want[3] = np.random.normal(0, 1, 4)
    np.random.seed(3)
    assert torch.equal(scene.change_flags(5, [3, 1, 3], 4, "cpu"), torch.from_numpy(want).float())
    assert scene.keep_mask(4, [1], "cpu").flatten().tolist() == [1, 0, 1, 1]


def test_scene_encoder_module_surface():
    cfg = cases.scene_cfg()
    specs = arch.scene_encoder_specs(cfg)
    m = modules.SceneEncoder()
    sd = m.state_dict()
    assert list(sd.keys()) == list(specs.keys())
    for k, s in specs.items():
        assert tuple(sd[k].shape) == tuple(s.shape), k
    # a full Sg2ScDiffModel state_dict carries other sub-modules: only the encoder slice is loaded, strictly
    full = dict(arch.make_state_dict(specs, cases.WEIGHT_SEED_SCENE))
    full["obj_embeddings_dc.weight"] = torch.zeros(37, 128)
    full["LayoutDiff.df.model.out.2.weight"] = torch.zeros(1)
    m.load_reference_state_dict(full, strict=True)
    assert torch.equal(m.state_dict()["rel_s_mlp.3.weight"], full["rel_s_mlp.3.weight"])
    v = modules.SceneEncoder.from_vocab({"object_idx_to_name": ["a", "b", "b"], "pred_idx_to_name": ["in", "on"]})
    assert v.state_dict()["obj_embeddings_ec.weight"].shape == (3, 128)
    assert v.state_dict()["pred_embeddings_ec.weight"].shape == (2, 128)
    # no CPU fallback, eval only
    g, objs, text, rel = cases.scene_inputs()
    with pytest.raises(_lib.EchoError):
        m.encode(objs, g.triples, text, rel)
    with pytest.raises(_lib.EchoError):
        m.rel_s(torch.zeros(2, 640))
    with pytest.raises(_lib.EchoError):
        modules.SceneEncoder(gconv_pooling="wAvg")
    m.train()
    with pytest.raises(_lib.EchoError):
        m.init_encoder(objs, g.triples, text, rel)


def test_scene_encoder_layout_only_variant():
    """SceneEncoder(man_dc_preds=True, with_rel_s=False) = the encoder slice of the layout-only Sg2BoxDiffModel."""
    sd = cases.scene_box_state_dict()
    m = modules.SceneEncoder(man_dc_preds=True, with_rel_s=False)
    assert list(m.state_dict().keys()) == list(sd.keys()) and "pred_embeddings_man_dc.weight" in sd
    full = dict(sd)
    full["obj_embeddings_dc.weight"] = torch.zeros(37, 128)
    full["rel_s_mlp.0.weight"] = torch.zeros(960, 640)             # a full-model checkpoint offered to the layout-only encoder
    m.load_reference_state_dict(full, strict=True)
    assert torch.equal(m.state_dict()["pred_embeddings_man_dc.weight"], sd["pred_embeddings_man_dc.weight"])
    # the oracle follows the same switch: the two tables give different manipulate outputs
    g, objs, text, rel = cases.scene_inputs()
    cfg = cases.scene_box_cfg()
    with torch.no_grad():
        a = orc.scene_encode(sd, cfg, objs, g.triples, text, rel)
        b = orc.scene_encode(sd, cases.scene_cfg(), objs, g.triples, text, rel)
    assert "uc_s" not in a and torch.equal(a["obj_embed"], b["obj_embed"]) and not torch.equal(a["latent"], b["latent"])


def test_scene_entry_points_reject_bad_arguments():
    import ctypes as C
    L = _lib.lib()
    h = C.c_void_p()
    assert L.echo_scene_create(C.byref(h), None, None, 0) == -1
    d = _lib.SceneDesc(64, 512, 37, 16, 5, 960, 1280, 32, 128, 1e-5)
    assert L.echo_scene_create(None, C.byref(d), None, 0) == -1
    bad = _lib.SceneDesc(63, 512, 37, 16, 5, 960, 1280, 32, 128, 1e-5)
    assert L.echo_scene_create(C.byref(h), C.byref(bad), None, 0) == -1 and b"multiples of 4" in L.echo_last_error()
    assert L.echo_scene_encode(None, None, None, None, None, None, None, None, None, None, None) == -1
    assert L.echo_scene_rel_s(None, None, 1, None, None) == -1
    L.echo_scene_destroy(None)


def test_reference_side_binding_shares_parameters():
    """integrate.scene_encoder_of(): the SceneEncoder behind the patched init_encoder / manipulate uses the model's own tensors,
    stays out of its state_dict, is rebuilt when the parameters are replaced, and refuses train mode / CPU tensors."""
    import torch.nn as nn
    from echoscene_b200 import integrate

    class Holder(nn.Module):                                     # the encoder sub-modules as EchoScene.py:46-100 builds them
        def __init__(self):
            super().__init__()
            self.clip, self.embedding_dim = True, 64
            self.obj_embeddings_ec = nn.Embedding(37, 128)
            self.pred_embeddings_ec = nn.Embedding(16, 128)
            kw = dict(hidden_dim=256, pooling="avg", mlp_normalization="batch", residual=True)
            self.gconv_net_ec = modules.GraphTripleConvNet(640, 640, num_layers=5, output_dim=640, **kw)
            self.gconv_net_manipulation = modules.GraphTripleConvNet(1344, 640, num_layers=5, output_dim=640, **kw)
            self.rel_s_mlp = nn.Sequential(nn.Linear(640, 960), nn.BatchNorm1d(960), nn.ReLU(), nn.Linear(960, 1280))
            self.obj_embeddings_dc = nn.Embedding(37, 128)       # present in the checkpoint, unused when sampling

        init_encoder, manipulate = integrate._init_encoder, integrate._manipulate

    h = Holder().eval()
    keys = list(h.state_dict().keys())
    enc = integrate.scene_encoder_of(h)
    assert integrate.scene_encoder_of(h) is enc                                   # cached
    assert list(h.state_dict().keys()) == keys                                    # not a registered sub-module
    assert enc.state_dict()["gconv_net_ec.gconvs.0.net1.0.weight"].data_ptr() == h.gconv_net_ec.state_dict()["gconvs.0.net1.0.weight"].data_ptr()
    assert enc.state_dict()["rel_s_mlp.3.bias"].data_ptr() == h.rel_s_mlp[3].bias.data_ptr()
    assert [k for k in enc.state_dict()] == list(arch.scene_encoder_specs(cases.scene_cfg()).keys())
    g, objs, text, rel = cases.scene_inputs()
    with pytest.raises(_lib.EchoError, match="CUDA"):
        h.init_encoder(objs, g.triples, text, rel)
    # self.rel_s_mlp(x) is routed to the library in eval mode (parameters stay in the nn.Sequential) ...
    with pytest.raises(_lib.EchoError, match="CUDA"):
        h.rel_s_mlp(torch.zeros(2, 640))
    assert list(h.state_dict().keys()) == keys
    h.train()
    assert h.rel_s_mlp(torch.zeros(2, 640)).shape == (2, 1280)                     # ... and is the reference's own forward in training
    with pytest.raises(_lib.EchoError, match="eval"):
        h.manipulate(torch.zeros(8, 704), objs, g.triples, text, rel)
    h.eval()
    h.obj_embeddings_ec.weight = nn.Parameter(h.obj_embeddings_ec.weight.detach().clone())   # what .cuda() / load does
    assert integrate.scene_encoder_of(h) is not enc
    # the layout-only model (model/EchoLayout.py) owns no rel_s_mlp and embeds the predicates of `manipulate` with
    # pred_embeddings_man_dc (patch_reference marks its class): the encoder is built that way
    del h.rel_s_mlp
    h.pred_embeddings_man_dc = nn.Embedding(16, 128)
    Holder._echo_man_dc_preds = True
    box = integrate.scene_encoder_of(h)
    assert not box.with_rel_s and not any(k.startswith("rel_s_mlp.") for k in box.state_dict())
    assert box.cfg.man_dc_preds
    assert box.state_dict()["pred_embeddings_man_dc.weight"].data_ptr() == h.pred_embeddings_man_dc.weight.data_ptr()
    with pytest.raises(_lib.EchoError, match="rel_s_mlp"):
        box.rel_s(torch.zeros(2, 640))


def test_load_reference_checkpoint(tmp_path):
    """A file with the key structure SGDiff.save writes (EchoScene.py:534-544) is distributed over the components the way
    SGDiff.load_networks does it (SGDiff.py:49-84)."""
    ecfg, lcfg, vcfg = cases.scene_cfg(), cases.layout_cfg(), cases.vqvae_cfg()
    esd = arch.make_state_dict(arch.scene_encoder_specs(ecfg), 1)
    lsd = arch.make_state_dict(arch.unet1d_specs(lcfg), 2)
    vsd = arch.make_state_dict(arch.vqvae_decode_specs(vcfg), 3)
    ckpt = dict(esd)
    ckpt["obj_embeddings_dc.weight"] = torch.zeros(37, 128)                       # unused sub-modules of the model
    ckpt["pred_embeddings_man_dc.weight"] = torch.zeros(16, 128)
    ckpt.update({"LayoutDiff.df.model." + k: v for k, v in lsd.items()})
    ckpt.update({"epoch": 7, "counter": 1234, "opt": {"state": {}, "param_groups": []}})
    ckpt["vqvae"] = {"module." + k: v for k, v in vsd.items()}                   # saved from a DDP wrapper
    ckpt["vqvae"]["module.encoder.conv_in.weight"] = torch.zeros(64, 1, 3, 3, 3)
    ckpt["shape_df"] = {"diffusion_net.out.2.bias": torch.ones(3)}
    path = tmp_path / "model7.pth"
    torch.save(ckpt, path)

    class Rec:                                                                   # stands in for the 430 M-parameter UNet3DModel
        def load_state_dict(self, sd, strict=True):
            self.sd, self.strict = sd, strict

    enc = modules.SceneEncoder()
    u1 = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                             channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                             crossattn_dim=1280, enable_t_emb=True)
    dd = dict(double_z=False, z_channels=vcfg.z_channels, resolution=vcfg.resolution, in_channels=1, out_ch=vcfg.out_ch, ch=vcfg.ch,
              ch_mult=list(vcfg.ch_mult), num_res_blocks=vcfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    vq, u3 = modules.VQVAE(dd, vcfg.n_embed, vcfg.embed_dim), Rec()
    info = scene.load_reference_checkpoint(str(path), encoder=enc, unet1d=u1, unet3d=u3, vqvae=vq)
    assert info["epoch"] == 7 and info["counter"] == 1234
    assert info["loaded"] == {"encoder": len(esd), "unet1d": len(lsd), "unet3d": 1, "vqvae": len(vsd) + 1}
    for k, v in esd.items():
        assert torch.equal(enc.state_dict()[k], v), k
    for k in ("out.2.weight", "input_blocks.0.0.weight", "box_graph_cov.gconvs.4.net2.3.bias"):
        assert torch.equal(u1.state_dict()[k], lsd[k]), k
    assert torch.equal(vq.state_dict()["decoder.conv_out.bias"], vsd["decoder.conv_out.bias"])
    assert list(u3.sd) == ["out.2.bias"] and u3.strict is True
    # layout-only checkpoint: strict raises for the shape branch, strict=False carries on as load_networks does
    lay = {k: v for k, v in ckpt.items() if k not in ("vqvae", "shape_df")}
    with pytest.raises(KeyError):
        scene.load_reference_checkpoint(lay, encoder=enc, unet1d=u1, unet3d=u3)
    assert scene.load_reference_checkpoint(lay, unet1d=u1, unet3d=u3, strict=False)["loaded"] == {"unet1d": len(lsd)}
    with pytest.raises(KeyError):
        scene.load_reference_checkpoint({"epoch": 1}, unet1d=u1)


def test_sample_scenes_batches_block_diagonal_scenes():
    """BASELINE config 4 through the public surface: a collated batch of scenes in one call.  On the oracle encoders the
    batched conditioning equals the per-scene conditioning (the graph is block-diagonal), every scene gets its own shape noise
    repeated over its objects, and split_by_scene cuts the rows back."""
    from echoscene_b200 import synth
    sizes = [5, 3, 8]
    parts = [cases.scene_inputs(cases.GraphCase(f"s{i}", n, n + 2, 60 + i)) for i, n in enumerate(sizes)]
    batch = synth.batch_scene_graphs([p[0] for p in parts])
    objs, text, rel = (torch.cat([p[j] for p in parts]) for j in (1, 2, 3))
    o2s = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(sizes)])
    lay = RecLayout()
    m = scene.Sg2ScDiffModel(OracleEncoder(), lay, shape=object(), ddim_sampler_cls=RecDDIM)
    x_T = torch.randn(3, 3, 16, 16, 16, generator=torch.Generator().manual_seed(1))
    shape_dict, layout_dict, got_o2s = m.sample_scenes(objs, batch.triples, text, rel, o2s, gen_shape=True, x_T_per_scene=x_T)
    assert torch.equal(got_o2s, o2s) and layout_dict["sizes"].shape == (16, 3)
    seen_x = RecDDIM.seen["x_T"]
    assert seen_x.shape == (16, 3, 16, 16, 16)
    assert torch.equal(seen_x[0], x_T[0]) and torch.equal(seen_x[4], x_T[0]) and torch.equal(seen_x[5], x_T[1]) and torch.equal(seen_x[15], x_T[2])
    batched_uc, batched_c = RecDDIM.seen["uc_s"], lay.seen["c_b"]
    off = 0
    for (g, o, t, r), n in zip(parts, sizes):                                   # scene by scene, as the reference would run them
        m.sample(o, g.triples, t, r, gen_shape=True, x_T=x_T[:1].repeat(n, 1, 1, 1, 1))
        # same arithmetic on different matrix heights: equal up to fp32 summation order inside the CPU GEMMs
        for a, b in ((batched_uc[off:off + n], RecDDIM.seen["uc_s"]), (batched_c[off:off + n], lay.seen["c_b"])):
            assert float((a - b).abs().max() / b.abs().max()) < 1e-5
        off += n
    per = scene.split_by_scene({**shape_dict, **layout_dict}, o2s)
    assert [p["sizes"].shape[0] for p in per] == sizes and per[2]["shapes"].shape == (8, 3, 16, 16, 16)
    with pytest.raises(_lib.EchoError):
        m.sample_scenes(objs, batch.triples, text, rel, o2s, gen_shape=True, x_T_per_scene=x_T[:2])


def test_select_sdfs_greedy_matches_the_reference():
    """scene.Sg2ScDiffModel.select_sdfs against the reference's own Sg2ScDiffModel.select_sdfs(sample_type='greedy')
    (model/EchoScene.py:246-319; tests/golden/select_sdfs.pt from oracle/gen_golden_select.py): whole scenes while they fit
    diffusion_bs, the triples among the selected nodes; and the two inputs the reference fails on, refused with a message."""
    import os
    import types
    from oracle import gen_golden_select as gs
    G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "select_sdfs.pt"), map_location="cpu")
    enc = types.SimpleNamespace(embedding_dim=64, out_dim_ini_encoder=640)
    for name, sizes, bs in gs.CASES:
        o2s, objs, triples, sdfs, uc, c = gs.batch(sizes, 100 + len(sizes))
        m = scene.Sg2ScDiffModel(enc, None, diffusion_bs=bs)
        cats, d = m.select_sdfs(o2s, objs, triples, sdfs, uc, c)
        ref = G[name]
        assert torch.equal(cats, ref["obj_cat_selected"]), name
        for k in ("sdf", "uc_s", "c_s", "triples"):
            assert torch.equal(d[k], ref[k]), (name, k)
        assert np.array_equal(np.asarray(d["scene_ids"]), ref["scene_ids"].numpy()), name
    o2s, objs, triples, sdfs, uc, c = gs.batch([20, 4], 7)
    with pytest.raises(_lib.EchoError, match="first scene"):                    # the reference: torch.cat of an empty list
        scene.Sg2ScDiffModel(enc, None, diffusion_bs=16).select_sdfs(o2s, objs, triples, sdfs, uc, c)
    with pytest.raises(_lib.EchoError, match="greedy"):
        scene.Sg2ScDiffModel(enc, None, diffusion_bs=16).select_sdfs(o2s, objs, triples, sdfs, uc, c, sample_type="balance")
    with pytest.raises(_lib.EchoError, match="sorted by scene"):
        scene.Sg2ScDiffModel(enc, None, diffusion_bs=64).select_sdfs(torch.tensor([1, 0, 1, 0]), objs[:4], triples[:0], sdfs[:4], uc[:4], c[:4])


def test_classes_handed_to_the_reference_refuse_train_mode():
    """integrate.patch_reference gives the reference subclasses with train_forward_values = False: in train() mode they raise before
    anything is computed (the reference's training loop needs an autograd tape), while the package's own classes -- the SGDiff
    facade's -- compute the training forward's values."""
    class Patched(modules.GraphTripleConvNet):
        train_forward_values = False
    m = Patched(64, 16, num_layers=1, hidden_dim=32, residual=True, mlp_normalization="batch")
    m.train()
    with pytest.raises(_lib.EchoError, match="inside the patched reference"):
        m(torch.zeros(4, 64), torch.zeros(3, 16), torch.zeros(3, 2, dtype=torch.int64))
    m.eval()
    with pytest.raises(_lib.EchoError, match="CUDA"):
        m(torch.zeros(4, 64), torch.zeros(3, 16), torch.zeros(3, 2, dtype=torch.int64))
    own = modules.GraphTripleConvNet(64, 16, num_layers=1, hidden_dim=32, residual=True, mlp_normalization="batch")
    own.train()
    with pytest.raises(_lib.EchoError, match="CUDA"):                           # goes on to the (CUDA-only) batch-statistics forward
        own(torch.zeros(4, 64), torch.zeros(3, 16), torch.zeros(3, 2, dtype=torch.int64))
