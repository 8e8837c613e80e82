"""GPU: the once-per-scene encoders (csrc/scene.cu behind modules.SceneEncoder / echo_scene_*) against
(a) the committed outputs of the reference's own init_encoder / manipulate / rel_s_mlp (tests/golden/scene_encode.pt),
(b) what the reference's sample_with_changes / sample_with_additions hand to the two branches (tests/golden/scene_glue.pt),
(c) the oracle on other sizes; then the whole `sample` surface (encoders -> DDPM layout chain -> DDIM shape chain) against
the oracle's chains.  fp32 contract: 1e-3 relative (tests/util.py)."""
import numpy as np
import pytest
import torch

from echoscene_b200 import _lib, arch, modules, samplers, scene
from oracle import cases, echoscene_oracle as orc
from util import FP32_TOL, assert_close, gold

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def enc():
    cfg = cases.scene_cfg()
    sd = arch.make_state_dict(arch.scene_encoder_specs(cfg), cases.WEIGHT_SEED_SCENE)
    m = modules.SceneEncoder()
    m.load_state_dict(sd, strict=True)
    return m.to(DEV), sd, cfg


def _cuda(*ts):
    return [t.to(DEV) for t in ts]


def test_scene_encode_vs_reference_golden(enc):
    m, sd, cfg = enc
    g, objs, text, rel = cases.scene_inputs()
    G = gold("scene_encode.pt")
    out = m.encode(*_cuda(objs, g.triples, text, rel))
    for k in ("obj_embed", "latent", "uc_s", "c_s"):
        assert_close(out[k], G[k], FP32_TOL, f"scene_encode {k}")
    # the embedding half of obj_embed is a row copy: bit-exact
    assert torch.equal(out["obj_embed"].cpu(), G["obj_embed"])
    # layout-only call: the shape conditionings are skipped
    out2 = m.encode(*_cuda(objs, g.triples, text, rel), shape_cond=False)
    assert set(out2) == {"obj_embed", "latent"} and torch.equal(out2["latent"], out["latent"])


def test_scene_stage_entry_points_vs_oracle(enc):
    """init_encoder / manipulate / rel_s one by one (the decomposition sample_with_changes needs), incl. a non-zero change."""
    m, sd, cfg = enc
    g, objs, text, rel = cases.scene_inputs()
    d_objs, d_tri, d_text, d_rel = _cuda(objs, g.triples, text, rel)
    oe, pe, lo, _ = m.init_encoder(d_objs, d_tri, d_text, d_rel)
    with torch.no_grad():
        w_oe, w_pe, w_lo, _ = orc.scene_init_encoder(sd, cfg, objs, g.triples, text, rel)
    assert torch.equal(oe.cpu(), w_oe) and torch.equal(pe.cpu(), w_pe)
    assert_close(lo, w_lo, FP32_TOL, "init_encoder latent")
    change = torch.randn(8, cfg.gconv_dim, generator=torch.Generator().manual_seed(9))
    latent_f = torch.cat([w_lo, change], dim=1)
    lt, _, oe2, pe2 = m.manipulate(latent_f.to(DEV), d_objs, d_tri, d_text, d_rel)
    with torch.no_grad():
        w_lt, _, w_oe2, w_pe2 = orc.scene_manipulate(sd, cfg, latent_f, objs, g.triples, text, rel)
        w_rs = orc.scene_rel_s(sd, w_lt)
    assert_close(lt, w_lt, FP32_TOL, "manipulate latent")
    assert torch.equal(oe2.cpu(), w_oe2) and torch.equal(pe2.cpu(), w_pe2)
    assert_close(m.rel_s(w_lt.to(DEV)), w_rs, FP32_TOL, "rel_s")
    assert m.rel_s(torch.zeros(0, 640, device=DEV)).shape == (0, 1280)
    # encode(change=...) == init_encoder -> manipulate -> rel_s
    out = m.encode(d_objs, d_tri, d_text, d_rel, change=change.to(DEV))
    with torch.no_grad():
        want = orc.scene_encode(sd, cfg, objs, g.triples, text, rel, change)
    for k in ("latent", "uc_s", "c_s"):
        assert_close(out[k], want[k], FP32_TOL, f"encode(change) {k}")


@pytest.mark.parametrize("n,t,seed", [(1, 0, 21), (2, 1, 22), (16, 64, 23), (96, 400, 24)])
def test_scene_encode_sizes_vs_oracle(enc, n, t, seed):
    """a lone node without edges, the BASELINE config-2 graph, a batch of scenes (rows > 64: the tiled GEMM path)."""
    m, sd, cfg = enc
    if n >= 3:
        g, objs, text, rel = cases.scene_inputs(cases.GraphCase("x", n, t, seed))
        tri = g.triples
    else:
        gen = torch.Generator().manual_seed(seed)
        objs = torch.randint(0, cfg.num_objs + 1, (n,), generator=gen)
        tri = torch.tensor([[1, 3, 0]], dtype=torch.int64)[:t].reshape(t, 3)
        text, rel = torch.randn(n, 512, generator=gen), torch.randn(t, 512, generator=gen)
    out = m.encode(*_cuda(objs, tri, text, rel))
    with torch.no_grad():
        want = orc.scene_encode(sd, cfg, objs, tri, text, rel)
    for k in ("obj_embed", "latent", "uc_s", "c_s"):
        assert_close(out[k], want[k], FP32_TOL, f"scene_encode n={n} {k}")


def test_scene_encoder_rejects_bad_ids(enc):
    m, sd, cfg = enc
    g, objs, text, rel = cases.scene_inputs()
    bad = objs.clone()
    bad[2] = cfg.num_objs + 1
    with pytest.raises(IndexError):
        m.encode(*_cuda(bad, g.triples, text, rel))
    tri = g.triples.clone()
    tri[5, 1] = cfg.num_preds
    with pytest.raises(_lib.EchoError, match="predicate ids"):
        m.encode(*_cuda(objs, tri, text, rel))
    with pytest.raises(_lib.EchoError):
        m.encode(*_cuda(objs, g.triples, text[:, :100], rel))


class RecLayout:
    def gen_samples_sg(self, shape, device, obj_embed, triples=None, condition=None, clip_denoised=True, **kw):
        self.seen = {"uc_b": obj_embed, "c_b": condition}
        return torch.zeros(shape, device=device)


class RecDDIM:
    seen = None

    def __init__(self, model):
        pass

    def sample(self, S, batch_size, shape, conditioning=None, x_T=None, unconditional_conditioning=None, **kw):
        RecDDIM.seen = {"c_s": conditioning, "uc_s": unconditional_conditioning}
        return x_T, {}


@pytest.mark.parametrize("name,fn,replace", [c for c in cases.SCENE_GLUE_CASES if c[1].startswith("sample_with")])
def test_changes_and_additions_vs_reference_golden(enc, name, fn, replace):
    """sample_with_changes / sample_with_additions with the CUDA encoders: the conditioning tensors that reach the two branches
    against those recorded from the reference's own methods."""
    m, sd, cfg = enc
    G = gold("scene_glue.pt")[name]
    lay = RecLayout()
    model = scene.Sg2ScDiffModel(m, lay, shape=object(), replace_latent=replace, ddim_sampler_cls=RecDDIM)
    args, marked = cases.scene_glue_inputs(name)
    np.random.seed(cases.SCENE_GLUE_NP_SEED)
    keep, shape_dict, layout_dict = getattr(model, fn)(*_cuda(*args), marked, gen_shape=True)
    assert torch.equal(keep.cpu(), G["keep"])
    for k in ("uc_b", "c_b"):
        assert_close(lay.seen[k], G[k], FP32_TOL, f"{name} {k}")
    for k in ("uc_s", "c_s"):
        assert_close(RecDDIM.seen[k], G[k], FP32_TOL, f"{name} {k}")
    assert layout_dict["sizes"].shape == (8, 3) and layout_dict["angles"].shape == (8, 2)


def test_sample_surface_end_to_end_vs_oracle(enc):
    """Sg2ScDiffModel.sample(gen_shape=True) on the B200 components: encoders -> 10-step DDPM layout chain -> 3-step DDIM shape
    chain (no VQ-VAE: the latents are returned), against the oracle's encoders and chains on the same noise."""
    m, sd, cfg = enc
    lcfg, scfg = cases.layout_cfg(), cases.shape_cfg()
    lsd = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
    ssd = arch.make_state_dict(arch.unet3d_specs(scfg), cases.WEIGHT_SEED_SHAPE)
    u1 = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                             channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                             crossattn_dim=1280, enable_t_emb=True)
    u1.load_state_dict(lsd, strict=True)
    u3 = modules.UNet3DModel(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                             attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3,
                             use_spatial_transformer=True, transformer_depth=1, context_dim=1280, legacy=False,
                             messsage_passing=True, conditioning_key="crossattn", enable_t_emb=True)
    u3.load_state_dict(ssd, strict=True)
    steps = 10
    dp = samplers.DiffusionPoint(u1.to(DEV), time_num=steps)
    n, t = 3, 4
    g, objs, text, rel = cases.scene_inputs(cases.GraphCase("e2e", n, t, 31))
    gen = torch.Generator().manual_seed(32)
    noises = [torch.randn(n, 8, generator=gen) for _ in range(steps + 1)]
    x_T = torch.randn(1, 3, 16, 16, 16, generator=gen).repeat(n, 1, 1, 1, 1)
    it = iter(noises)
    real = dp.gen_samples_sg

    def with_noise(shape, device, obj_embed, triples=None, condition=None, clip_denoised=False, **kw):   # the reference's RNG hook
        return real(shape, device, obj_embed, triples, condition, noise_fn=lambda size, dtype, device: next(it).to(device),
                    clip_denoised=clip_denoised)

    dp.gen_samples_sg = with_noise
    model = scene.Sg2ScDiffModel(m, dp, shape=u3.to(DEV), vqvae=None, ddim_steps=100)
    # three DDIM iterations only: a sampler that stops after SHAPE_CHAIN_STEPS of the S = 100 schedule
    k = cases.SHAPE_CHAIN_STEPS

    class ShortDDIM(samplers.DDIMSampler):
        def sample(self, *a, **kw):
            # run the LAST k indices of the schedule (index 99, 98, 97), as oracle.shape_chain(n_steps=k) does
            S = kw["S"]
            self.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
            img = kw["x_T"].float().contiguous().clone()
            nxt = torch.empty_like(img)
            self.unet._ensure(img.shape[0], kw["triplet"].shape[0])
            for i in range(k):
                self.unet.ddim_step(img, kw["unconditional_conditioning"], kw["triplet"], S - 1 - i, out=nxt)
                img, nxt = nxt, img
            return img, {}

    model._ddim_cls = ShortDDIM
    shape_dict, layout_dict = model.sample(*_cuda(objs, g.triples, text, rel), gen_shape=True, x_T=x_T.to(DEV))
    with torch.no_grad():
        e = orc.scene_encode(sd, cfg, objs, g.triples, text, rel)
        boxes = orc.layout_chain(lsd, lcfg, e["obj_embed"], g.triples, noises[0], noises[1:], steps)
        lat = orc.shape_chain(ssd, scfg, e["uc_s"], g.triples, x_T, 100, k)
    got = torch.cat([layout_dict["sizes"], layout_dict["translations"], layout_dict["angles"]], dim=1)
    assert_close(got, boxes, FP32_TOL, "sample(): layout chain")        # 10 chained steps, as test_model_gpu's chain test
    assert_close(shape_dict["shapes"], lat, FP32_TOL, "sample(): shape chain")


def test_config1_echolayout_sampleBoxes_vs_oracle(enc):
    """BASELINE config 1 through the native surface: echolayout, N = 8 nodes / 32 triples, layout only, 10 DDPM steps --
    Sg2BoxDiffModel.sampleBoxes (EchoLayout.py:291-307) = scene encoders -> DiffusionPoint chain, against the oracle."""
    m, sd, cfg = enc
    lcfg = cases.layout_cfg()
    lsd = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
    u1 = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                             channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                             crossattn_dim=1280, enable_t_emb=True)
    u1.load_state_dict(lsd, strict=True)
    steps = cases.LAYOUT_CHAIN_STEPS
    dp = samplers.DiffusionPoint(u1.to(DEV), time_num=steps)
    g, objs, text, rel = cases.scene_inputs()                                   # the 8-node / 32-triple scene
    gen = torch.Generator().manual_seed(41)
    noises = [torch.randn(8, 8, generator=gen) for _ in range(steps + 1)]
    it = iter(noises)
    real = dp.gen_samples_sg
    dp.gen_samples_sg = lambda shape, device, obj_embed, triples=None, condition=None, clip_denoised=False, **kw: real(
        shape, device, obj_embed, triples, condition, noise_fn=lambda size, dtype, device: next(it).to(device),
        clip_denoised=clip_denoised)
    model = scene.Sg2BoxDiffModel(m, dp)
    out = model.sample_box_and_shape(*_cuda(objs, g.triples, text, rel))         # SGDiff.sample_box_and_shape, type_ 'echolayout'
    assert set(out) == {"sizes", "translations", "angles"}
    with torch.no_grad():
        e = orc.scene_encode(sd, cfg, objs, g.triples, text, rel)
        boxes = orc.layout_chain(lsd, lcfg, e["obj_embed"], g.triples, noises[0], noises[1:], steps)
    got = torch.cat([out["sizes"], out["translations"], out["angles"]], dim=1)
    assert got.shape == (8, 8)
    assert_close(got, boxes, FP32_TOL, "config 1: sampleBoxes 10-step chain")


@pytest.mark.parametrize("name,fn,replace", [c for c in cases.SCENE_GLUE_CASES if c[0].startswith("box_")])
def test_layout_only_model_vs_reference_golden(name, fn, replace):
    """The layout-only Sg2BoxDiffModel (model/EchoLayout.py:291-401): SceneEncoder(man_dc_preds=True, with_rel_s=False) --
    `manipulate` embeds predicates with pred_embeddings_man_dc (:154) -- under sampleBoxes / _with_changes / _with_additions,
    against the conditioning recorded from the reference's own methods."""
    sd = cases.scene_box_state_dict()
    m = modules.SceneEncoder(man_dc_preds=True, with_rel_s=False)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    G = gold("scene_glue.pt")[name]
    lay = RecLayout()
    model = scene.Sg2BoxDiffModel(m, lay, replace_latent=replace)
    args, marked = cases.scene_glue_inputs(name)
    np.random.seed(cases.SCENE_GLUE_NP_SEED)
    if fn == "sampleBoxes":
        layout_dict = model.sampleBoxes(*_cuda(*args))
    else:
        keep, layout_dict = getattr(model, fn)(*_cuda(*args), marked)
        assert (keep.cpu().flatten().tolist() if torch.is_tensor(keep) else keep) == (
            G["keep"].flatten().tolist() if torch.is_tensor(G["keep"]) else G["keep"])
    for k in ("uc_b", "c_b"):
        assert_close(lay.seen[k], G[k], FP32_TOL, f"{name} {k}")
    assert layout_dict["translations"].shape == (8, 3)
    g, objs, text, rel = cases.scene_inputs()
    with pytest.raises(_lib.EchoError, match="rel_s_mlp"):
        m.encode(*_cuda(objs, g.triples, text, rel), shape_cond=True)
