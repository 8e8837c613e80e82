"""GPU: the GraphTripleConv net, the layout step/chain and the shape step/chain through the reference-facing modules
(which call the C ABI), against (a) the committed outputs of the reference itself (tests/golden) and (b) the oracle on
fresh seeded inputs; plus size-independent properties at the BASELINE sizes."""
import pytest
import torch

from echoscene_b200 import _lib, arch, modules, samplers, synth
from oracle import cases, echoscene_oracle as orc
from util import BF16_TOL, FP32_TOL, assert_close, gold

pytestmark = pytest.mark.gpu
DEV = "cuda"


def layout_model(sd, **kw):
    m = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                            channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                            conditioning_key="crossattn", concat_dim=1280, crossattn_dim=1280, use_checkpoint=True,
                            enable_t_emb=True, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def shape_model(sd, **kw):
    m = modules.UNet3DModel(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                            attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3,
                            use_spatial_transformer=True, transformer_depth=1, context_dim=1280, use_checkpoint=True,
                            legacy=False, messsage_passing=True, conditioning_key="crossattn", enable_t_emb=True, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


@pytest.fixture(scope="module")
def layout_sd():
    return arch.make_state_dict(arch.unet1d_specs(cases.layout_cfg()), cases.WEIGHT_SEED_LAYOUT)


@pytest.fixture(scope="module")
def shape_sd():
    return arch.make_state_dict(arch.unet3d_specs(cases.shape_cfg()), cases.WEIGHT_SEED_SHAPE)


# ---------------------------------------------------------------------------------------------------------------- GCN
def test_gcn_vs_reference_golden():
    cfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(cfg), cases.WEIGHT_SEED_GCN)
    net = modules.GraphTripleConvNet(cfg.input_dim_obj, cfg.input_dim_pred, num_layers=5, hidden_dim=256, residual=True,
                                     pooling="avg", mlp_normalization="batch", output_dim=1280)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    g, obj, pred = cases.gcn_inputs(cases.GCN_CASE, cfg)
    edges, _ = orc.edges_of(g.triples)
    G = gold("gcn_layout_n8.pt")
    o, p = net(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    assert_close(o, G["obj"], FP32_TOL, "gcn obj")
    assert_close(p, G["pred"], FP32_TOL, "gcn pred")
    # single layer through the GraphTripleConv surface
    one = modules.GraphTripleConv(cfg.input_dim_obj, cfg.input_dim_pred, output_dim=cfg.input_dim_obj, hidden_dim=256,
                                  pooling="avg", mlp_normalization="batch", residual=True)
    one.load_state_dict({k[len("gconvs.0."):]: v for k, v in sd.items() if k.startswith("gconvs.0.")}, strict=True)
    one = one.to(DEV)
    o1, p1 = one(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    assert_close(o1, G["layer0_obj"], 1e-4, "gcn layer0 obj")
    assert_close(p1, G["layer0_pred"], 1e-4, "gcn layer0 pred")


@pytest.mark.parametrize("n,t,seed", [(1, 0, 1), (5, 4, 2), (32, 128, 3), (200, 900, 4)])
def test_gcn_edge_cases_vs_oracle(n, t, seed):
    """isolated nodes (empty edge list), self loops (counted twice, graph.py:191-192), duplicates, many rows."""
    cfg = arch.GCNConfig(96, 32, 2, 64, 160, True, "avg", "batch")
    sd = arch.make_state_dict(arch.gcn_specs(cfg), 77)
    net = modules.GraphTripleConvNet(96, 32, num_layers=2, hidden_dim=64, residual=True, mlp_normalization="batch", output_dim=160)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    gen = torch.Generator().manual_seed(seed)
    obj, pred = torch.randn(n, 96, generator=gen), torch.randn(t, 32, generator=gen)
    edges = torch.randint(0, n, (t, 2), generator=gen)
    if t >= 4:
        edges[0] = torch.tensor([0, 0])        # self loop
        edges[1] = edges[2]                    # duplicate edge
    o, p = net(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    wo, wp = orc.graph_triple_conv_net(sd, "", obj, pred, edges, num_layers=2)
    assert_close(o, wo, 1e-4, "gcn obj")
    if t:
        assert_close(p, wp, 1e-4, "gcn pred")


def test_graph_rejects_out_of_range_index():
    tri = torch.tensor([[0, 1, 9]], dtype=torch.int64, device=DEV)
    with pytest.raises(_lib.EchoError):
        _lib.Graph(tri, 4)


# ------------------------------------------------------------------------------------------------------------- layout
def test_layout_step_vs_reference_golden(layout_sd):
    cfg = cases.layout_cfg()
    m = layout_model(layout_sd)
    g, obj_embed, x, t = cases.layout_step_inputs(cases.LAYOUT_CASE, cfg)
    out = m(x.to(DEV), obj_embed.to(DEV), g.triples.to(DEV), t.to(DEV), None)
    assert out.shape == (8, 8, 1)
    assert_close(out, gold("layout_n8.pt")["step"], FP32_TOL, "layout step")


def test_layout_chain_config1_vs_reference_golden(layout_sd):
    """BASELINE config 1: N=8, layout only, 10 DDPM steps, through DiffusionPoint.gen_samples_sg with injected noise."""
    cfg = cases.layout_cfg()
    m = layout_model(layout_sd)
    dp = samplers.DiffusionPoint(m, {}, time_num=cases.LAYOUT_CHAIN_STEPS, beta_start=1e-4, beta_end=0.02)
    g, obj_embed, x_T, noises = cases.layout_chain_inputs(cases.LAYOUT_CASE, cfg, cases.LAYOUT_CHAIN_STEPS)
    it = iter([x_T] + list(noises))
    out = dp.gen_samples_sg((8, 8), DEV, obj_embed.to(DEV), triples=g.triples.to(DEV), condition=None,
                            noise_fn=lambda size, dtype, device: next(it).to(device), clip_denoised=False)
    assert_close(out, gold("layout_n8.pt")["chain"], FP32_TOL, "layout 10-step chain")


def test_layout_schedule_tables(layout_sd):
    m = layout_model(layout_sd, time_num=1000)
    assert_close(m.schedule_tables(), orc.DDPMSchedule(time_num=1000).tables(), 1e-6, "ddpm tables")


@pytest.mark.parametrize("n,t", [(16, 64), (32, 128), (1, 0)])
def test_layout_step_sizes_vs_oracle(layout_sd, n, t):
    cfg = cases.layout_cfg()
    m = layout_model(layout_sd)
    g = synth.make_scene_graph(n, t, 20 + n) if n > 1 else synth.SceneGraph(1, torch.zeros(0, 3, dtype=torch.int64))
    obj_embed, x = synth.layout_inputs(n, 30 + n)
    ts = torch.randint(0, 1000, (n,), generator=torch.Generator().manual_seed(n))
    out = m(x.to(DEV), obj_embed.to(DEV), g.triples.to(DEV), ts.to(DEV)).squeeze(-1)
    with torch.no_grad():
        want = orc.unet1d_forward(layout_sd, cfg, x, obj_embed, g.triples, ts).squeeze(-1)
    assert_close(out, want, FP32_TOL, f"layout step N={n}")


def test_layout_batched_scenes_are_independent(layout_sd):
    """collate-style disjoint union (threedfront_dataset.py:698-701): a batch equals its scenes run one by one."""
    m = layout_model(layout_sd)
    gs = [synth.make_scene_graph(8, 32, 1), synth.make_scene_graph(16, 64, 2)]
    ins = [synth.layout_inputs(8, 41), synth.layout_inputs(16, 42)]
    b = synth.batch_scene_graphs(gs)
    obj = torch.cat([i[0] for i in ins]).to(DEV)
    x = torch.cat([i[1] for i in ins]).to(DEV)
    ts = torch.cat([torch.full((8,), 500), torch.full((16,), 30)]).to(DEV)
    full = m(x, obj, b.triples.to(DEV), ts)
    a = m(x[:8], obj[:8], gs[0].triples.to(DEV), ts[:8])
    c = m(x[8:], obj[8:], gs[1].triples.to(DEV), ts[8:])
    # few-row and tiled kernels sum K in different orders (the batch has 96 edge rows > 64): equal to rounding
    assert_close(full[:8], a, 1e-5, "scene 0")
    assert_close(full[8:], c, 1e-5, "scene 1")


def test_layout_large_batch_runs_on_the_tensor_core_path_and_equals_its_scenes(layout_sd):
    """Six scenes collated (96 nodes, 384 triples: more than 64 rows, so every Linear's prologue is materialised and the contraction
    runs as the 3 x TF32 tensor-core GEMM, csrc/sgemm_x3.cu) against the same scenes one by one on the few-row path, which is the
    path pinned to the oracle: the forward, and a DDPM step through the replayed graph."""
    m = layout_model(layout_sd)
    gs = [synth.make_scene_graph(16, 64, 60 + i) for i in range(6)]
    ins = [synth.layout_inputs(16, 70 + i) for i in range(6)]
    b = synth.batch_scene_graphs(gs)
    obj = torch.cat([i[0] for i in ins]).to(DEV)
    x = torch.cat([i[1] for i in ins]).to(DEV)
    ts = torch.cat([torch.full((16,), 100 + 150 * i) for i in range(6)]).to(DEV)
    full = m(x, obj, b.triples.to(DEV), ts)
    noise = torch.randn(96, 8, generator=torch.Generator().manual_seed(3)).to(DEV)
    step = m.ddpm_step(x, obj, b.triples.to(DEV), 321, noise)
    step2 = m.ddpm_step(x, obj, b.triples.to(DEV), 321, noise)               # second call: the replayed graph
    assert torch.equal(step, step2)
    for i, g in enumerate(gs):
        r = slice(16 * i, 16 * i + 16)
        one = m(x[r], obj[r], g.triples.to(DEV), ts[r])
        assert_close(full[r], one, 1e-4, f"forward, scene {i}")
        assert_close(step[r], m.ddpm_step(x[r], obj[r], g.triples.to(DEV), 321, noise[r]), 1e-4, f"DDPM step, scene {i}")


# -------------------------------------------------------------------------------------------------------------- shape
def test_shape_step_vs_reference_golden(shape_sd):
    cfg = cases.shape_cfg()
    m = shape_model(shape_sd)
    g, uc, x, t = cases.shape_step_inputs(cases.SHAPE_CASE, cfg)
    out = m(x.to(DEV), uc.to(DEV), g.triples.to(DEV), t.to(DEV), context=uc.to(DEV))
    G = gold("shape.pt")
    assert_close(m.last_latent(4), G["latent"], FP32_TOL, "latent_shape_rel")
    assert_close(out, G["step"], FP32_TOL, "shape step")


def test_shape_chain_vs_reference_golden(shape_sd):
    """first 3 iterations of the 100-step DDIM schedule through DDIMSampler-style stepping."""
    cfg = cases.shape_cfg()
    m = shape_model(shape_sd)
    g, uc, x_T, _ = cases.shape_step_inputs(cases.SHAPE_CHAIN_CASE, cfg, same_noise=True)
    x = x_T.to(DEV)
    for i in range(cases.SHAPE_CHAIN_STEPS):
        x = m.ddim_step(x, uc.to(DEV), g.triples.to(DEV), 100 - i - 1)
    assert_close(x, gold("shape.pt")["chain"], FP32_TOL, "shape 3-step DDIM chain")


def test_shape_schedule_tables(shape_sd):
    m = shape_model(shape_sd)
    coef, ts = m.schedule_tables()
    sch = orc.DDIMSchedule(100)
    assert ts.tolist() == sch.ddim_timesteps.tolist()
    assert_close(coef, sch.table(), 1e-6, "ddim coefficients")
    m.set_schedule(250)                       # BASELINE config 3 realised as S = 250 (timesteps range(0,1000,4)+1)
    coef, ts = m.schedule_tables()
    sch = orc.DDIMSchedule(250)
    assert ts.tolist() == sch.ddim_timesteps.tolist()
    assert_close(coef, sch.table(), 1e-6, "ddim coefficients S=250")


def test_shape_sampler_surface_and_sharded_trunk(shape_sd):
    """DDIMSampler.sample over 2 steps == manual stepping; per-object shard (embed -> codes -> trunk on a slice) is
    bit-identical to the unsharded step."""
    m = shape_model(shape_sd, ddim_steps=10)
    g = synth.make_scene_graph(3, 4, 9)
    uc, x_T = synth.shape_inputs(3, 90, same_noise=True)
    tri, ucd, xd = g.triples.to(DEV), uc.to(DEV), x_T.to(DEV)

    class Holder:                              # stands in for EchoToShape: exposes .df.diffusion_net
        num_timesteps = 1000
    h = Holder()
    h.df = type("DF", (), {})()
    h.df.diffusion_net = m
    smp = samplers.DDIMSampler(h)
    out, _ = smp.sample(S=10, batch_size=3, shape=(3, 16, 16, 16), conditioning=ucd, x_T=xd,
                        unconditional_guidance_scale=3.0, unconditional_conditioning=ucd, triplet=tri, eta=0.0, verbose=False)
    # a full 10-step chain must stay finite; the 2-step chain is compared exactly below
    m2 = shape_model(shape_sd, ddim_steps=2)
    a = m2.ddim_step(xd, ucd, tri, 1)
    b = m2.ddim_step(a, ucd, tri, 0)
    smp2 = samplers.DDIMSampler(m2)
    out2, _ = smp2.sample(S=2, batch_size=3, shape=(3, 16, 16, 16), x_T=xd, unconditional_conditioning=ucd, triplet=tri, verbose=False)
    assert torch.equal(out2, b)
    assert torch.isfinite(out).all()
    # shard: ranks own objects [0,2) and [2,3)
    full = m2.ddim_step(xd, ucd, tri, 1)
    ms = shape_model(shape_sd, ddim_steps=2)
    codes = torch.cat([ms.embed_local(xd[:2], 3, 4), ms.embed_local(xd[2:], 3, 4)])
    p0 = ms.trunk_local(xd[:2], 0, codes, ucd, tri, index=1)
    p1 = ms.trunk_local(xd[2:], 2, codes, ucd, tri, index=1)
    assert torch.equal(torch.cat([p0, p1]), full)


def test_shape_objects_couple_only_through_the_graph(shape_sd):
    """Two disjoint scenes in one batch == each scene alone (block-diagonal graph, per-object trunk)."""
    m = shape_model(shape_sd)
    gs = [synth.make_scene_graph(2, 1, 1), synth.make_scene_graph(3, 4, 2)]
    b = synth.batch_scene_graphs(gs)
    uc, x = synth.shape_inputs(5, 50, same_noise=False)
    ts = torch.tensor([991, 991, 401, 401, 401])
    full = m(x.to(DEV), uc.to(DEV), b.triples.to(DEV), ts.to(DEV))
    a = m(x[:2].to(DEV), uc[:2].to(DEV), gs[0].triples.to(DEV), ts[:2].to(DEV))
    c = m(x[2:].to(DEV), uc[2:].to(DEV), gs[1].triples.to(DEV), ts[2:].to(DEV))
    assert torch.equal(full[:2], a) and torch.equal(full[2:], c)


def test_shape_step_bf16_vs_reference_golden(shape_sd):
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    cfg = cases.shape_cfg()
    m = shape_model(shape_sd, precision="bf16")
    g, uc, x, t = cases.shape_step_inputs(cases.SHAPE_CASE, cfg)
    out = m(x.to(DEV), uc.to(DEV), g.triples.to(DEV), t.to(DEV))
    assert_close(out, gold("shape.pt")["step"], BF16_TOL, "shape step bf16")


def test_full_size_determinism_n16(shape_sd):
    """BASELINE config 2 size (N=16, T=64): two runs of the same DDIM step are bit-identical (no float atomics),
    the output is finite, and all objects that share x_T but differ in conditioning get different updates."""
    m = shape_model(shape_sd, precision="bf16" if _lib.lib().echo_has_tcgen05() else "fp32")
    g = synth.make_scene_graph(16, 64, 2)
    uc, x_T = synth.shape_inputs(16, 2, same_noise=True)
    a = m.ddim_step(x_T.to(DEV), uc.to(DEV), g.triples.to(DEV), 99)
    b = m.ddim_step(x_T.to(DEV), uc.to(DEV), g.triples.to(DEV), 99)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    assert (a[0] - a[1]).abs().max() > 0


def test_shape_bf16_full_size_matches_small_batches(shape_sd):
    """BASELINE config 2 size in the tensor-core mode: 4 disjoint scenes of 4 nodes batched into one N = 16 step.
    The launch plans differ with the object count (sub-blocks per CTA, split-K, tile widths), so the full-size step is
    checked (a) against the fp32 parity path on the same inputs at the bf16 tolerance (the fp32 path is pinned to the
    reference at 1e-3 by the golden fixtures) and (b) against each scene stepped alone in bf16: only summation order
    differs, but every layer re-rounds to bf16, so two bf16 evaluations sit ~sqrt(2) x the bf16-vs-fp32 distance apart
    (measured 0.9e-2 rel-L2); a broken plan would be O(1) off."""
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    m = shape_model(shape_sd, precision="bf16")
    gs = [synth.make_scene_graph(4, 6, 10 + i) for i in range(4)]
    b = synth.batch_scene_graphs(gs)
    uc, x = synth.shape_inputs(16, 77, same_noise=False)
    ts = torch.full((16,), 501, dtype=torch.int64)
    full = m(x.to(DEV), uc.to(DEV), b.triples.to(DEV), ts.to(DEV))
    assert torch.isfinite(full).all()
    m32 = shape_model(shape_sd)
    want = m32(x.to(DEV), uc.to(DEV), b.triples.to(DEV), ts.to(DEV))
    assert_close(full, want, BF16_TOL, "N=16 bf16 step vs the fp32 parity path")
    for i, g in enumerate(gs):
        sl = slice(4 * i, 4 * i + 4)
        alone = m(x[sl].to(DEV), uc[sl].to(DEV), g.triples.to(DEV), ts[sl].to(DEV))
        assert_close(full[sl], alone, 2e-2, f"scene {i} inside the N=16 batch vs alone")


def test_shape_bf16_n32_finite_and_deterministic(shape_sd):
    """BASELINE config 3 size (N = 32, T = 128), S = 250 schedule: finite, run-to-run bit-identical."""
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    m = shape_model(shape_sd, precision="bf16", ddim_steps=250)
    g = synth.make_scene_graph(32, 128, 3)
    uc, x_T = synth.shape_inputs(32, 3, same_noise=True)
    a = m.ddim_step(x_T.to(DEV), uc.to(DEV), g.triples.to(DEV), 249)
    b = m.ddim_step(x_T.to(DEV), uc.to(DEV), g.triples.to(DEV), 249)
    assert torch.equal(a, b) and torch.isfinite(a).all()


def test_sharded_trunk_with_codes_on_their_own_stream(shape_sd):
    """echo_shape_trunk_async: embed + (stand-in for the) all-gather queued on a second stream, trunk on the current one;
    result is bit-identical to the unsharded step."""
    m = shape_model(shape_sd, ddim_steps=2)
    g = synth.make_scene_graph(3, 4, 9)
    uc, x_T = synth.shape_inputs(3, 90, same_noise=True)
    tri, ucd, xd = g.triples.to(DEV), uc.to(DEV), x_T.to(DEV)
    full = m.ddim_step(xd, ucd, tri, 1)
    ms = shape_model(shape_sd, ddim_steps=2)
    s2 = torch.cuda.Stream()
    codes_all = torch.empty(3, 64, device=DEV)
    outs = []
    for (lo, hi) in [(0, 2), (2, 3)]:
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            codes_all[0:2] = ms.embed_local(xd[:2], 3, 4)     # every "rank"'s codes (the all-gather)
            codes_all[2:3] = ms.embed_local(xd[2:], 3, 4)
        outs.append(ms.trunk_local(xd[lo:hi], lo, codes_all, ucd, tri, index=1, codes_stream=s2))
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(outs), full)


def test_sharded_trunk_restricted_to_the_components_of_its_objects(shape_sd):
    """trunk_local(restrict_to_components=True): a rank that owns the second scene of a two-scene batch runs the echo GCN on that
    scene only (shard.echo_components) and reproduces the batched step; with the exchange on its own stream too."""
    m = shape_model(shape_sd, ddim_steps=2)
    gs = [synth.make_scene_graph(3, 4, 1), synth.make_scene_graph(4, 6, 2)]
    b = synth.batch_scene_graphs(gs)
    uc, x_T = synth.shape_inputs(7, 50, same_noise=False)
    tri, ucd, xd = b.triples.to(DEV), uc.to(DEV), x_T.to(DEV)
    full = m.ddim_step(xd, ucd, tri, 1)
    ms = shape_model(shape_sd, ddim_steps=2)
    codes = torch.cat([ms.embed_local(xd[:3], 7, tri.shape[0]), ms.embed_local(xd[3:], 7, tri.shape[0])])
    plain = ms.trunk_local(xd[3:], 3, codes, ucd, tri, index=1)
    assert torch.equal(plain, full[3:])
    got = ms.trunk_local(xd[3:], 3, codes, ucd, tri, index=1, restrict_to_components=True)
    assert_close(got, full[3:], 1e-5, "component-restricted echo, rank 1")
    got0 = ms.trunk_local(xd[:3], 0, codes, ucd, tri, index=1, restrict_to_components=True)
    assert_close(got0, full[:3], 1e-5, "component-restricted echo, rank 0")
    s2 = torch.cuda.Stream()
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        codes2 = codes.clone()
    got2 = ms.trunk_local(xd[3:], 3, codes2, ucd, tri, index=1, codes_stream=s2, restrict_to_components=True)
    torch.cuda.synchronize()
    assert torch.equal(got2, got)
    # one connected scene: nothing to restrict, the call is the plain one
    g1 = synth.make_scene_graph(3, 4, 9)
    uc1, x1 = synth.shape_inputs(3, 90, same_noise=True)
    t1, u1, xx1 = g1.triples.to(DEV), uc1.to(DEV), x1.to(DEV)
    c1 = torch.cat([ms.embed_local(xx1[:2], 3, 4), ms.embed_local(xx1[2:], 3, 4)])
    assert torch.equal(ms.trunk_local(xx1[:2], 0, c1, u1, t1, index=1, restrict_to_components=True), ms.trunk_local(xx1[:2], 0, c1, u1, t1, index=1))
