"""CPU: the oracle restatement against the committed reference outputs (tests/golden/, made by oracle/gen_golden.py
from the reference's own modules).  Bit-exact on the machine that generated them; 1e-6 elsewhere (BLAS blocking)."""
import json
import os

import torch

from echoscene_b200 import arch
from oracle import cases, echoscene_oracle as orc
from util import GOLD, gold, rel_err

TOL = 1e-5


def test_pinning_record():
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))
    for name, rec in pin["cases"].items():
        if "rel_l2" in rec:
            assert rec["rel_l2"] < 1e-6, name
    assert pin["cases"]["ddim_tables_100"]["timesteps_equal"]
    assert pin["param_counts"] == {"unet1d": arch.count_params(arch.unet1d_specs(cases.layout_cfg())),
                                   "unet3d": arch.count_params(arch.unet3d_specs(cases.shape_cfg()))}


def test_gcn_oracle_vs_reference():
    cfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(cfg), cases.WEIGHT_SEED_GCN)
    g, obj, pred = cases.gcn_inputs(cases.GCN_CASE, cfg)
    edges, _ = orc.edges_of(g.triples)
    G = gold("gcn_layout_n8.pt")
    o_obj, o_pred = orc.graph_triple_conv_net(sd, "", obj, pred, edges)
    assert max(rel_err(o_obj, G["obj"])) < TOL and max(rel_err(o_pred, G["pred"])) < TOL
    l_obj, l_pred = orc.graph_triple_conv(sd, "gconvs.0.", obj, pred, edges)
    assert max(rel_err(l_obj, G["layer0_obj"])) < TOL and max(rel_err(l_pred, G["layer0_pred"])) < TOL
    assert torch.equal(orc.gather_rows(obj, edges[:, 0]), G["gather_s"])
    assert torch.equal(orc.gather_rows(obj, edges[:, 1]), G["gather_o"])


def test_layout_oracle_vs_reference():
    cfg = cases.layout_cfg()
    sd = arch.make_state_dict(arch.unet1d_specs(cfg), cases.WEIGHT_SEED_LAYOUT)
    g, obj_embed, x, t = cases.layout_step_inputs(cases.LAYOUT_CASE, cfg)
    G = gold("layout_n8.pt")
    with torch.no_grad():
        o = orc.unet1d_forward(sd, cfg, x, obj_embed, g.triples, t)
    assert o.abs().max() > 0.1, "vacuous parity (zero-initialised output)"
    assert max(rel_err(o, G["step"])) < TOL
    g, obj_embed, x_T, noises = cases.layout_chain_inputs(cases.LAYOUT_CASE, cfg, cases.LAYOUT_CHAIN_STEPS)
    with torch.no_grad():
        c = orc.layout_chain(sd, cfg, obj_embed, g.triples, x_T, noises, cases.LAYOUT_CHAIN_STEPS)
    assert max(rel_err(c, G["chain"])) < TOL


def test_shape_oracle_vs_reference():
    cfg = cases.shape_cfg()
    sd = arch.make_state_dict(arch.unet3d_specs(cfg), cases.WEIGHT_SEED_SHAPE)
    g, uc, x, t = cases.shape_step_inputs(cases.SHAPE_CASE, cfg)
    G = gold("shape.pt")
    with torch.no_grad():
        o = orc.unet3d_forward(sd, cfg, x, uc, g.triples, t)
    assert o.abs().max() > 0.1
    assert max(rel_err(o, G["step"])) < TOL


def test_schedules_closed_form():
    s = orc.DDIMSchedule(100)
    assert list(s.ddim_timesteps[:3]) == [1, 11, 21] and s.ddim_timesteps[-1] == 991 and len(s.ddim_timesteps) == 100
    assert s.alphas_prev[0] == s.alphas[0] * 0 + s.alphas_prev[0] and (s.sigmas == 0).all()
    d = orc.DDPMSchedule(time_num=1000)
    assert d.tables().shape == (5, 1000) and torch.isfinite(d.tables()).all()
    # x_{t-1} at t = 0 carries no noise (diffusion_ddpm.py:304-305)
    x = torch.randn(4, 8)
    e = torch.randn(4, 8)
    assert torch.equal(orc.ddpm_update(d, x, e, 0, torch.randn(4, 8)), orc.ddpm_update(d, x, e, 0, torch.zeros(4, 8)))


def test_vqvae_decode_oracle_vs_reference():
    """VQVAE.decode_no_quant (quantize -> post_quant_conv -> Decoder3D) against the reference's output for
    the seeded case of oracle/gen_golden_vqvae.py; one object is decoded here (objects are independent)."""
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))["cases"]["vqvae_decode_no_quant"]
    assert pin["rel_l2"] < 1e-6 and pin["indices_equal"] and pin["quant_equal"]
    cfg = cases.vqvae_cfg()
    specs = arch.vqvae_decode_specs(cfg)
    assert arch.count_params(specs) == pin["params"]
    sd = arch.make_state_dict(specs, cases.WEIGHT_SEED_VQVAE)
    z = cases.vqvae_inputs()[:1]
    G = gold("vqvae_decode.pt")
    quant, idx = orc.vq_quantize(sd, z)
    assert torch.equal(idx.to(torch.int32), G["indices"][: idx.numel()])          # integer work: bit-exact
    assert torch.equal(quant, G["quant"][:1])
    with torch.no_grad():
        dec = orc.vqvae_decode_no_quant(sd, cfg, z)
    assert dec.shape == (1, 1, 64, 64, 64)
    assert max(rel_err(dec[:, :, ::2, ::2, ::2], G["dec_sub"][:1])) < TOL


def test_scene_encoder_oracle_vs_reference():
    """what Sg2ScDiffModel.sample computes before the chains (init_encoder, manipulate, rel_s_mlp) against the
    outputs of the reference's own methods (oracle/gen_golden_scene.py)."""
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))["cases"]["scene_encode"]
    assert pin["rel_l2"] < 1e-6
    cfg = cases.scene_cfg()
    specs = arch.scene_encoder_specs(cfg)
    assert arch.count_params(specs) == pin["detail"]["params"]
    sd = arch.make_state_dict(specs, cases.WEIGHT_SEED_SCENE)
    g, objs, text, rel = cases.scene_inputs()
    G = gold("scene_encode.pt")
    with torch.no_grad():
        out = orc.scene_encode(sd, cfg, objs, g.triples, text, rel)
    for k in ("obj_embed", "latent", "uc_s", "c_s"):
        assert out[k].shape == G[k].shape, k
        assert max(rel_err(out[k], G[k])) < TOL, k
    assert out["uc_s"].shape == (8, 1, 1280) and out["obj_embed"].shape == (8, 640)


def test_vqvae_encode_oracle_vs_reference():
    """SURVEY 8f-3 (oracle stage): VQVAE.encode_no_quant (Encoder3D -> quant_conv) against the reference's output for the
    seeded SDF volume of oracle/gen_golden_vqvae.py."""
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))["cases"]["vqvae_encode_no_quant"]
    assert pin["rel_l2"] < 1e-6
    cfg = cases.vqvae_cfg()
    specs = arch.vqvae_encode_specs(cfg)
    assert arch.count_params(specs) == pin["params"]
    sd = arch.make_state_dict(specs, cases.WEIGHT_SEED_VQVAE + 1)
    x = cases.vqvae_sdf_inputs()
    with torch.no_grad():
        z = orc.vqvae_encode_no_quant(sd, cfg, x)
    G = gold("vqvae_encode.pt")
    assert z.shape == (1, 3, 16, 16, 16) == G["z"].shape
    assert max(rel_err(z, G["z"])) < TOL


def test_gcn_batch_statistics_oracle_matches_the_reference_in_train_mode():
    """oracle.graph_triple_conv_net(batch_stats=True) against the reference's GraphTripleConvNet under .train()
    (tests/golden/gcn_train.pt, oracle/gen_golden_train.py)."""
    from oracle import gen_golden_train as gt
    G = gold("gcn_train.pt")
    gcfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    for name, n, t, seed in gt.TRAIN_CASES:
        g, obj, pred = gt.inputs(n, t, seed, gcfg)
        edges, _ = orc.edges_of(g.triples)
        with torch.no_grad():
            o_obj, o_pred = orc.graph_triple_conv_net(sd, "", obj, pred, edges, batch_stats=True)
        assert torch.equal(o_obj, G[name]["obj"]) and torch.equal(o_pred, G[name]["pred"]), name


def test_gcn_backward_oracle_matches_the_reference_autograd():
    """oracle.gcn_backward (autograd over the oracle's forward restatement) against the reference's GraphTripleConvNet under .train()
    differentiated by torch autograd (tests/golden/gcn_bwd.pt, oracle/gen_golden_gcn_bwd.py): input gradients, parameter gradients and
    the BatchNorm1d buffers after the forward."""
    from oracle import gcn_backward, gen_golden_gcn_bwd as gb
    G = gold("gcn_bwd.pt")
    sd = gb.state_dict()
    for name, n, t, seed, with_pred in gb.CASES:
        g, obj, pred, d_obj, d_pred = gb.inputs(n, t, seed, with_pred)
        edges, _ = orc.edges_of(g.triples)
        o, p, gi, gp, grads, track = gcn_backward.graph_triple_conv_net_backward(sd, obj, pred, edges, d_obj, d_pred, gb.CFG["num_layers"])
        ref = G[name]
        assert max(rel_err(o, ref["obj_out"])) < 1e-6 and max(rel_err(gi, ref["d_obj"])) < 1e-6 and max(rel_err(gp, ref["d_pred"])) < 1e-6
        assert set(grads) == set(ref["grads"]) and set(track) == set(ref["buffers"])
        for k, want in ref["grads"].items():
            if float(want.abs().max()) == 0.0:
                assert float(grads[k].abs().max()) == 0.0, k
            else:
                assert max(rel_err(grads[k], want)) < 1e-5, (name, k)
        for k, want in ref["buffers"].items():
            assert max(rel_err(track[k].float(), want.float())) < 1e-6, (name, k)


def test_layout_denoiser_train_mode_oracle_forward_and_backward_match_the_reference():
    """oracle.unet1d_forward(batch_stats=True) and autograd over it against the reference's UNet1DModel under .train() and ITS autograd
    (tests/golden/layout_train.pt, oracle/gen_golden_layout_train.py): the forward output bit for bit, and for every parameter the
    digest of its gradient (L2 norm, eight entries).  This is the parity contract the layout trunk's backward pass will be built
    against; no CUDA code is involved yet."""
    from oracle import gen_golden_layout_train as gt
    G = gold("layout_train.pt")
    lcfg = cases.layout_cfg()
    sd = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
    g, obj_embed, x, t, noise = gt.inputs(lcfg)
    out, loss, grads = gt.oracle_grads(sd, lcfg, g, obj_embed, x, t, noise)
    assert torch.equal(out, G["out"]) and abs(loss - G["loss"]) < 1e-6
    assert set(G["no_grad"]) == {"box_graph_cov.gconvs.4.linear_projection_pred.weight", "box_graph_cov.gconvs.4.linear_projection_pred.bias"}
    scale, noisy = G["grad_scale"], set(G["noise_level"])
    assert len(noisy) == 90 and any(k.endswith("attn1.to_q.weight") for k in noisy)     # softmax over one token: q / k get no gradient
    for k, d in G["grads"].items():
        got = gt.digest(grads[k])
        if k in noisy:
            assert float(grads[k].abs().max()) < 1e-5 * scale, k
            continue
        assert abs(got["norm"] - d["norm"]) <= 1e-3 * d["norm"], k
        assert float((got["samples"] - d["samples"]).abs().max()) <= 1e-3 * max(float(d["samples"].abs().max()), 1e-3 * d["norm"]), k


def test_shape_denoiser_train_mode_oracle_forward_and_backward_match_the_reference():
    """oracle.unet3d_forward(batch_stats=True) and autograd over it against the reference's UNet3DModel under .train() and ITS autograd
    on three objects with one timestep each (tests/golden/shape_train.pt, oracle/gen_golden_shape_train.py): the forward output bit
    for bit; per parameter the digest of its gradient, 1e-3 of its norm plus a rounding floor scaled by the largest gradient of the
    model (most biases of this network carry gradients 1e-6 of the output convolution's).  The parity contract of the shape trunk's
    backward kernels (conv dgrad / wgrad, GroupNorm, attention, GEGLU); no CUDA code is involved yet."""
    from oracle import gen_golden_layout_train as gl, gen_golden_shape_train as gs
    G = gold("shape_train.pt")
    scfg = cases.shape_cfg()
    sd = arch.make_state_dict(arch.unet3d_specs(scfg), cases.WEIGHT_SEED_SHAPE)
    g, uc, x, t, noise = gs.inputs(scfg)
    out, loss, grads = gs.oracle_grads(sd, scfg, g, uc, x, t, noise)
    assert torch.equal(out, G["out"]) and abs(loss - G["loss"]) < 1e-6
    assert set(G["no_grad"]) == {"shape_code_graph_cov.gconvs.4.linear_projection_pred.weight", "shape_code_graph_cov.gconvs.4.linear_projection_pred.bias"}
    scale, noisy = G["grad_scale"], set(G["noise_level"])
    for k, d in G["grads"].items():
        if k in noisy:
            assert float(grads[k].abs().max()) < 1e-5 * scale, k
            continue
        got = gl.digest(grads[k])
        floor = 1e-7 * scale * grads[k].numel() ** 0.5
        assert abs(got["norm"] - d["norm"]) <= 1e-3 * d["norm"] + floor, (k, got["norm"], d["norm"])
        assert float((got["samples"] - d["samples"]).abs().max()) <= 1e-3 * float(d["samples"].abs().max()) + 1e-6 * scale, k
