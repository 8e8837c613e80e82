"""GPU: the training-side kernels of SURVEY 8f-3 (csrc/train.cu) against the reference's own arithmetic in torch:
q_sample (diffusion_ddpm.py:191-201, echo2shape.py:254-258), the diffusion losses (diffusion_ddpm.py:451-477, echo2shape.py:297-331)
and the optimizer step of scripts/train_3dfront.py:247-259 (clip_grad_norm_ + NaN scrub loop + torch.optim.AdamW)."""
import pytest
import torch

from echoscene_b200 import train
from util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def schedule(T=1000, b0=1e-4, b1=0.02):
    betas = torch.linspace(b0, b1, T, dtype=torch.float64)
    ac = torch.cumprod(1.0 - betas, 0)
    return ac.sqrt().float().to(DEV), (1.0 - ac).sqrt().float().to(DEV)


@pytest.mark.parametrize("shape", [(37, 8), (16, 3, 16, 16, 16), (1, 8), (0, 8)])
def test_q_sample_matches_the_reference_expression(shape):
    g = torch.Generator().manual_seed(3)
    x0, noise = torch.randn(shape, generator=g).to(DEV), torch.randn(shape, generator=g).to(DEV)
    t = torch.randint(0, 1000, (shape[0],), generator=g).to(DEV)
    a, b = schedule()
    got = train.q_sample(x0, t, noise, a, b)
    ex = (shape[0],) + (1,) * (len(shape) - 1)
    want = a[t].reshape(ex) * x0 + b[t].reshape(ex) * noise          # _extract(...) * x_start + _extract(...) * noise
    assert torch.equal(got, want), "q_sample must be bit-exact: two products and one sum in fp32"


def test_layout_loss_parts_match_the_reference_expression():
    g = torch.Generator().manual_seed(4)
    out, tgt = torch.randn(203, 8, generator=g).to(DEV), torch.randn(203, 8, generator=g).to(DEV)
    loss, parts = train.layout_diffusion_loss(out, tgt)
    ref = {"loss.size": ((tgt[:, 0:3] - out[:, 0:3]) ** 2).mean(dim=1).mean(), "loss.trans": ((tgt[:, 3:6] - out[:, 3:6]) ** 2).mean(dim=1).mean(),
           "loss.angle": ((tgt[:, 6:8] - out[:, 6:8]) ** 2).mean(dim=1).mean(), "loss.bbox": ((tgt - out) ** 2).mean(dim=1).mean()}
    for k, v in ref.items():
        assert abs(float(parts[k]) - float(v)) <= 1e-6 * abs(float(v)), k
    assert abs(float(loss) - float(((tgt - out) ** 2).mean(dim=1).mean())) <= 1e-6 * float(loss)


def test_shape_loss_matches_the_reference_expression():
    g = torch.Generator().manual_seed(5)
    out, tgt = torch.randn(16, 3, 16, 16, 16, generator=g).to(DEV), torch.randn(16, 3, 16, 16, 16, generator=g).to(DEV)
    t = torch.randint(0, 1000, (16,), generator=g).to(DEV)
    logvar, lvlb = torch.zeros(1000), torch.rand(1000, generator=g)
    loss, parts = train.shape_diffusion_loss(out, tgt, t, logvar, lvlb, 1.0, 0.0)
    ls = torch.nn.functional.mse_loss(tgt, out, reduction="none").mean([1, 2, 3, 4])
    want = (ls / torch.exp(logvar.to(DEV)[t]) + logvar.to(DEV)[t]).mean()
    assert abs(float(loss) - float(want)) <= 2e-6 * float(want)
    assert abs(float(parts["loss_vlb"]) - float((lvlb.to(DEV)[t] * ls).mean())) <= 2e-6 * float(parts["loss_vlb"])


def _reference_step(params, clip, opt, max_norm=5.0):
    """scripts/train_3dfront.py:250-258, literally"""
    torch.nn.utils.clip_grad_norm_(clip, max_norm)
    for group in opt.param_groups:
        for p in group["params"]:
            if p.grad is not None and p.requires_grad and torch.isnan(p.grad).any():
                p.grad[torch.isnan(p.grad)] = 0
    opt.step()


@pytest.mark.parametrize("nan_in", [None, "other", "clip"])
def test_fused_optimizer_step_matches_clip_scrub_adamw(nan_in):
    g = torch.Generator().manual_seed(6)
    shapes = [(224, 224, 3, 3, 3), (1280,), (512, 2048), (3,), (672, 1344)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    clip_idx = [0, 4]                                  # "the shape denoiser's parameters"
    opt = torch.optim.AdamW(ref_p, lr=1e-4)
    ours = train.FusedAdamW(our_p, lr=1e-4, clip_params=[our_p[i] for i in clip_idx], clip_max_norm=5.0)
    for step in range(3):
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            gr = torch.randn(a.shape, generator=g).to(DEV) * (30.0 if i in clip_idx else 1.0)    # the clip is active
            if step == 1 and nan_in == "other" and i == 1:
                gr[7] = float("nan")
            if step == 1 and nan_in == "clip" and i == 4:
                gr[3, 5] = float("nan")                # clip_grad_norm_ then turns EVERY clipped gradient into NaN -> all scrubbed to 0
            a.grad, b.grad = gr.clone(), gr.clone()
        _reference_step(ref_p, [ref_p[i] for i in clip_idx], opt)
        ours.step()
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            assert torch.isfinite(b).all()
            assert_close(b.detach(), a.detach(), 2e-6, f"step {step} parameter {i} ({nan_in})")
            assert_close(b.grad, a.grad, 1e-6, f"step {step} gradient left in place {i} ({nan_in})")
    st = opt.state[ref_p[4]]
    assert_close(ours.exp_avg[4], st["exp_avg"], 2e-6, "exp_avg")
    assert_close(ours.exp_avg_sq[4], st["exp_avg_sq"], 2e-6, "exp_avg_sq")
    info = ours.info()
    assert info["steps"] == 3 and info["parameters"] == sum(p.numel() for p in our_p)
    if nan_in == "other":
        assert info["nan_gradients_scrubbed"] == 1
    if nan_in == "clip":
        assert info["nan_gradients_scrubbed"] == sum(our_p[i].numel() for i in clip_idx)


def test_fused_optimizer_refuses_cpu_and_missing_gradients():
    from echoscene_b200._lib import EchoError
    with pytest.raises(EchoError):
        train.FusedAdamW([torch.nn.Parameter(torch.zeros(4))])
    p = torch.nn.Parameter(torch.zeros(4, device=DEV))
    with pytest.raises(EchoError, match="gradient"):
        train.FusedAdamW([p]).step()


def test_gcn_forward_on_batch_statistics_matches_the_reference_in_train_mode():
    """echo_gcn_forward_train (BatchNorm1d on the statistics of the batch, model/layers.py:29-30 under model.train()) against the
    reference's GraphTripleConvNet in .train() mode: a single scene (few-row kernels) and a collated batch (tiled GEMM path)."""
    from echoscene_b200 import arch, modules
    from echoscene_b200._lib import EchoError
    from oracle import cases, echoscene_oracle as orc, gen_golden_train as gt
    from util import FP32_TOL, gold
    G = gold("gcn_train.pt")
    gcfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    m = modules.GraphTripleConvNet(gcfg.input_dim_obj, gcfg.input_dim_pred, num_layers=gcfg.num_layers, hidden_dim=gcfg.hidden_dim,
                                   residual=True, pooling="avg", mlp_normalization="batch", output_dim=gcfg.output_dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    for name, n, t, seed in gt.TRAIN_CASES:
        g, obj, pred = gt.inputs(n, t, seed, gcfg)
        edges, _ = orc.edges_of(g.triples)
        o, p = m.forward_batch_stats(obj.to(DEV), pred.to(DEV), edges.to(DEV))
        assert_close(o, G[name]["obj"], FP32_TOL, f"train-mode GCN {name}: nodes")
        assert_close(p, G[name]["pred"], FP32_TOL, f"train-mode GCN {name}: predicates")
        e, _ = m(obj.to(DEV), pred.to(DEV), edges.to(DEV))                       # the eval forward still works on the same handle
        assert float((e.cpu() - G[name]["obj"]).abs().max()) > 1e-2
    m.train()                                                                   # train(): forward IS the batch-statistics forward
    o2, p2 = m(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    assert torch.equal(o2, o) and torch.equal(p2, p)


def _check_grad(got, want, tol, what):
    """A bias whose only way to the loss is a BatchNorm1d on batch statistics has a mathematically zero gradient (the normalisation
    removes the column mean): the Linear biases inside the MLPs, and linear_projection_pred.bias of a layer whose predicate output is
    read by the next layer only.  Both sides hold rounding noise of the order of 1e-6 there (real gradients here are >= 1e-2):
    compared on an absolute scale."""
    if 0.0 < float(want.abs().max()) < 1e-5:
        assert float(got.abs().max()) < 1e-5, what
    elif float(want.abs().max()) == 0.0:
        assert float(got.abs().max()) == 0.0, what
    else:
        assert_close(got, want, tol, what)


def _gcn_bwd_module():
    from echoscene_b200 import modules
    from oracle import gen_golden_gcn_bwd as gb
    c = gb.CFG
    m = modules.GraphTripleConvNet(c["input_dim_obj"], c["input_dim_pred"], num_layers=c["num_layers"], hidden_dim=c["hidden_dim"],
                                   residual=True, pooling="avg", mlp_normalization="batch", output_dim=c["output_dim"])
    m.load_state_dict(gb.state_dict(), strict=True)
    return m.to(DEV), gb


def test_gcn_backward_matches_the_reference_autograd():
    """echo_gcn_train_forward / _backward against the reference's GraphTripleConvNet under .train() differentiated by torch autograd
    (tests/golden/gcn_bwd.pt, oracle/gen_golden_gcn_bwd.py): outputs, input gradients, every parameter gradient, and the BatchNorm1d
    buffers after the forward -- a single scene with both cotangents, and a collated batch whose predicate output has none."""
    from oracle import echoscene_oracle as orc
    from util import gold
    G = gold("gcn_bwd.pt")
    tol = 2e-4
    for name, n, t, seed, with_pred in _gcn_bwd_module()[1].CASES:
        m, gb = _gcn_bwd_module()
        tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)
        g, obj, pred, d_obj, d_pred = gb.inputs(n, t, seed, with_pred)
        edges, _ = orc.edges_of(g.triples)
        o, p = tr.forward(obj.to(DEV), pred.to(DEV), edges.to(DEV))
        assert_close(o, G[name]["obj_out"], tol, f"{name}: obj_out")
        assert_close(p, G[name]["pred_out"], tol, f"{name}: pred_out")
        gi, gp = tr.backward(d_obj.to(DEV), d_pred.to(DEV) if d_pred is not None else None)
        assert_close(gi, G[name]["d_obj"], tol, f"{name}: d_obj_vecs")
        assert_close(gp, G[name]["d_pred"], tol, f"{name}: d_pred_vecs")
        for k, want in G[name]["grads"].items():
            _check_grad(dict(m.named_parameters())[k].grad, want, tol, f"{name}: grad of {k}")
        for k, want in G[name]["buffers"].items():
            got = dict(m.named_buffers())[k]
            if want.dtype == torch.int64:
                assert torch.equal(got.cpu(), want), k
            else:
                assert_close(got, want, tol, f"{name}: buffer {k} after the forward")


def test_gcn_backward_accumulates_is_deterministic_and_reads_parameters_in_place():
    from oracle import echoscene_oracle as orc
    m, gb = _gcn_bwd_module()
    name, n, t, seed, with_pred = gb.CASES[0]
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)
    g, obj, pred, d_obj, d_pred = gb.inputs(n, t, seed, True)
    edges, _ = orc.edges_of(g.triples)
    args = (obj.to(DEV), pred.to(DEV), edges.to(DEV))
    rm0 = {k: v.clone() for k, v in m.named_buffers()}
    o1, _ = tr.forward(*args)
    gi1, gp1 = tr.backward(d_obj.to(DEV), d_pred.to(DEV))
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    for k, v in m.named_buffers():          # second pass from the same buffers: bit-identical results (no atomics anywhere)
        v.copy_(rm0[k])
    handle = tr._handle.value
    o2, _ = tr.forward(*args)
    gi2, gp2 = tr.backward(d_obj.to(DEV), d_pred.to(DEV))
    assert torch.equal(o1, o2) and torch.equal(gi1, gi2) and torch.equal(gp1, gp2)
    for k, p in m.named_parameters():       # .grad accumulates as autograd does: exactly twice the first pass
        assert torch.equal(p.grad, 2 * g1[k]), k
    # an optimizer step writes the parameters in place: the next forward sees them without a rebuild
    tr.zero_grad()
    with torch.no_grad():
        dict(m.named_parameters())["gconvs.1.net2.3.weight"].mul_(1.5)
    o3, _ = tr.forward(*args)
    assert tr._handle.value == handle and not torch.equal(o3, o1)
    # the input gradients are optional; backward before forward on another graph is refused
    assert tr.backward(d_obj.to(DEV), None, need_input_grads=False) == (None, None)
    from echoscene_b200._lib import EchoError
    with pytest.raises(EchoError, match="more than one row"):
        tr.forward(args[0][:1], args[1][:1], torch.zeros(1, 2, dtype=torch.int64, device=DEV))


def test_gcn_training_iteration_as_one_cuda_graph():
    """capture(): forward + backward replayed as one CUDA graph give bit-identical outputs and gradients to the launch-by-launch
    calls, and follow new input values copied into the static buffers."""
    from oracle import echoscene_oracle as orc
    m, gb = _gcn_bwd_module()
    name, n, t, seed, _ = gb.CASES[0]
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)
    g, obj, pred, d_obj, d_pred = gb.inputs(n, t, seed, True)
    edges = orc.edges_of(g.triples)[0].to(DEV)
    obj, pred, d_obj, d_pred = obj.to(DEV), pred.to(DEV), d_obj.to(DEV), d_pred.to(DEV)
    bufs0 = {k: v.clone() for k, v in m.named_buffers()}
    o, p = tr.forward(obj, pred, edges)
    gi, gp = tr.backward(d_obj, d_pred)
    want = {k: q.grad.clone() for k, q in m.named_parameters()}
    bufs1 = {k: v.clone() for k, v in m.named_buffers()}
    for k, v in m.named_buffers():
        v.copy_(bufs0[k])
    it = tr.capture(obj, pred, edges, d_obj, d_pred)
    for k, v in m.named_buffers():            # the warm-up iterations of capture() advanced the running statistics
        v.copy_(bufs0[k])
    tr.zero_grad()
    o2, p2, gi2, gp2 = it.replay()
    assert torch.equal(o, o2) and torch.equal(p, p2) and torch.equal(gi, gi2) and torch.equal(gp, gp2)
    for k, q in m.named_parameters():
        assert torch.equal(q.grad, want[k]), k
    for k, v in m.named_buffers():
        assert torch.equal(v, bufs1[k]), k
    tr.zero_grad()
    o3, _, gi3, _ = it.replay(obj=2 * obj, d_obj=-d_obj)       # new values through the static buffers
    o4, _ = tr.forward(2 * obj, pred, edges)
    assert torch.equal(o3, o4) and not torch.equal(o3, o)


def test_gcn_backward_many_rows_two_stage_kernels():
    """A collated batch of 70 scenes (560 nodes, 2240 triples) at the golden case's small widths: the row counts at which BatchNorm runs
    as two-stage kernels (partials per 256-row block, merged in block order) and the weight gradient's reduction is split into chunks,
    against autograd over the oracle's forward restatement in FP64 (at these widths fp32 stays well conditioned: 1e-3)."""
    from echoscene_b200 import synth
    from oracle import echoscene_oracle as orc, gcn_backward
    m, gb = _gcn_bwd_module()
    sd = gb.state_dict()
    g = synth.batch_scene_graphs([synth.make_scene_graph(8, 32, 300 + i) for i in range(70)])
    gen = torch.Generator().manual_seed(12)
    n, t = g.n_nodes, g.triples.shape[0]
    c = gb.CFG
    obj, pred = torch.randn(n, c["input_dim_obj"], generator=gen), torch.randn(t, c["input_dim_pred"], generator=gen)
    d_obj, d_pred = torch.randn(n, c["output_dim"], generator=gen), torch.randn(t, c["input_dim_pred"], generator=gen)
    edges, _ = orc.edges_of(g.triples)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    w = gcn_backward.graph_triple_conv_net_backward(sd64, obj.double(), pred.double(), edges, d_obj.double(), d_pred.double(), c["num_layers"])
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)
    o, p = tr.forward(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    gi, gp = tr.backward(d_obj.to(DEV), d_pred.to(DEV))
    tol = 1e-3
    assert_close(o, w[0].float(), tol, "obj_out"); assert_close(p, w[1].float(), tol, "pred_out")
    assert_close(gi, w[2].float(), tol, "d_obj_vecs"); assert_close(gp, w[3].float(), tol, "d_pred_vecs")
    scale = max(float(v.abs().max()) for v in w[4].values())
    params, bufs = dict(m.named_parameters()), dict(m.named_buffers())
    for k, want in w[4].items():
        if float(want.abs().max()) < 1e-6 * scale:
            assert float(params[k].grad.abs().max()) < 1e-5 * scale, k
        else:
            assert_close(params[k].grad, want.float(), tol, f"grad of {k}")
    for k, want in w[5].items():
        if want.dtype != torch.int64:
            assert_close(bufs[k], want.float(), tol, f"buffer {k}")


def test_gcn_backward_at_the_denoiser_width_single_scene_tight():
    """box_graph_cov's widths (768 / 128 / hidden 256 / out 1280; two layers) on ONE scene (8 nodes, 32 triples): with a few
    thousand activations no ReLU mask sits within rounding of zero and fp32 is well conditioned (the fp32 oracle is 2e-6 from fp64),
    so every gradient is held to 1e-4 against the fp32 oracle.  Covers the multi-block paths of the dgrad / wgrad kernels (K = 1664)."""
    import dataclasses
    from echoscene_b200 import arch, modules, synth
    from oracle import cases, echoscene_oracle as orc, gcn_backward
    gcfg = dataclasses.replace(cases.layout_cfg().gcn(), num_layers=2)
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    m = modules.GraphTripleConvNet(gcfg.input_dim_obj, gcfg.input_dim_pred, num_layers=2, hidden_dim=gcfg.hidden_dim, residual=True,
                                   pooling="avg", mlp_normalization="batch", output_dim=gcfg.output_dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    g = synth.make_scene_graph(8, 32, 2)
    gen = torch.Generator().manual_seed(10)
    obj, pred = torch.randn(8, gcfg.input_dim_obj, generator=gen), torch.randn(32, gcfg.input_dim_pred, generator=gen)
    d_obj, d_pred = torch.randn(8, gcfg.output_dim, generator=gen), torch.randn(32, gcfg.input_dim_pred, generator=gen)
    edges, _ = orc.edges_of(g.triples)
    w = gcn_backward.graph_triple_conv_net_backward(sd, obj, pred, edges, d_obj, d_pred, 2)
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=8, max_triples=32)
    o, p = tr.forward(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    gi, gp = tr.backward(d_obj.to(DEV), d_pred.to(DEV))
    tol = 1e-4
    assert_close(o, w[0], tol, "obj_out"); assert_close(p, w[1], tol, "pred_out")
    assert_close(gi, w[2], tol, "d_obj_vecs"); assert_close(gp, w[3], tol, "d_pred_vecs")
    scale = max(float(v.abs().max()) for v in w[4].values())
    for k, want in w[4].items():
        got = dict(m.named_parameters())[k].grad
        if float(want.abs().max()) < 1e-6 * scale:
            assert float(got.abs().max()) < 1e-5 * scale, k
        else:
            assert_close(got, want, tol, f"grad of {k}")


def test_gcn_backward_at_the_denoiser_width_against_the_oracle():
    """The layout denoiser's box_graph_cov (5 layers, the widths of config/full_mp.yaml: 768 / 128 / hidden 256 / out 1280) on a ragged
    collated batch of 12 scenes (162 nodes, 552 triples): CUDA backward against autograd over the oracle's forward restatement
    (oracle/gcn_backward.py), computed here on the CPU.

    Tolerance: at these widths the GRADIENTS are ill-conditioned in fp32.  The oracle's own fp32 autograd sits between 5e-4 and
    1.2e-2 (rel-L2) from the same computation in fp64, depending on nothing but the summation order of the host's GEMM (1 thread
    against 16 threads, one CPU model against another; one layer is enough to see 7e-5 against 2.5e-3): a pre-activation within
    rounding of zero flips its ReLU mask, and BatchNorm's backward spreads that flip over the whole column.  So the forward outputs
    are held to 1e-3 against the fp32 oracle, every gradient is measured against the FP64 oracle and must stay within 3e-2 of it --
    the band the reference's own fp32 arithmetic occupies (this kernel measured 2e-5 .. 1.4e-2 over its versions, moving with the
    GEMM's summation order exactly as the CPU oracle does) -- and the measured distances of both are printed.  The tight check of
    the backward's structure is the golden test above (2e-4 at small widths, where no mask flips)."""
    from echoscene_b200 import arch, modules, synth
    from oracle import cases, echoscene_oracle as orc, gcn_backward
    from util import rel_err
    gcfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    m = modules.GraphTripleConvNet(gcfg.input_dim_obj, gcfg.input_dim_pred, num_layers=gcfg.num_layers, hidden_dim=gcfg.hidden_dim,
                                   residual=True, pooling="avg", mlp_normalization="batch", output_dim=gcfg.output_dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    g = synth.batch_scene_graphs([synth.make_scene_graph(8 + i, 24 + 4 * i, 50 + i) for i in range(12)])   # ragged scenes, > 64 nodes
    gen = torch.Generator().manual_seed(8)
    n, t = g.n_nodes, g.triples.shape[0]
    obj, pred = torch.randn(n, gcfg.input_dim_obj, generator=gen), torch.randn(t, gcfg.input_dim_pred, generator=gen)
    d_obj, d_pred = torch.randn(n, gcfg.output_dim, generator=gen), torch.randn(t, gcfg.input_dim_pred, generator=gen)
    edges, _ = orc.edges_of(g.triples)
    w32 = gcn_backward.graph_triple_conv_net_backward(sd, obj, pred, edges, d_obj, d_pred, gcfg.num_layers)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    w64 = gcn_backward.graph_triple_conv_net_backward(sd64, obj.double(), pred.double(), edges, d_obj.double(), d_pred.double(),
                                                      gcfg.num_layers)
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)
    o, p = tr.forward(obj.to(DEV), pred.to(DEV), edges.to(DEV))
    gi, gp = tr.backward(d_obj.to(DEV), d_pred.to(DEV))
    assert_close(o, w32[0], 1e-3, "obj_out"); assert_close(p, w32[1], 1e-3, "pred_out")

    worst = [0.0, 0.0]
    scale = max(float(v.abs().max()) for v in w64[4].values())
    def check(got, a32, a64, what):
        if float(a64.abs().max()) < 1e-5:                      # mathematically zero (see _check_grad): rounding noise of the sums
            assert float(got.abs().max()) < 1e-5 * scale, f"{what}: {float(got.abs().max()):.2e} against gradients of scale {scale:.1e}"
            return
        mine, ref = rel_err(got, a64)[1], rel_err(a32, a64)[1]
        worst[0], worst[1] = max(worst[0], mine), max(worst[1], ref)
        assert mine <= 3e-2, f"{what}: rel-L2 {mine:.2e} from the fp64 oracle; the fp32 oracle is {ref:.2e} from it"
    check(gi, w32[2], w64[2], "d_obj_vecs"); check(gp, w32[3], w64[3], "d_pred_vecs")
    params, bufs = dict(m.named_parameters()), dict(m.named_buffers())
    for k in w32[4]:
        check(params[k].grad, w32[4][k], w64[4][k], f"grad of {k}")
    for k, want in w32[5].items():
        if want.dtype != torch.int64:
            assert_close(bufs[k], want, 1e-3, f"buffer {k}")
    print(f"gradients vs the fp64 oracle: worst rel-L2 {worst[0]:.2e} (CUDA), {worst[1]:.2e} (fp32 oracle)")
