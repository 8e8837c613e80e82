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
    m.train()
    with pytest.raises(EchoError):                                              # forward() keeps refusing train(): no autograd tape
        m(obj.to(DEV), pred.to(DEV), edges.to(DEV))
