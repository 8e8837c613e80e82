"""Training-mode golden vectors: the reference's GraphTripleConvNet under .train() -- BatchNorm1d on batch statistics
(model/layers.py:29-30) -- imported in place from /root/reference (build container only), on the round-1 GCN case (N = 8, T = 32)
and on a collated batch (N = 40, T = 160).  Pins oracle.graph_triple_conv_net(batch_stats=True) and writes
tests/golden/gcn_train.pt.  Usage: python oracle/gen_golden_train.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from echoscene_b200 import arch, synth  # noqa: E402
from oracle import cases, echoscene_oracle as orc, ref_import  # noqa: E402

TRAIN_CASES = [("n8", 8, 32, 1), ("batch_n40", 40, 160, 7)]


def inputs(n, t, seed, cfg):
    if n <= 16:
        g = synth.make_scene_graph(n, t, seed)
    else:
        g = synth.batch_scene_graphs([synth.make_scene_graph(8, 32, seed + i) for i in range(n // 8)])
    gen = torch.Generator().manual_seed(seed + 100)
    return g, torch.randn(g.n_nodes, cfg.input_dim_obj, generator=gen), torch.randn(g.triples.shape[0], cfg.input_dim_pred, generator=gen)


def tables_main():
    """The q_sample / loss tables the two branches build for TRAINING, from the reference's own constructors: GaussianDiffusion.__init__
    (diffusion_ddpm.py:119-166, through DiffusionPoint with a stub denoiser) and EchoToShape.register_schedule (echo2shape.py:173-226,
    called unbound on a bare namespace: it only reads device / v_posterior / parameterization).  -> tests/golden/train_tables.pt, pins
    echoscene_b200.train.layout_train_tables / shape_train_tables bit for bit."""
    import importlib
    import types
    from echoscene_b200 import train
    ref = ref_import.load()
    dp = ref.DiffusionPoint(torch.nn.Identity(), config={}, schedule_type="linear", beta_start=1e-4, beta_end=0.02, time_num=1000,
                            loss_type="mse", model_mean_type="eps", model_var_type="fixedsmall", loss_separate=True, loss_iou=False,
                            iou_type="obb", train_stats_file=None)
    g = dp.diffusion
    e2s = importlib.import_module("model.networks.diffusion_shape.echo2shape")
    ns = types.SimpleNamespace(device="cpu", v_posterior=0.0, parameterization="eps")
    e2s.EchoToShape.register_schedule(ns, timesteps=1000, linear_start=0.00085, linear_end=0.012)
    out = {"layout": {"sqrt_alphas_cumprod": g.sqrt_alphas_cumprod.clone(), "sqrt_one_minus_alphas_cumprod": g.sqrt_one_minus_alphas_cumprod.clone()},
           "shape": {"sqrt_alphas_cumprod": ns.sqrt_alphas_cumprod.clone(), "sqrt_one_minus_alphas_cumprod": ns.sqrt_one_minus_alphas_cumprod.clone(),
                     "lvlb_weights": ns.lvlb_weights.clone()}}
    a, b = train.layout_train_tables(1000, 1e-4, 0.02)
    t = train.shape_train_tables(1000, 0.00085, 0.012)
    ok = (torch.equal(a, out["layout"]["sqrt_alphas_cumprod"]) and torch.equal(b, out["layout"]["sqrt_one_minus_alphas_cumprod"])
          and all(torch.equal(t[k], v) for k, v in out["shape"].items()))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "train_tables.pt"))
    print(f"training tables (layout q_sample, shape q_sample + lvlb_weights): bit-equal to the reference's constructors: {ok}")
    assert ok


def main():
    ref = ref_import.load()
    gcfg = cases.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), cases.WEIGHT_SEED_GCN)
    out, worst = {}, 0.0
    for name, n, t, seed in TRAIN_CASES:
        net = ref.GraphTripleConvNet(input_dim_obj=gcfg.input_dim_obj, input_dim_pred=gcfg.input_dim_pred, num_layers=gcfg.num_layers,
                                     hidden_dim=gcfg.hidden_dim, residual=True, pooling="avg", mlp_normalization="batch",
                                     output_dim=gcfg.output_dim)
        net.load_state_dict(sd, strict=True)
        net.train()
        g, obj, pred = inputs(n, t, seed, gcfg)
        edges, _ = orc.edges_of(g.triples)
        with torch.no_grad():
            r_obj, r_pred = net(obj, pred, edges)
            o_obj, o_pred = orc.graph_triple_conv_net(sd, "", obj, pred, edges, batch_stats=True)
            e_obj, _ = orc.graph_triple_conv_net(sd, "", obj, pred, edges, batch_stats=False)
        worst = max(worst, float((r_obj - o_obj).abs().max()), float((r_pred - o_pred).abs().max()))
        assert float((r_obj - e_obj).abs().max()) > 1e-2, "batch statistics must change the result (else the case is vacuous)"
        out[name] = {"obj": r_obj, "pred": r_pred}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "gcn_train.pt"))
    tables_main()
    print(f"GraphTripleConvNet under .train(): oracle(batch_stats=True) vs reference max-abs {worst:.3e} over {len(TRAIN_CASES)} cases")


if __name__ == "__main__":
    main()
