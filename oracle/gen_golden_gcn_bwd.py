"""Backward-pass golden vectors: the reference's GraphTripleConvNet under .train(), imported in place from
/root/reference (build container only), differentiated by torch autograd for random cotangents of both outputs -- gradients of the
inputs, of every parameter, and the BatchNorm1d buffers after the forward.  Small widths keep the fixture small; the layer structure
is the reference's (5 layers would only repeat it: 2 here, the last with its own output width, residual projections on).
Pins oracle.gcn_backward and writes tests/golden/gcn_bwd.pt.  Usage: python oracle/gen_golden_gcn_bwd.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from echoscene_b200 import arch, synth  # noqa: E402
from oracle import echoscene_oracle as orc, gcn_backward, ref_import  # noqa: E402

CFG = dict(input_dim_obj=48, input_dim_pred=16, num_layers=2, hidden_dim=64, output_dim=32)
CASES = [("n8", 8, 32, 3, True), ("batch_n40_no_pred_cotangent", 40, 160, 9, False)]
WEIGHT_SEED = 31


def state_dict():
    g = arch.GCNConfig(CFG["input_dim_obj"], CFG["input_dim_pred"], CFG["num_layers"], CFG["hidden_dim"], CFG["output_dim"], True, "avg",
                       "batch")
    return arch.make_state_dict(arch.gcn_specs(g), WEIGHT_SEED)


def inputs(n, t, seed, with_pred):
    if n <= 16:
        g = synth.make_scene_graph(n, t, seed)
    else:
        g = synth.batch_scene_graphs([synth.make_scene_graph(8, 32, seed + i) for i in range(n // 8)])
    gen = torch.Generator().manual_seed(seed + 200)
    T = g.triples.shape[0]
    obj, pred = torch.randn(g.n_nodes, CFG["input_dim_obj"], generator=gen), torch.randn(T, CFG["input_dim_pred"], generator=gen)
    d_obj = torch.randn(g.n_nodes, CFG["output_dim"], generator=gen)
    d_pred = torch.randn(T, CFG["input_dim_pred"], generator=gen) if with_pred else None
    return g, obj, pred, d_obj, d_pred


def main():
    ref = ref_import.load()
    sd = state_dict()
    out, worst = {}, 0.0
    for name, n, t, seed, with_pred in CASES:
        net = ref.GraphTripleConvNet(input_dim_obj=CFG["input_dim_obj"], input_dim_pred=CFG["input_dim_pred"], num_layers=CFG["num_layers"],
                                     hidden_dim=CFG["hidden_dim"], residual=True, pooling="avg", mlp_normalization="batch",
                                     output_dim=CFG["output_dim"])
        net.load_state_dict(sd, strict=True)
        net.train()
        g, obj, pred, d_obj, d_pred = inputs(n, t, seed, with_pred)
        edges, _ = orc.edges_of(g.triples)
        o, p = obj.clone().requires_grad_(True), pred.clone().requires_grad_(True)
        r_obj, r_pred = net(o, p, edges)
        loss = (r_obj * d_obj).sum() + ((r_pred * d_pred).sum() if d_pred is not None else 0.0)
        loss.backward()
        # a parameter the loss does not reach keeps .grad = None in torch (the last layer's linear_projection_pred when the predicate
        # output has no cotangent): recorded as zeros, which is what an accumulating backward leaves in a zeroed buffer
        ref_grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in net.named_parameters()}
        ref_bufs = {k: v.detach().clone() for k, v in net.named_buffers()}
        o_obj, o_pred, g_obj, g_pred, grads, track = gcn_backward.graph_triple_conv_net_backward(sd, obj, pred, edges, d_obj, d_pred,
                                                                                               CFG["num_layers"])
        def dev(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        w = max(dev(o_obj, r_obj.detach()), dev(o_pred, r_pred.detach()), dev(g_obj, o.grad), dev(g_pred, p.grad))
        assert set(grads) == set(ref_grads), (sorted(set(grads) ^ set(ref_grads)))
        for k, v in ref_grads.items():
            w = max(w, dev(grads[k], v))
        for k, v in track.items():
            w = max(w, dev(v.float(), ref_bufs[k].float()))
        assert set(track) == set(ref_bufs)
        worst = max(worst, w)
        out[name] = {"obj_out": r_obj.detach(), "pred_out": r_pred.detach(), "d_obj": o.grad.clone(), "d_pred": p.grad.clone(),
                     "grads": ref_grads, "buffers": ref_bufs}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "gcn_bwd.pt"))
    print(f"GraphTripleConvNet backward: oracle (autograd over the restated forward) vs the reference's autograd, max relative "
          f"deviation {worst:.3e} over {len(CASES)} cases; fixture {os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'gcn_bwd.pt')) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
