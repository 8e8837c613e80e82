"""Forward + backward of box_graph_cov (the layout denoiser's GraphTripleConvNet, 5 layers, config/full_mp.yaml widths) in training
mode on a collated batch: the reference module under torch autograd (eager, TF32 off, baseline/_ref) against train.GraphTripleConvNetTrainer on the same GPU.  MEASUREMENT INFRASTRUCTURE.

  python tools/time_gcn_train.py [--scenes 64] [--out gpurun_out/gcn_train_timing.json]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(a):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from echoscene_b200 import arch, modules, synth, train
    dev = "cuda"
    gcfg = synth.layout_cfg().gcn()
    sd = arch.make_state_dict(arch.gcn_specs(gcfg), synth.WEIGHT_SEED_GCN)
    g = synth.batch_scene_graphs([synth.make_scene_graph(8 + i % 9, 24 + 4 * (i % 9), 50 + i) for i in range(a.scenes)])
    gen = torch.Generator().manual_seed(8)
    n, t = g.n_nodes, g.triples.shape[0]
    obj, pred = torch.randn(n, gcfg.input_dim_obj, generator=gen).to(dev), torch.randn(t, gcfg.input_dim_pred, generator=gen).to(dev)
    d_obj, d_pred = torch.randn(n, gcfg.output_dim, generator=gen).to(dev), torch.randn(t, gcfg.input_dim_pred, generator=gen).to(dev)
    edges = torch.stack([g.triples[:, 0], g.triples[:, 2]], dim=1).to(dev)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # reference arm
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "model")):
        raise SystemExit("baseline/_ref is not installed (python baseline/install_ref.py in the build container): no reference arm")
    from baseline import ref_runner
    ref = ref_runner.load_reference()
    ref_net = ref.GraphTripleConvNet(input_dim_obj=gcfg.input_dim_obj, input_dim_pred=gcfg.input_dim_pred, num_layers=gcfg.num_layers,
                                     hidden_dim=gcfg.hidden_dim, residual=True, pooling="avg", mlp_normalization="batch",
                                     output_dim=gcfg.output_dim)
    ref_net.load_state_dict(sd, strict=True)
    ref_net = ref_net.to(dev).train()
    kind = "baseline/_ref GraphTripleConvNet under .train(), torch autograd, eager fp32"

    def ref_step():
        o, p = obj.clone().requires_grad_(True), pred.clone().requires_grad_(True)
        ref_net.zero_grad(set_to_none=True)
        ro, rp = ref_net(o, p, edges)
        ((ro * d_obj).sum() + (rp * d_pred).sum()).backward()

    m = modules.GraphTripleConvNet(gcfg.input_dim_obj, gcfg.input_dim_pred, num_layers=gcfg.num_layers, hidden_dim=gcfg.hidden_dim,
                                   residual=True, pooling="avg", mlp_normalization="batch", output_dim=gcfg.output_dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    tr = train.GraphTripleConvNetTrainer(m, max_nodes=n, max_triples=t)

    def my_step():
        tr.zero_grad()
        tr.forward(obj, pred, edges)
        tr.backward(d_obj, d_pred)

    def my_fwd():
        tr.forward(obj, pred, edges)

    if a.profile_one_step:   # ncu --profile-from-start off: exactly one forward + backward of the B200 arm in the capture
        my_step(); my_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        my_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    r_ms, m_ms, f_ms = timed(ref_step, a.reps), timed(my_step, a.reps), timed(my_fwd, a.reps)
    it = tr.capture(obj, pred, edges, d_obj, d_pred)

    def graph_step():
        tr.zero_grad()
        it.replay()

    g_ms = timed(graph_step, a.reps)
    params = sum(p.numel() for p in m.parameters())
    res = {"scenes": a.scenes, "nodes": n, "triples": t, "parameters": params, "reference": kind, "reference_ms_fwd_bwd": r_ms,
           "b200_ms_fwd_bwd": m_ms, "b200_ms_fwd": f_ms, "speedup": r_ms / m_ms, "b200_ms_fwd_bwd_cuda_graph": g_ms,
           "speedup_cuda_graph": r_ms / g_ms,
           "param_plus_grad_bytes_per_step": 3 * 4 * params,
           "hbm_gbs_on_params_and_grads": 3 * 4 * params / (m_ms * 1e-3) / 1e9}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--profile-one-step", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gcn_train_timing.json"))
    main(ap.parse_args())
