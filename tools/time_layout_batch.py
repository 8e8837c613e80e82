"""Layout DDPM step on a collated batch of scenes (BASELINE config 4 shape): ms per batched step and scene-steps/s for S scenes of
16 nodes / 64 triples each.  MEASUREMENT INFRASTRUCTURE.   python tools/time_layout_batch.py 1 4 16 64"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(S, steps=30):
    from echoscene_b200 import arch, modules, synth
    dev = torch.device("cuda:0")
    sd = arch.make_state_dict(arch.unet1d_specs(synth.layout_cfg()), synth.WEIGHT_SEED_LAYOUT)
    m = modules.UNet1DModel(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2, attention_resolutions=[4, 2],
                            channel_mult=[1, 1, 1, 1], num_heads=8, use_spatial_transformer=True, concat_dim=1280,
                            crossattn_dim=1280, enable_t_emb=True, precision="fp32", time_num=1000)
    m.load_state_dict(sd)
    m = m.to(dev)
    g = synth.batch_scene_graphs([synth.make_scene_graph(16, 64, 2 + i) for i in range(S)])
    n = g.n_nodes
    tri = g.triples.to(dev)
    gen = torch.Generator().manual_seed(5)
    obj_embed, x = torch.randn(n, 640, generator=gen).to(dev), torch.randn(n, 8, generator=gen).to(dev)
    noise = torch.randn(steps + 5, n, 8, device=dev)
    m._ensure(n, tri.shape[0])
    m.frozen = True
    for i in range(5):
        x = m.ddpm_step(x, obj_embed, tri, 999 - i, noise[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(steps):
        x = m.ddpm_step(x, obj_embed, tri, 994 - i, noise[5 + i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"scenes": S, "nodes": n, "triples": int(tri.shape[0]), "ms_per_batched_step": ms, "scene_steps_per_s": S * 1e3 / ms,
            "finite": bool(torch.isfinite(x).all())}


if __name__ == "__main__":
    out = [run(int(a)) for a in (sys.argv[1:] or ["1", "4", "16", "64"])]
    for r in out:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/layout_batch.json", "w"), indent=1)
