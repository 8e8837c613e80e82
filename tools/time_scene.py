"""Once-per-scene stages around the chains: scene encode and VQ-VAE encode (8f-3), ms per call.
Usage: python tools/time_scene.py [n_nodes] [n_triples] [n_sdf_objects]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import _lib, arch, modules  # noqa: E402
from oracle import cases  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_sdf = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    _lib.lib().echo_launch_count_reset()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, _lib.lib().echo_launch_count() // reps


cfg = cases.scene_cfg()
enc = modules.SceneEncoder()
enc.load_state_dict(arch.make_state_dict(arch.scene_encoder_specs(cfg), cases.WEIGHT_SEED_SCENE))
enc = enc.to(dev)
g, objs, text, rel = cases.scene_inputs(cases.GraphCase("time_scene", n, t, 2))
a = [x.to(dev) for x in (objs, g.triples, text, rel)]
params = sum(v.numel() for v in enc.state_dict().values() if v.dtype == torch.float32)
ms, launches = timed(lambda: enc.encode(*a), 20)
print(f"scene encode N={n} T={t}: {ms:.3f} ms, {launches} launches, {params * 4 / ms / 1e6:.0f} GB/s of {params / 1e6:.1f} M fp32 parameters")
ms, launches = timed(lambda: enc.init_encoder(*a), 20)
print(f"  init_encoder alone: {ms:.3f} ms, {launches} launches")

vcfg = cases.vqvae_cfg()
dd = dict(double_z=False, z_channels=3, resolution=64, in_channels=1, out_ch=1, ch=64, ch_mult=[1, 2, 4], num_res_blocks=1,
          attn_resolutions=[], dropout=0.0)
vq = modules.VQVAE(dd, vcfg.n_embed, vcfg.embed_dim, with_encoder=True)
sd = dict(arch.make_state_dict(arch.vqvae_encode_specs(vcfg), cases.WEIGHT_SEED_VQVAE + 1))
sd.update(arch.make_state_dict(arch.vqvae_decode_specs(vcfg), cases.WEIGHT_SEED_VQVAE))
vq.load_state_dict(sd)
vq = vq.to(dev)
x = cases.vqvae_sdf_inputs(n_sdf, seed=5).to(dev)
ms, launches = timed(lambda: vq.encode_no_quant(x), 3, warm=1)
print(f"vqvae encode fp32 N={n_sdf}: {ms:.2f} ms ({ms / n_sdf:.2f} ms/object, {271e9 * n_sdf / ms / 1e9:.1f} TFLOP/s at 271 GFLOP/object (SURVEY 8f)), "
      f"{launches} launches")
