"""Pin the VQ-VAE decode oracle against the reference and write tests/golden/vqvae_decode.pt --
TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference):  python oracle/gen_golden_vqvae.py

Builds the reference ``VQVAE`` (model/networks/vqvae_networks/network.py:56-103) with config/vqvae_snet.yaml, loads the
seeded decode-path tensors of ``arch.vqvae_decode_specs`` (strict=False: the only keys allowed to be missing are the
encoder's and quant_conv's, which ``decode_no_quant`` never touches; no key may be unexpected), runs
``decode_no_quant`` on CPU fp32, compares the oracle restatement and merges the record into tests/golden/PINNING.json.
The fixture keeps every second voxel of the reference output plus whole-tensor statistics (1 MB per object otherwise).
"""
from __future__ import annotations

import json
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echoscene_b200 import arch                      # noqa: E402
from oracle import cases, echoscene_oracle as orc    # noqa: E402
from oracle import ref_import                        # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    ref = ref_import.load()
    with open(os.path.join(ref_import.REF_ROOT, "config/vqvae_snet.yaml")) as f:
        vq = yaml.safe_load(f)["model"]["params"]
    cfg = cases.vqvae_cfg()
    assert (vq["embed_dim"], vq["n_embed"], tuple(vq["ddconfig"]["ch_mult"]), vq["ddconfig"]["ch"]) == \
        (cfg.embed_dim, cfg.n_embed, tuple(cfg.ch_mult), cfg.ch)
    net = ref.VQVAE(vq["ddconfig"], vq["n_embed"], vq["embed_dim"]).eval()
    sd = arch.make_state_dict(arch.vqvae_decode_specs(cfg), cases.WEIGHT_SEED_VQVAE)
    res = net.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("encoder.") or k.startswith("quant_conv.") for k in res.missing_keys), res.missing_keys
    for k, v in sd.items():
        assert tuple(net.state_dict()[k].shape) == tuple(v.shape), k
    z = cases.vqvae_inputs()
    with torch.no_grad():
        want = net.decode_no_quant(z)
        quant_ref, _, info = net.quantize(z, is_voxel=True)
        got = orc.vqvae_decode_no_quant(sd, cfg, z)
        quant_o, idx_o = orc.vq_quantize(sd, z)
    d = (got.double() - want.double())
    rec = {"max_abs": float(d.abs().max()), "rel_l2": float(d.norm() / want.double().norm()), "ref_abs_max": float(want.abs().max()),
           "indices_equal": bool(torch.equal(idx_o, info[2])), "quant_equal": bool(torch.equal(quant_o, quant_ref)),
           "objects": int(z.shape[0]), "params": arch.count_params(arch.vqvae_decode_specs(cfg))}
    torch.save({"dec_sub": want[:, :, ::2, ::2, ::2].contiguous(), "dec_sum": want.double().sum(), "dec_abs_sum": want.double().abs().sum(),
                "indices": info[2].to(torch.int32), "quant": quant_ref}, os.path.join(GOLD, "vqvae_decode.pt"))
    # ---- encode_no_quant (oracle only): encoder + quant_conv, one object ----
    esd = arch.make_state_dict(arch.vqvae_encode_specs(cfg), cases.WEIGHT_SEED_VQVAE + 1)
    net2 = ref.VQVAE(vq["ddconfig"], vq["n_embed"], vq["embed_dim"]).eval()
    res2 = net2.load_state_dict(esd, strict=False)
    assert not res2.unexpected_keys, res2.unexpected_keys
    assert all(k.startswith("decoder.") or k.startswith("post_quant_conv.") or k.startswith("quantize.") for k in res2.missing_keys)
    xs = cases.vqvae_sdf_inputs()
    with torch.no_grad():
        ewant = net2.encode_no_quant(xs)
        egot = orc.vqvae_encode_no_quant(esd, cfg, xs)
    de = (egot.double() - ewant.double())
    erec = {"max_abs": float(de.abs().max()), "rel_l2": float(de.norm() / ewant.double().norm()), "ref_abs_max": float(ewant.abs().max()),
            "params": arch.count_params(arch.vqvae_encode_specs(cfg))}
    torch.save({"z": ewant}, os.path.join(GOLD, "vqvae_encode.pt"))
    path = os.path.join(GOLD, "PINNING.json")
    pin = json.load(open(path))
    pin["cases"]["vqvae_encode_no_quant"] = erec
    print(json.dumps(erec, indent=1))
    pin["cases"]["vqvae_decode_no_quant"] = rec
    with open(path, "w") as f:
        json.dump(pin, f, indent=1)
    print(json.dumps(rec, indent=1))
    assert rec["rel_l2"] < 1e-5 and rec["indices_equal"], rec


if __name__ == "__main__":
    main()
