"""Pin the scene-encoder oracle against the reference and write tests/golden/scene_encode.pt --
TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference):  python oracle/gen_golden_scene.py

`Sg2ScDiffModel.__init__` builds the two diffusion branches, loads a VQ-VAE checkpoint and opens log files
(model/EchoScene.py:91-113), none of which the encoders need.  The reference's own METHODS are therefore run unbound on a
holder module that owns exactly the sub-modules they read, each built with the reference's own constructors
(nn.Embedding, model.graph.GraphTripleConvNet, model.graph.make_mlp) the way EchoScene.py:46-100 builds them:
`Sg2ScDiffModel.init_encoder(holder, ...)`, `Sg2ScDiffModel.manipulate(holder, ...)`, `holder.rel_s_mlp(...)`, in the
order of `Sg2ScDiffModel.sample` (EchoScene.py:388-410).
"""
from __future__ import annotations

import importlib
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echoscene_b200 import arch                      # noqa: E402
from oracle import cases, echoscene_oracle as orc    # noqa: E402
from oracle import ref_import                        # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.manual_seed(0)
    ref = ref_import.load()
    graph = importlib.import_module("model.graph")
    es = importlib.import_module("model.EchoScene")
    cfg = cases.scene_cfg()
    gd, add = cfg.gconv_dim, cfg.add_dim

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.clip = True
            self.embedding_dim = gd
            self.obj_embeddings_ec = nn.Embedding(cfg.num_objs + 1, gd * 2)
            self.pred_embeddings_ec = nn.Embedding(cfg.num_preds, gd * 2)
            kw = dict(hidden_dim=gd * 4, pooling="avg", mlp_normalization="batch", residual=cfg.residual)
            self.gconv_net_ec = ref.GraphTripleConvNet(input_dim_obj=gd * 2 + add, input_dim_pred=gd * 2 + add, num_layers=cfg.num_layers,
                                                       output_dim=gd * 2 + add, **kw)
            self.gconv_net_manipulation = ref.GraphTripleConvNet(input_dim_obj=(gd * 2 + add) + gd + gd * 2 + add,
                                                                 input_dim_pred=gd * 2 + add, num_layers=min(cfg.num_layers, 5),
                                                                 output_dim=gd * 2 + add, **kw)
            self.rel_s_mlp = graph.make_mlp([gd * 2 + add, 960, 1280], batch_norm="batch", norelu=True)

    h = Holder().eval()
    sd = arch.make_state_dict(arch.scene_encoder_specs(cfg), cases.WEIGHT_SEED_SCENE)
    h.load_state_dict(sd, strict=True)
    g, objs, text, rel = cases.scene_inputs()
    M = es.Sg2ScDiffModel
    with torch.no_grad():
        obj_embed, pred_embed, latent_obj, _ = M.init_encoder(h, objs, g.triples, text, rel)
        change = torch.zeros(latent_obj.shape[0], gd)                     # EchoScene.py:393-397 (np.zeros per node, on the GPU there)
        latent_, _, obj_embed_, _ = M.manipulate(h, torch.cat([latent_obj, change], dim=1), objs, g.triples, text, rel)
        uc = torch.unsqueeze(h.rel_s_mlp(obj_embed_), dim=1)
        c = torch.unsqueeze(h.rel_s_mlp(latent_), dim=1)
        got = orc.scene_encode(sd, cfg, objs, g.triples, text, rel)
    want = {"obj_embed": obj_embed_, "latent": latent_, "uc_s": uc, "c_s": c}
    rec = {}
    for k, v in want.items():
        d = (got[k].double() - v.double())
        rec[k] = {"max_abs": float(d.abs().max()), "rel_l2": float(d.norm() / v.double().norm())}
    rec["params"] = arch.count_params(arch.scene_encoder_specs(cfg))
    torch.save(want, os.path.join(GOLD, "scene_encode.pt"))
    path = os.path.join(GOLD, "PINNING.json")
    pin = json.load(open(path))
    pin["cases"]["scene_encode"] = {"rel_l2": max(r["rel_l2"] for r in rec.values() if isinstance(r, dict)), "detail": rec}
    with open(path, "w") as f:
        json.dump(pin, f, indent=1)
    print(json.dumps(rec, indent=1))
    assert pin["cases"]["scene_encode"]["rel_l2"] < 1e-5


if __name__ == "__main__":
    main()
