"""Golden vectors for the host glue of Sg2ScDiffModel.sample / sample_with_changes / sample_with_additions
-- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference):  python oracle/gen_golden_scene_glue.py

The reference's own methods (model/EchoScene.py:388-532) are run unbound on a holder module that owns the encoder
sub-modules (built by the reference's constructors, as in oracle/gen_golden_scene.py) and two recording stubs in place of the
diffusion branches: `LayoutDiff.set_input` records the conditioning the layout chain would receive, `ShapeDiff.rel2shape`
records the shape conditioning.  What is pinned is therefore everything the glue decides: change flags (np.random stream),
inserted zero rows, which latent rows are replaced, rel_s_mlp inputs, `keep`.  `.cuda()` is made the identity for the
duration of the run (the methods hard-code it, :396, :438, :484) -- this container has no GPU.
Writes tests/golden/scene_glue.pt (the recorded tensors); inputs are regenerated from seeds by oracle/cases.py.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from echoscene_b200 import arch              # noqa: E402
from oracle import cases, ref_import         # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _Rec:
    def __init__(self):
        self.layout_in = None
        self.shape_in = None


def main():
    torch.manual_seed(0)
    ref = ref_import.load()
    graph = importlib.import_module("model.graph")
    es = importlib.import_module("model.EchoScene")
    cfg = cases.scene_cfg()
    gd, add = cfg.gconv_dim, cfg.add_dim
    M = es.Sg2ScDiffModel
    MB = importlib.import_module("model.EchoLayout").Sg2BoxDiffModel
    rec = _Rec()

    class LayoutStub:
        def set_input(self, d):
            rec.layout_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}

        def generate_layout_sg(self, box_dim):
            n = rec.layout_in["uc_b"].shape[0]
            z = torch.zeros(n, box_dim)
            return {"sizes": z[:, 0:3], "translations": z[:, 3:6], "angles": z[:, 6:8]}

    class ShapeStub:
        def rel2shape(self, d):
            rec.shape_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}
            return torch.zeros(d["c_s"].shape[0], 1, 2, 2, 2)

    class Holder(nn.Module):
        def __init__(self, replace_latent):
            super().__init__()
            self.clip = True
            self.embedding_dim = gd
            self.out_dim_ini_encoder = gd * 2 + add
            self.replace_all_latent = replace_latent
            self.obj_embeddings_ec = nn.Embedding(cfg.num_objs + 1, gd * 2)
            self.pred_embeddings_ec = nn.Embedding(cfg.num_preds, gd * 2)
            kw = dict(hidden_dim=gd * 4, pooling="avg", mlp_normalization="batch", residual=cfg.residual)
            self.gconv_net_ec = ref.GraphTripleConvNet(input_dim_obj=gd * 2 + add, input_dim_pred=gd * 2 + add, num_layers=cfg.num_layers,
                                                       output_dim=gd * 2 + add, **kw)
            self.gconv_net_manipulation = ref.GraphTripleConvNet(input_dim_obj=(gd * 2 + add) + gd + gd * 2 + add,
                                                                 input_dim_pred=gd * 2 + add, num_layers=min(cfg.num_layers, 5),
                                                                 output_dim=gd * 2 + add, **kw)
            self.rel_s_mlp = graph.make_mlp([gd * 2 + add, 960, 1280], batch_norm="batch", norelu=True)
            self.diff_cfg = types.SimpleNamespace(layout_branch=types.SimpleNamespace(denoiser_kwargs=types.SimpleNamespace(in_channels=8)))
            self.LayoutDiff = LayoutStub()
            self.ShapeDiff = ShapeStub()

        init_encoder = M.init_encoder
        manipulate = M.manipulate
        prepare_boxes = M.prepare_boxes

    class BoxHolder(Holder):                    # the layout-only model: its own methods and predicate table, no rel_s_mlp
        def __init__(self, replace_latent):
            super().__init__(replace_latent)
            del self.rel_s_mlp
            self.pred_embeddings_man_dc = nn.Embedding(cfg.num_preds, gd * 2)

        init_encoder = MB.init_encoder
        manipulate = MB.manipulate              # looks predicates up in pred_embeddings_man_dc (EchoLayout.py:154)
        prepare_input = MB.prepare_input

    sd = arch.make_state_dict(arch.scene_encoder_specs(cfg), cases.WEIGHT_SEED_SCENE)
    sd_box = cases.scene_box_state_dict()
    out = {}
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, fn, replace in cases.SCENE_GLUE_CASES:
            if fn.startswith("sampleBoxes"):
                h = BoxHolder(replace).eval()
                h.load_state_dict(sd_box, strict=True)
            else:
                h = Holder(replace).eval()
                h.load_state_dict(sd, strict=True)
            args, marked = cases.scene_glue_inputs(name)
            np.random.seed(cases.SCENE_GLUE_NP_SEED)
            rec.layout_in = rec.shape_in = None
            if fn == "sample":
                res = M.sample(h, *args, gen_shape=True)
                keep = None
            elif fn == "sampleBoxes":
                res = MB.sampleBoxes(h, *args)
                keep = None
            elif fn.startswith("sampleBoxes"):
                res = getattr(MB, fn)(h, *args, marked)
                keep = res[0]
            else:
                res = getattr(M, fn)(h, *args, marked, gen_shape=True)
                keep = res[0]
            out[name] = {"uc_b": rec.layout_in["uc_b"], "c_b": rec.layout_in["c_b"], "preds": rec.layout_in["preds"], "keep": keep}
            if rec.shape_in is not None:
                out[name].update({"uc_s": rec.shape_in["uc_s"], "c_s": rec.shape_in["c_s"]})
            print(name, fn, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out[name].items()})
    finally:
        torch.Tensor.cuda = real_cuda
    torch.save(out, os.path.join(GOLD, "scene_glue.pt"))
    pin_glue(out)


def pin_glue(gold):
    """Runs echoscene_b200/scene.py on the oracle encoders for every case and records its deviation from the reference's
    recorded tensors in tests/golden/PINNING.json (expected: exactly 0)."""
    import json
    from echoscene_b200 import scene
    from oracle.scene_encoder import OracleSceneEncoder

    class Lay:
        def gen_samples_sg(self, shape, device, obj_embed, triples=None, condition=None, **kw):
            self.seen = {"uc_b": obj_embed, "c_b": condition}
            return torch.zeros(shape)

    class DDIM:
        seen = None

        def __init__(self, model):
            pass

        def sample(self, S, batch_size, shape, conditioning=None, x_T=None, unconditional_conditioning=None, **kw):
            DDIM.seen = {"c_s": conditioning, "uc_s": unconditional_conditioning}
            return x_T, {}

    rec = {}
    for name, fn, replace in cases.SCENE_GLUE_CASES:
        box = fn.startswith("sampleBoxes")
        lay = Lay()
        cls = scene.Sg2BoxDiffModel if box else scene.Sg2ScDiffModel
        kw = {} if box else dict(shape=object(), ddim_sampler_cls=DDIM)
        m = cls(OracleSceneEncoder(box=box), lay, replace_latent=replace, **kw)
        args, marked = cases.scene_glue_inputs(name)
        np.random.seed(cases.SCENE_GLUE_NP_SEED)
        if fn == "sample":
            m.sample(*args, gen_shape=True)
        elif fn == "sampleBoxes":
            m.sampleBoxes(*args)
        elif box:
            getattr(m, fn)(*args, marked)
        else:
            getattr(m, fn)(*args, marked, gen_shape=True)
        seen = dict(lay.seen)
        if not box:
            seen.update(DDIM.seen)
        rec[name] = max(float((seen[k].double() - gold[name][k].double()).abs().max()) for k in seen)
    path = os.path.join(GOLD, "PINNING.json")
    pin = json.load(open(path))
    pin["cases"]["scene_glue"] = {"max_abs": max(rec.values()), "detail": rec,
                                  "what": "tensors handed to LayoutDiff.set_input / ShapeDiff.rel2shape by the reference's own "
                                          "sample* / sampleBoxes* methods vs echoscene_b200/scene.py on the oracle encoders"}
    with open(path, "w") as f:
        json.dump(pin, f, indent=1)
    print("scene_glue pinning:", rec)
    assert pin["cases"]["scene_glue"]["max_abs"] == 0.0


if __name__ == "__main__":
    main()
