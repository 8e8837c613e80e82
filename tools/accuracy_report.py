"""Measured deviation of both precision modes from the committed reference outputs (tests/golden) — run on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from echoscene_b200 import _lib, arch  # noqa: E402
from oracle import cases  # noqa: E402
from test_model_gpu import layout_model, shape_model  # noqa: E402
from util import gold, rel_err  # noqa: E402

out = {}
cfg = cases.shape_cfg()
sd = arch.make_state_dict(arch.unet3d_specs(cfg), cases.WEIGHT_SEED_SHAPE)
g, uc, x, t = cases.shape_step_inputs(cases.SHAPE_CASE, cfg)
G = gold("shape.pt")
for prec in ("fp32", "bf16"):
    if prec == "bf16" and not _lib.lib().echo_has_tcgen05():
        continue
    m = shape_model(sd, precision=prec)
    o = m(x.cuda(), uc.cuda(), g.triples.cuda(), t.cuda())
    out[f"shape_step_{prec}"] = dict(zip(("max_rel", "rel_l2"), rel_err(o, G["step"])))
    gc, ucc, xT, _ = cases.shape_step_inputs(cases.SHAPE_CHAIN_CASE, cfg, same_noise=True)
    xx = xT.cuda()
    for i in range(cases.SHAPE_CHAIN_STEPS):
        xx = m.ddim_step(xx, ucc.cuda(), gc.triples.cuda(), 100 - i - 1)
    out[f"shape_chain3_{prec}"] = dict(zip(("max_rel", "rel_l2"), rel_err(xx, G["chain"])))
    del m
lcfg = cases.layout_cfg()
sdl = arch.make_state_dict(arch.unet1d_specs(lcfg), cases.WEIGHT_SEED_LAYOUT)
g, obj_embed, x, t = cases.layout_step_inputs(cases.LAYOUT_CASE, lcfg)
for prec in ("fp32", "bf16"):
    m = layout_model(sdl, precision=prec)
    o = m(x.cuda(), obj_embed.cuda(), g.triples.cuda(), t.cuda())
    out[f"layout_step_{prec}"] = dict(zip(("max_rel", "rel_l2"), rel_err(o, gold("layout_n8.pt")["step"])))
print(json.dumps(out, indent=1))
