"""GPU: the SGDiff facade (echoscene_b200/sgdiff.py) built from YAML with the reference's key structure, sampling through
the CUDA library end to end: echolayout (BASELINE config 1 shape: N = 8, 10 DDPM steps) and echoscene (N = 3 objects, full-size
denoisers, the 7-step debug DDIM schedule, VQ-VAE decode).  Parity of every stage is asserted in the tests of that stage;
here the wiring is: shapes, finiteness, determinism under a fixed seed, agreement with the components called by hand."""
import copy

import numpy as np
import pytest
import torch
import yaml

from echoscene_b200 import sgdiff
from oracle import cases
from test_sgdiff_host import MAIN, VOCAB, VQ

pytestmark = pytest.mark.gpu
DEV = "cuda"

DF_FULL = {"model": {"params": {"linear_start": 0.00085, "linear_end": 0.012, "conditioning_key": "crossattn", "timesteps": 1000}},
           "unet": {"params": {"image_size": 16, "in_channels": 3, "out_channels": 3, "model_channels": 224, "num_res_blocks": 2,
                               "attention_resolutions": [4, 2], "channel_mult": [1, 2, 3], "num_heads": 8, "dims": 3,
                               "use_spatial_transformer": True, "transformer_depth": 1, "context_dim": 1280, "use_checkpoint": True,
                               "legacy": False, "messsage_passing": True, "enable_t_emb": True}}}


@pytest.fixture()
def cfg_dir(tmp_path):
    d = tmp_path / "config"
    d.mkdir()
    for name, body in (("df.yaml", DF_FULL), ("vq.yaml", VQ)):
        (d / name).write_text(yaml.safe_dump(body))
    return d


def _cuda(*ts):
    return [t.to(DEV) for t in ts]


def test_echolayout_facade_samples_boxes(cfg_dir):
    opt = copy.deepcopy(MAIN)
    opt["layout_branch"]["diffusion_kwargs"]["time_num"] = 10
    m = sgdiff.SGDiff("echolayout", opt, VOCAB, residual=True, config_dir=str(cfg_dir)).cuda().eval()
    g, objs, text, rel = cases.scene_inputs()
    torch.manual_seed(3)
    out = m.sample_box_and_shape(*_cuda(objs, g.triples, text, rel))
    assert set(out) == {"sizes", "translations", "angles"}
    assert out["sizes"].shape == (8, 3) and out["translations"].shape == (8, 3) and out["angles"].shape == (8, 2)
    assert all(torch.isfinite(v).all() for v in out.values())
    torch.manual_seed(3)
    again = m.sample_box_and_shape(*_cuda(objs, g.triples, text, rel))
    assert all(torch.equal(out[k], again[k]) for k in out)                     # same torch seed -> same chain
    # by hand: encoders, then the DiffusionPoint chain on their obj_embed
    torch.manual_seed(3)
    e = m.encoder.encode(*_cuda(objs, g.triples, text, rel), shape_cond=False)
    boxes = m.layout.gen_samples_sg((8, 8), torch.device(DEV), e["obj_embed"], g.triples.to(DEV), condition=e["latent"],
                                    clip_denoised=False)
    assert torch.equal(boxes[:, 0:3], out["sizes"]) and torch.equal(boxes[:, 6:8], out["angles"])
    np.random.seed(1)
    keep, out2 = m.sample_boxes_and_shape_with_changes(*_cuda(objs, g.triples, text, rel, objs, g.triples, text, rel), [2, 4])
    assert keep.flatten().tolist() == [1, 1, 0, 1, 0, 1, 1, 1] and out2["sizes"].shape == (8, 3)


def test_echoscene_facade_samples_boxes_and_shapes(cfg_dir):
    opt = copy.deepcopy(MAIN)
    opt["layout_branch"]["diffusion_kwargs"]["time_num"] = 5
    opt["misc"]["debug"] = 1                                                     # ddim_steps = 7 (echo2shape.py:116-120)
    m = sgdiff.SGDiff("echoscene", opt, VOCAB, residual=True, config_dir=str(cfg_dir)).cuda().eval()
    g, objs, text, rel = cases.scene_inputs(cases.GraphCase("facade", 3, 4, 51))
    torch.manual_seed(4)
    out = m.sample_box_and_shape(*_cuda(objs, g.triples, text, rel), gen_shape=True)
    assert set(out) == {"shapes", "sizes", "translations", "angles"}
    assert out["shapes"].shape == (3, 1, 64, 64, 64) and torch.isfinite(out["shapes"]).all()
    assert out["sizes"].shape == (3, 3) and torch.isfinite(out["sizes"]).all()
    layout_only = m.sample_box_and_shape(*_cuda(objs, g.triples, text, rel), gen_shape=False)
    assert layout_only["shapes"] is None
