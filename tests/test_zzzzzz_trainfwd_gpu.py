"""GPU: the training forward, loss against loss.  The UNMODIFIED reference (baseline/_ref) under model.train() runs
one SGDiff.forward_mani on a collated batch of three scenes (scripts/train_3dfront.py:237-241), eager fp32 on this GPU; the B200
components, built from the same YAML files and loaded with the same weights, run train().forward_mani on the same batch under the
same numpy / torch seeds.  Every stage that carries a BatchNorm1d (the two scene encoders, rel_s_mlp, both denoisers' GCNs) is
compared on its own first (the denoiser OUTPUTS: with random-initialised weights the losses themselves are dominated by the noise
target), then every entry of the loss dictionary and both loss totals -- for a manipulation batch, for an addition (one node
missing from the encoder-side scene, replace_latent = False) and for the layout-only model: all within north_star's 1e-3."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "trainfwd_check.py")


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "model")), reason="baseline/_ref (the reference copy) is not installed")
def test_forward_mani_losses_match_the_reference(tmp_path):
    out = str(tmp_path / "trainfwd.json")
    r = subprocess.run([sys.executable, TOOL, "--out", out], cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True,
                       text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["selected_identical"] and res["worst_rel"] < 1e-3
    assert all(v < 1e-3 for v in res["stages"].values()), res["stages"]
    assert {"loss_simple", "loss_vlb", "loss.bbox", "loss.angle", "Shape_loss", "Layout_loss"} <= set(res["losses"])
    # an addition with replace_latent = False (zero row inserted, touched rows only), and the layout-only model
    assert {"loss_simple", "Layout_loss"} <= set(res["losses_addition"]) and "Layout_loss" in res["losses_layout_only"]
