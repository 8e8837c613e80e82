"""GPU: the drop-in claim, run for real (VERDICT r1 task 6).  The UNMODIFIED reference (baseline/_ref, the verbatim copy made by
baseline/install_ref.py) builds its own SGDiff('echoscene', config/full_mp.yaml) and samples a scene through its own
Sg2ScDiffModel.sample -- once as it is (eager PyTorch on this GPU, TF32 off), once after integrate.patch_reference() rebound
its hot-path classes to libechoscene_b200.so -- from the same synthetic checkpoint, scene graph and RNG seed.  Boxes (sizes,
translations, angles) and shape latents must agree within north_star's 1e-3.  Each arm is its own process (the patch is global).
The layout chain runs 20 DDPM steps and the shape chain the reference's debug setting of 7 DDIM steps (echo2shape.py:116-120) so
that the eager arm stays short; the code paths are the full ones (model/SGDiff.py:87-95, model/EchoScene.py:388-420)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "refbind_check.py")


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "model")), reason="baseline/_ref (the reference copy) is not installed")
def test_reference_sample_unpatched_vs_patched(tmp_path):
    ref, b200 = str(tmp_path / "ref.pt"), str(tmp_path / "b200.pt")
    env = dict(os.environ, PYTHONPATH=ROOT)
    for cmd in ([sys.executable, TOOL, "--arm", "reference", "--out", ref],
                [sys.executable, TOOL, "--arm", "patched", "--out", b200, "--like", ref]):
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        print(r.stdout.strip().splitlines()[-1])
    r = subprocess.run([sys.executable, TOOL, "--compare", ref, b200], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, "patched and unpatched reference disagree beyond 1e-3:\n" + r.stdout + r.stderr[-2000:]
