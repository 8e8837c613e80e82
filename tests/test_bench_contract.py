"""CPU: the committed bench lines (profiles/r1_bench_*.json, written by bench.py on a B200) carry every key of the
measurement contract, so a change to bench.py that drops one shows up here before a GPU run does."""
import json
import os

import pytest

PROF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "clocks", "gpu_launches", "e2e", "roofline"]


def _line(name):
    path = os.path.join(PROF, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.load(open(path))


def test_single_gpu_line():
    d = _line("r1_bench_n1.json")
    for k in BASE + ["cpu_baseline", "layout_branch"]:
        assert k in d, k
    assert d["metric"] == "denoiser-steps/sec" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and "sample" in c
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_round2_single_gpu_line_has_the_secondary_figures():
    """The round's final default line (profiles/r2_bench_n1.json): contract keys, the reference on the same box as cpu_baseline, and
    every secondary figure without an error entry."""
    d = _line("r2_bench_n1.json")
    for k in BASE + ["cpu_baseline", "layout_branch", "launch_mode"]:
        assert k in d, k
    assert d["steps"] >= 100 and d["cpu_baseline"]["kind"] == "reference" and "replayed CUDA graph" in d["launch_mode"]
    for k in ("layout_branch_batched_64_scenes", "sdf_to_mesh", "parity_mode_x3", "config3_n32_s250", "config4_scene_sharded",
              "gpu_eager_baseline", "scene_encode", "vqvae_decode", "full_chain_seconds_per_scene", "full_chain_seconds_per_scene_batched"):
        assert k in d and "error" not in d[k], k
    assert d["layout_branch"]["executor"]["persistent_kernel_steps"] > 0
    assert d["layout_branch_batched_64_scenes"]["value"] > 3 * d["layout_branch"]["value"]          # batching pays on the layout branch
    assert d["full_chain_seconds_per_scene_batched"]["total"] < d["full_chain_seconds_per_scene"]["total"]
    assert d["roofline"]["traffic_source"].startswith("static")


@pytest.mark.parametrize("name,n", [("r1_bench_n2.json", 2), ("r1_bench_n4.json", 4), ("r1_bench_n8.json", 8)])
def test_multi_gpu_lines(name, n):
    d = _line(name)
    for k in BASE:
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["config"]["scenes"] == n


def test_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "denoiser-steps/sec"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"]


def test_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the driver's reference arm) runs without a GPU and prints one contract line."""
    import subprocess
    import sys
    root = os.path.dirname(PROF)
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoiser-steps/sec" and d["value"] > 0
    # the unmodified reference from baseline/_ref when it is installed (build container, GPU box), else the oracle port
    from baseline import ref_runner
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_runner.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] >= 1 and d["config"]["n_nodes"] == 16 and "scaled" not in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_secondary_figures_are_never_fatal():
    """bench.py computes the scene-encode figure last and through optional_figure: without a GPU it degrades to an error entry."""
    import bench
    out = bench.optional_figure(bench.scene_encode_time, "cpu")
    assert set(out) == {"error"} and "EchoError" in out["error"]
    assert bench.optional_figure(lambda: {"ms": 1.0}) == {"ms": 1.0}
