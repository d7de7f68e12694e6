"""bench.py's JSON-line contract: the reference arm runs on CPU here; the GPU
arm's line is checked on the GPU box (marked gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    env = dict(os.environ)
    env.pop("RANK", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the reference's CPU path on the host cores — the unmodified reference from
    baseline/_ref when installed (kind "reference"), else the oracle port (kind "port") —, a bounded sample
    per step, same metric / unit / config as the GPU arm."""
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], timeout=600)
    assert j["impl"] == "reference"
    assert BASE_KEYS <= set(j)
    assert j["unit"] == "rays/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["value"] > 0 and j["gpu_launches"] == 0
    installed = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "models", "nvfi.py"))
    assert j["cpu_baseline"]["kind"] == ("reference" if installed else "port") and j["cpu_baseline"]["cores"] >= 1
    assert "stride" in j["cpu_baseline"]["sample"] and "valid-sample fraction" in j["cpu_baseline"]["sample"]
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                       env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    j = _run(["--steps", "1", "--warmup", "3", "--no-cpu", "--rows", "40"], timeout=900)
    assert BASE_KEYS <= set(j)
    assert j["n_gpus"] == 1 and j["value"] > 0 and j["gpu_launches"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "k_march" in j["kernels"] and "k_sample_advect_h" in j["kernels"]
    assert j["invalid_for_bench"]          # a 40-row band is a profiling aid, not the bench workload
