"""bench.py contract: one JSON line with the keys the driver reads, for both arms."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(*args):
    pr = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert pr.returncode == 0, pr.stderr[-2000:]
    lines = [ln for ln in pr.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, pr.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--size", "XS", "--steps", "2", "--warmup", "1", "--no-direct-sample")
    assert "direct_solver_sample" not in d
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["unit"] == "solves/s" and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly(monkeypatch):
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    pr = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "XS"],
                        capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert pr.returncode == 0 and pr.stdout.strip() == ""


def test_direct_solver_sample_of_the_reference_arm():
    # the reference's solver class (sparse direct) on a slab that factors in a blink here; the arm uses 48x36x30
    sys.path.insert(0, str(ROOT))
    import bench
    d = bench.direct_solver_sample((12, 9, 8))
    assert d["nodes"] > 500 and d["rel_residual"] < 1e-9 and d["fill_nnz"] > d["matrix_nnz"] and d["solves_per_s"] > 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run("--size", "S", "--steps", "2", "--warmup", "3", "--cpu-iters", "20")
    assert BASE_KEYS <= set(d) and "impl" not in d and d["value"] > 0 and d["n_gpus"] == 1
    assert d["gpu_launches"] > 0 and d["scaling"] == "weak" and d["higher_is_better"] is True
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    ref = _run("--impl", "reference", "--size", "S", "--steps", "1", "--warmup", "1", "--no-direct-sample")
    assert ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
