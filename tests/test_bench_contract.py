"""bench.py's JSON contract.  CPU: the --impl reference arm (the oracle on host cores).  GPU: the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_json_cpu():
    d = run_bench("--impl", "reference", "--m", "8", "--n", "24", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "bases/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 1e5
    assert d["e2e"] == {"value": d["value"], "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    rc = d["reference_code"]            # side figure: the reference's own per-basis code (oracle/_ref), when built
    assert rc is None or rc.get("value", 0) > 1e3 or "unavailable" in rc


@pytest.mark.gpu
def test_product_arm_json_gpu(gpu_lib):
    d = run_bench("--m", "10", "--n", "30", "--steps", "3", "--warmup", "3", "--cpu-ranks", "3000000")
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["gpu_launches"] == d["steps"]           # ONE kernel per enumeration (fused finalize, no memset)
    assert d["ms_per_step_best"] <= d["ms_per_step_median"] and d["roofline"]["launch_ms_best"] <= d["roofline"]["launch_ms_median"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 256 and d["e2e"]["value"] > 1e8
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and r["peak"] > 10 and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["per_basis_lu_kernel"]["frac"] < 1.0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 1e5
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["result"]["best_rank"] == 1579911 and d["result"]["n_feasible"] == 66804      # the oracle's answer for this LP
