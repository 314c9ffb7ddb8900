"""bench.py prints exactly one JSON line on stdout with the contract's keys (checked here through the reference
arm on the small workload, which needs no GPU; the GPU arm is checked by the gpu-marked test)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--workload", "cube", "--steps", "1", "--warmup", "0", "--cpu-spp", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cube", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--workload", "cube", "--steps", "2", "--warmup", "3")
    assert BASE_KEYS <= set(d) and "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["parity"]["identical"] >= 0.999 and d["parity"]["rmse"] <= 1e-3          # the bench line carries its own parity evidence
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
