"""bench.py prints exactly ONE JSON line on stdout with the keys the driver reads (both arms).

CPU: the reference arm (`--impl reference`, the oracle port on host cores) on a tiny step count.
GPU: our arm on a short run; also checks that the line is internally consistent (value = particles x steps / time,
roofline fraction = achieved / peak, e2e no faster than the device-timed value).
(Named zz so that it runs after the parity tests under `pytest -x`.)
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
          "cpu_baseline")


def run_bench(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must be one JSON line, got {len(lines)}: {out.stdout[:500]}"
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    for k in COMMON:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line():
    steps, warmup = 24, 4
    d = run_bench("--steps", str(steps), "--warmup", str(warmup), "--no-cpu")
    for k in COMMON + ("gpu_launches", "clocks", "roofline", "phases_ms"):
        assert k in d, k
    assert "impl" not in d or d["impl"] != "reference"
    assert d["metric"] == "particle_steps_per_s" and d["n_gpus"] == 1 and d["steps"] == steps and d["warmup"] >= 3
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    n = d["config"]["particles"]
    assert n == 1024 * 1024 and "dambreak_1m" in d["config"]["workload"]
    assert abs(d["value"] - n * steps / (d["ms_per_step"] * steps * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] <= d["value"] * 1.02  # a frame includes the step
    assert e["d2h_bytes_per_step"] == n * 24 and e["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["gpu_launches"] % steps == 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert "color_sweep_flow_kernel" in r["kernel"]  # the one-launch sweep is what a 1M scene runs
    if r.get("issue"):
        assert 0.3 < r["issue"]["frac"] <= 1.0
    assert abs(sum(v for k, v in d["phases_ms"].items()) - d["ms_per_step"]) < 0.35 * d["ms_per_step"]
