"""The reference's benchmark protocol (`B` key: app.h:24-30, app.cpp:88-161, 238-274) as tools/benchmark_recorder.py
runs it headless: N iterations x F frames, the scene reloaded per iteration, min / avg / max of the nine
SPHStatistics.time buckets and of the whole Update over all frames."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "benchmark_recorder.py")
BUCKETS = {"emitters", "integration", "viscosityForces", "predict", "updateGrid", "neighborSearch", "densityAndPressure", "deltaPositions", "collisions"}


def record(impl, scenario=2, iterations=2, frames=6):
    out = subprocess.run([sys.executable, TOOL, "--impl", impl, "--scenario", str(scenario), "--iterations", str(iterations), "--frames", str(frames)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def check(doc, impl, particles):
    assert doc["impl"] == impl and doc["iterations"] == 2 and doc["frames"] == 6
    assert doc["particles_at_end"] == particles  # the scene was reloaded per iteration, not accumulated
    assert set(doc["phases_ms"]) == BUCKETS  # sph.h:131-141
    for b in doc["phases_ms"].values():
        assert 0.0 <= b["min"] <= b["avg"] <= b["max"]
    u = doc["update_ms"]
    assert 0.0 < u["min"] <= u["avg"] <= u["max"]
    assert doc["particle_steps_per_s_avg"] > 0
    # on the CPU the pair passes are where the time goes (27 of 28 ms in the survey's probe); on the GPU a 1400-particle
    # scene is launch-latency bound in every phase, so no share is asserted there (a timing assertion would be flaky)
    if impl != "b200":
        heavy = sum(doc["phases_ms"][k]["avg"] for k in ("viscosityForces", "neighborSearch", "densityAndPressure", "deltaPositions"))
        assert heavy > 0.5 * sum(b["avg"] for b in doc["phases_ms"].values())


def test_recorder_on_the_oracle_in_reference_semantics():
    check(record("oracle"), "oracle", 1400)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsphref.so")), reason="libsphref.so not built")
def test_recorder_on_the_reference_binary():
    check(record("ref"), "ref", 1400)


@pytest.mark.gpu
def test_recorder_on_the_gpu():
    doc = record("b200")
    check(doc, "b200", 1400)
    assert doc["phases_ms"]["neighborSearch"]["max"] == 0.0  # no neighbour lists are materialised
