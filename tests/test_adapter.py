"""The drop-in boundary compiles against the REFERENCE's own headers ("adapter was never compiled"):
examples/demo_b200.h - the BaseSimulation subclass of INTEGRATION.md section 1 - is built against
/root/reference/NBodySimulation/base.h, sph.h, vecmath.h (under the oracle's g++ shim) and render.h and linked with
libsphb200.so by oracle/Makefile's `adapter` target.  The binary replays LoadScenario + the app loop (app.cpp:228-236,
477-534) through a BaseSimulation pointer; on a box without a GPU it stops at sph_create's refusal."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adapter_check")
HAVE_REFERENCE = os.path.isdir("/root/reference/NBodySimulation")


def test_integration_md_shows_the_adapter_that_is_compiled():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(ROOT, "examples", "demo_b200.h")).read()
    body = src[src.index("#ifndef DEMO_B200_H"):]
    assert body.strip() in md, "INTEGRATION.md section 1 and examples/demo_b200.h have drifted apart"


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference tree (authoring container)")
def test_adapter_compiles_against_the_reference_headers_and_links():
    from nbodysimulation_experiment_b200 import build

    build.build()
    subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "_ref/adapter_check"], check=True)
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adapter linked" in out.stdout or "adapter ran" in out.stdout


@pytest.mark.gpu
def test_adapter_runs_the_app_loop_through_the_base_class_pointer():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/adapter_check was not built (needs the reference tree at build time)")
    out = subprocess.run([BIN, "need-gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adapter ran: 1400 particles" in out.stdout
