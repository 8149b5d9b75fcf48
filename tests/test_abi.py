"""CPU: the C-ABI library builds, loads, and exports every symbol include/sphb200.h declares; the
ctypes mirror agrees with the header's struct layout; and without a GPU the product refuses to run
(no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sphb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from nbodysimulation_experiment_b200 import _lib, build

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    assert lib.sph_abi_version() == 1


def test_ctypes_structs_match_header():
    from nbodysimulation_experiment_b200 import _lib

    src = '#include "sphb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(SphConfig), sizeof(SphParams), sizeof(SphStats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.SphConfig), ctypes.sizeof(_lib.SphParams), ctypes.sizeof(_lib.SphStats)]


def test_sass_is_sm100a():
    from nbodysimulation_experiment_b200 import build

    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nbodysimulation_experiment_b200 import ParticleSimulation, SphError

    with pytest.raises(SphError) as e:
        ParticleSimulation()
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "nbodysimulation_experiment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "sph_oracle" not in text and "libsphref" not in text and "oracle_lib" not in text, f


def _build_demo(tmp_path):
    from nbodysimulation_experiment_b200 import build

    lib = build.build()
    exe = str(tmp_path / "demo_host")
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "demo_host.cpp"), "-L", os.path.dirname(lib),
                    "-lsphb200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe], check=True)
    return exe


def test_cpp_host_class_builds_and_refuses_without_gpu(tmp_path):
    """include/sphb200_sim.hpp (the BaseSimulation-shaped C++ host class) compiles against the C ABI
    alone; without a device the program fails loudly instead of falling back."""
    import torch

    exe = _build_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 3 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_cpp_host_class_runs_the_app_loop(tmp_path):
    exe = _build_demo(tmp_path)
    out = subprocess.run([exe, "0", "16"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "5300 particles" in out.stdout
