"""GPU trajectory parity against the REFERENCE through the C ABI: scenes 0-3 at K in {1, 8, 64, 256}
steps, kinetic energy / centre of mass / mean density / count / extent inside the reference's own MT-vs-ST envelope
(tests/envelope_lib.py, tests/golden/envelope.json = unmodified demo4.cpp via tools/make_envelope.py); and bit parity
with the oracle's coloured mode at the sizes of BASELINE.json configs[2] and configs[4]."""
import os

import numpy as np
import pytest

import envelope_lib as E
from oracle_lib import MODE_COLORED, CpuSim

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
DT = E.DT


def bits_equal_report(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0))
    return int((~same.all(1)).sum()), float(np.nanmax(np.abs(a - b)))


@pytest.mark.parametrize("scene", [0, 1, 2, 3])
def test_gpu_trajectory_inside_the_reference_envelope(scene):
    import nbodysimulation_experiment_b200 as pkg

    env = E.load()
    gpu = pkg.ParticleSimulation()
    gpu.LoadScenario(scene, seed=1)
    init = np.load(os.path.join(GOLDEN_DIR, f"scene{scene}.npz"))["init"]
    bad, _ = bits_equal_report(gpu.particles(), init)
    assert bad == 0, "the initial state must equal the reference's (glibc rand() jitter, seed 1)"
    done = 0
    for k in (1, 8, 64, 256):
        for _ in range(k - done):
            gpu.Update(DT)
        done = k
        print(E.check_aggregates(env, scene, k, gpu.particles()))
    gpu.close()


def mirror(gpu, gravity, mode=MODE_COLORED, threads=1):
    """the CPU oracle with the GPU simulation's domain, parameters, bodies and particles (creation order kept)"""
    from nbodysimulation_experiment_b200 import scenes

    cpu = CpuSim("oracle", width=gpu.scene["width"], height=gpu.scene["height"], cell=scenes.KERNEL_HEIGHT, mode=mode, threads=threads)
    assert cpu.dims() == gpu.grid_dims()
    cpu.put_params(gpu.params_array())
    cpu.set_gravity(float(np.float32(gravity[0])), float(np.float32(gravity[1])))
    cpu.add_bodies(gpu.bodies)
    cpu.add_particle_array(gpu.particles()[:, 0:2])
    return cpu


def test_bodies_scene_bitwise_vs_oracle():
    """BASELINE.json configs[4] in small (sph.h:420-436 style bodies, 10x viscosity, dense block): 128 x 128 particles,
    48 steps, every particle's full ParticleData row equal to the oracle's coloured mode bit for bit."""
    from nbodysimulation_experiment_b200 import scenes

    g = (0.0, -3.0)
    gpu = scenes.fill_block(scenes.bodies_scene(128, gravity=g))
    assert gpu.GetParticleCount() == 128 * 128 and len(gpu.bodies) == 4 + 3 + 2
    cpu = mirror(gpu, g)
    for step in range(48):
        gpu.Update(DT)
        cpu.advance(DT)
        if step in (0, 7, 23):
            bad, diff = bits_equal_report(gpu.particles(), cpu.particles())
            assert bad == 0, f"step {step + 1}: {bad} particles differ, max abs diff {diff:.3e}"
    a, b = gpu.particles(), cpu.particles()
    bad, diff = bits_equal_report(a, b)
    assert bad == 0, f"after 48 steps: {bad} particles differ, max abs diff {diff:.3e}"
    assert np.array_equal(gpu.cell_counts(), cpu.cell_counts())
    # the bodies did something: particles rest against the three circles' skin
    moved = np.abs(a[:, 0:2] - gpu.particles()[:, 2:4]).max()
    assert np.isfinite(a).all() and moved > 0
    gpu.close()
    cpu.close()


def test_million_particle_block_bitwise_vs_oracle():
    """BASELINE.json configs[2] at FULL size: 1 048 576 particles, gravity (0,-10), 3 steps (step 1 has no viscosity pass,
    steps 2-3 exercise the stale-grid viscosity sweep), GPU one-launch sweep == oracle coloured mode, bit for bit."""
    from nbodysimulation_experiment_b200 import scenes

    g = (0.0, -10.0)
    gpu = scenes.fill_block(scenes.block_scene(1024, spacing=0.1, gravity=g))
    assert gpu.GetParticleCount() == 1024 * 1024
    cpu = mirror(gpu, g)
    for _ in range(3):
        gpu.Update(DT)
        cpu.advance(DT)
    bad, diff = bits_equal_report(gpu.particles(), cpu.particles())
    assert bad == 0, f"1M block after 3 steps: {bad} particles differ, max abs diff {diff:.3e}"
    assert np.array_equal(gpu.cell_counts(), cpu.cell_counts())
    gpu.close()
    cpu.close()
