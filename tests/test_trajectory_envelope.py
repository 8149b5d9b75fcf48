"""Trajectory parity of the coloured Gauss-Seidel order against the REFERENCE, on the CPU (the GPU is bit-identical to
the oracle's coloured mode, tests/test_gpu_parity.py; tests/test_gpu_envelope.py repeats the aggregate part through the
C ABI).  Bars and their derivation: tests/envelope_lib.py; numbers of the reference: tests/golden/envelope.json
(tools/make_envelope.py, from the unmodified demo4.cpp)."""
import os

import numpy as np
import pytest

import envelope_lib as E
from oracle_lib import MODE_COLORED, MODE_GS_INDEX, CpuSim

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def test_envelope_fixture_present():
    env = E.load()
    assert sorted(env["scenes"]) == ["0", "1", "2", "3"] and env["K"] == [1, 8, 32, 64, 128, 256]
    for sc in env["scenes"].values():
        assert len(sc["mt"]) >= 3 and len(sc["one_step_from_state8"]["mt_vs_st"]) >= 5


def one_step_from(scene, state, mode):
    sim = CpuSim("oracle", mode=mode)
    sim.load_scenario(scene, 1)
    sim.put_particles(state)
    sim.pass_neighbor_search()
    sim.advance(E.DT, 1)
    out = sim.particles()
    sim.close()
    return out


@pytest.mark.parametrize("scene", [0, 1, 2, 3])
def test_one_step_gap_to_the_reference_order_is_the_size_of_its_own_mt_scatter(scene):
    """demo4.cpp:223-255 from an injected reference state: one Update in the coloured order and in the reference's index
    order; the per-particle gap against the reference's own multithreaded-vs-single-threaded gap from that state."""
    env = E.load()
    state = np.load(os.path.join(GOLDEN_DIR, f"scene{scene}.npz"))["state8"]
    in_reference_order = one_step_from(scene, state, MODE_GS_INDEX)
    colored = one_step_from(scene, state, MODE_COLORED)
    print(E.check_one_step_gap(env, scene, colored, in_reference_order))
    # the single-threaded reference run from the same injected state, as stored by make_envelope: same aggregates
    st = env["scenes"][str(scene)]["one_step_from_state8"]["st_aggregates"]
    a = E.aggregates(in_reference_order)
    assert abs(a["ke"] / st["ke"] - 1.0) < 1e-5 and np.allclose(a["com"], st["com"], atol=1e-6)


@pytest.mark.parametrize("scene", [0, 1, 2, 3])
def test_coloured_order_stays_inside_the_reference_envelope(scene):
    """K in {1, 8, 32, 64, 128, 256} steps of scenes 0-3: kinetic energy, centre of mass, mean density,
    count and extent within the reference's own MT-vs-ST spread (factors in envelope_lib)."""
    env = E.load()
    sim = CpuSim("oracle", mode=MODE_COLORED)
    sim.load_scenario(scene, 1)
    done = 0
    for k in env["K"]:
        sim.advance(E.DT, k - done)
        done = k
        print(E.check_aggregates(env, scene, k, sim.particles()))
    sim.close()
