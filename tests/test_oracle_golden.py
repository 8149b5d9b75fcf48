"""CPU: the oracle restatement (oracle/sph_oracle.c, gs_index mode, one thread) against the golden
dumps of the reference's own demo4.cpp (tests/golden/, made by tools/make_golden.py).  Bit-exact."""
import glob
import os

import numpy as np
import pytest

from oracle_lib import MODE_GS_INDEX, CpuSim

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "scene*.npz")))


def candidate_checksums(sim):
    sums = np.zeros(sim.n, np.uint64)
    xors = np.zeros(sim.n, np.uint32)
    for i in range(sim.n):
        nb = sim.neighbors(i)
        sums[i] = nb.astype(np.uint64).sum()
        xors[i] = np.bitwise_xor.reduce(nb * np.uint32(2654435761)) if len(nb) else 0
    return sums, xors


def test_fixtures_present():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_dump(path):
    g = np.load(path)
    sim = CpuSim("oracle", mode=MODE_GS_INDEX, threads=1)
    sim.load_scenario(int(g["scene"]), int(g["seed"]))
    assert sim.dims() == (33, 18)  # kSPHGridCountX/Y, sph.h:61-62
    assert np.array_equal(sim.params(), g["params"])
    assert np.array_equal(sim.gravity(), g["gravity"])
    assert np.array_equal(sim.particles(), g["init"])
    for k, (t, nv, f) in enumerate(sim.bodies()):
        assert np.array_equal(np.concatenate([[t, nv], f]).astype(np.float32), g[f"body{k}"])
    assert f"body{len(sim.bodies())}" not in g.files
    dt, done = float(g["dt"]), 0
    for s in g["steps"]:
        sim.advance(dt, int(s) - done)
        done = int(s)
        assert np.array_equal(sim.particles(), g[f"state{s}"]), f"state after {s} steps"
    assert np.array_equal(sim.cell_of_particle(), g["cell_of_particle"])
    assert np.array_equal(sim.cell_counts(), g["cell_counts"])
    assert np.array_equal(sim.neighbor_counts(), g["neighbor_counts"])
    sums, xors = candidate_checksums(sim)
    assert np.array_equal(sums, g["cand_sum"]) and np.array_equal(xors, g["cand_xor"])
    assert np.array_equal(sim.stats()[0], g["stats"])
    assert np.array_equal(sim.colors(), g["colors"])
    # per-pass pair: state X -> re-file, NeighborSearch, DensityAndPressure
    sim.put_particles(sim.particles())
    sim.pass_neighbor_search()
    sim.pass_density()
    assert np.array_equal(sim.particles()[:, 8:12], g["density_from_last"])
    assert np.array_equal(sim.cell_of_particle(), g["refiled_cell_of_particle"])
    assert np.array_equal(sim.cell_counts(), g["refiled_cell_counts"])
    assert np.array_equal(sim.neighbor_counts(), g["refiled_neighbor_counts"])
    sums, xors = candidate_checksums(sim)
    assert np.array_equal(sums, g["refiled_cand_sum"]) and np.array_equal(xors, g["refiled_cand_xor"])
    sim.close()
