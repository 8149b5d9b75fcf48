"""Generates tests/golden/*.npz from the UNMODIFIED reference solver (oracle/_ref/libsphref.so).

The reference ships no tests or golden vectors, so the pins are made here: its
own demo4.cpp, compiled headless by oracle/Makefile, run single-threaded (the only deterministic
mode), and dumped.  Only runnable where /root/reference exists; the outputs are committed.

    python tools/make_golden.py

Per scene: the initial state after LoadScenario (captures the glibc rand() jitter, seed 1), full
ParticleData dumps after a few Update(1/60) calls, and at the last dump the integer side of the
grid (cell of each particle, per-cell counts, candidate-list length and an order-independent
checksum of every candidate set) plus a per-pass pair: state X -> NeighborSearch +
DensityAndPressure -> (rho, rhoNear, P, PNear).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import CpuSim, build_oracle, have_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DT = float(np.float32(1.0) / np.float32(60.0))
SEED = 1
PLAN = {0: [1, 2, 8], 1: [1, 8, 32], 2: [1, 8, 32], 3: [1, 8, 32], 5: [200], 7: [100, 200]}


def candidate_checksums(sim):
    n = sim.n
    sums = np.zeros(n, np.uint64)
    xors = np.zeros(n, np.uint32)
    for i in range(n):
        nb = sim.neighbors(i)
        sums[i] = nb.astype(np.uint64).sum()
        xors[i] = np.bitwise_xor.reduce((nb * np.uint32(2654435761)) if len(nb) else np.zeros(1, np.uint32))
    return sums, xors


def main():
    build_oracle()
    assert have_ref(), "libsphref.so missing: needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for scene, steps in PLAN.items():
        sim = CpuSim("ref", threads=1)
        sim.load_scenario(scene, SEED)
        data = {
            "scene": np.int32(scene), "seed": np.int32(SEED), "dt": np.float32(DT), "steps": np.array(steps, np.int32),
            "params": sim.params(), "gravity": sim.gravity(), "init": sim.particles(),
        }
        for k, (t, nv, f) in enumerate(sim.bodies()):
            data[f"body{k}"] = np.concatenate([[t, nv], f]).astype(np.float32)
        done = 0
        for s in steps:
            sim.advance(DT, s - done)
            done = s
            data[f"state{s}"] = sim.particles()
        data["cell_of_particle"] = sim.cell_of_particle()
        data["cell_counts"] = sim.cell_counts()
        data["neighbor_counts"] = sim.neighbor_counts()
        data["cand_sum"], data["cand_xor"] = candidate_checksums(sim)
        data["stats"] = sim.stats()[0]
        data["colors"] = sim.colors()
        # per-pass pair from the last state: re-file, list, density
        x = sim.particles()
        sim.put_particles(x)
        sim.neighbor_search(DT)
        sim.density_pressure(DT)
        data["density_from_last"] = sim.particles()[:, 8:12].copy()
        # the integer side of that re-filed grid (the step's own grid above was filed from the
        # PREDICTED positions, demo4.cpp:342-356, so it is stale with respect to the final state)
        data["refiled_cell_of_particle"] = sim.cell_of_particle()
        data["refiled_cell_counts"] = sim.cell_counts()
        data["refiled_neighbor_counts"] = sim.neighbor_counts()
        data["refiled_cand_sum"], data["refiled_cand_xor"] = candidate_checksums(sim)
        path = os.path.join(OUT, f"scene{scene}.npz")
        np.savez_compressed(path, **data)
        print(f"scene {scene}: n={sim.n} steps={steps} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
        sim.close()


if __name__ == "__main__":
    main()
