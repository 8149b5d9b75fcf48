"""The reference's built-in benchmark protocol (the `B` key: app.h:24-30, app.cpp:238-274, 88-161), headless,
with JSON instead of the on-screen bar chart.

For the chosen scenario: `iterations` x `frames` Update(1/60) calls (reference: 16 x 64), the scenario is
reloaded before every iteration (app.cpp:256-262), and every frame records the nine SPHStatistics.time
buckets (sph.h:131-141) plus the whole-Update time.  Per bucket the recorder keeps min / avg / max over all
frames of all iterations (app.cpp:88-161); the reference's chart shows the max (app.cpp:186-195).

    python tools/benchmark_recorder.py --scenario 0                 # our solver on cuda:0
    python tools/benchmark_recorder.py --scenario 0 --impl oracle   # the CPU restatement, reference semantics, all cores
    python tools/benchmark_recorder.py --scenario 0 --impl ref      # the reference's own binary (where libsphref.so exists)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BUCKETS = ("emitters", "integration", "viscosityForces", "predict", "updateGrid", "neighborSearch", "densityAndPressure", "deltaPositions", "collisions")
DT = float(np.float32(1.0) / np.float32(60.0))


def run_gpu(scenario, iterations, frames):
    from nbodysimulation_experiment_b200 import SPH_FLAG_PHASE_TIMING, ParticleSimulation

    sim = ParticleSimulation(flags=SPH_FLAG_PHASE_TIMING)
    rows, totals = [], []
    for it in range(iterations):
        sim.LoadScenario(scenario, seed=1)  # also ResetStats (app.cpp:480)
        for _ in range(frames):
            sim.ResetStats()  # one-frame averages = that frame's phase times
            t0 = time.perf_counter()
            sim.Update(DT)
            st = sim.GetStats()  # waits for the step, like the app reading GetStats() after Update (app.cpp:240)
            totals.append((time.perf_counter() - t0) * 1e3)
            rows.append([st.time_emitters, st.time_integration, st.time_viscosity_forces, st.time_predict, st.time_update_grid,
                         st.time_neighbor_search, st.time_density_and_pressure, st.time_delta_positions, st.time_collisions])
    n = sim.GetParticleCount()
    sim.close()
    return n, np.array(rows), np.array(totals)


def run_cpu(kind, scenario, iterations, frames):
    from oracle_lib import MODE_GS_INDEX, CpuSim

    threads = os.cpu_count() or 1
    sim = CpuSim(kind, mode=MODE_GS_INDEX, threads=threads) if kind == "oracle" else CpuSim("ref", threads=threads)
    rows, totals = [], []
    for it in range(iterations):
        sim.load_scenario(scenario, 1)
        for _ in range(frames):
            t0 = time.perf_counter()
            sim.advance(DT)
            totals.append((time.perf_counter() - t0) * 1e3)
            rows.append(sim.stats()[1].tolist())
    n = sim.n
    sim.close()
    return n, np.array(rows), np.array(totals)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=16)  # kBenchmarkIterationCount, app.h
    ap.add_argument("--frames", type=int, default=64)      # kBenchmarkFrameCount, app.h:24
    ap.add_argument("--impl", default="b200", choices=["b200", "oracle", "ref"])
    a = ap.parse_args()
    n, rows, totals = run_gpu(a.scenario, a.iterations, a.frames) if a.impl == "b200" else run_cpu(a.impl, a.scenario, a.iterations, a.frames)
    out = {"impl": a.impl, "scenario": a.scenario, "particles_at_end": int(n), "iterations": a.iterations, "frames": a.frames,
           "update_ms": {"min": float(totals.min()), "avg": float(totals.mean()), "max": float(totals.max())},
           "particle_steps_per_s_avg": float(n / (totals.mean() * 1e-3)),
           "phases_ms": {b: {"min": float(rows[:, k].min()), "avg": float(rows[:, k].mean()), "max": float(rows[:, k].max())} for k, b in enumerate(BUCKETS)}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
