"""Small driver for ncu: a few steps of the 1M dam-break scene (bench.py's workload) and nothing else."""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nbodysimulation_experiment_b200 import SPH_FP_EXACT, SPH_FP_FAST, SPH_SOLVER_COLORED_GS, SPH_SOLVER_GATHER, _lib, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--spacing", type=float, default=0.1)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--solver", default="gs")
ap.add_argument("--fp", default="exact")
ap.add_argument("--gravity", type=float, default=-10.0)
ap.add_argument("--sweep", default="auto", choices=["auto", "flow", "warp", "team"])
a = ap.parse_args()
sim = scenes.fill_block(scenes.block_scene(a.nx, spacing=a.spacing, gravity=(0.0, a.gravity), fp_mode=SPH_FP_FAST if a.fp == "fast" else SPH_FP_EXACT,
                                           solver=SPH_SOLVER_GATHER if a.solver == "gather" else SPH_SOLVER_COLORED_GS,
                                           flags={"auto": 0, "flow": _lib.SPH_FLAG_SWEEP_FLOW, "warp": _lib.SPH_FLAG_SWEEP_WARP, "team": _lib.SPH_FLAG_SWEEP_TEAM}[a.sweep]))
dt = float(np.float32(1.0) / np.float32(60.0))
for _ in range(a.steps):
    sim.Update(dt)
sim.Sync()
st = sim.GetStats()
print("particles", sim.GetParticleCount(), "candidates/particle", st.pair_candidates / sim.GetParticleCount(), "max cell", st.max_cell_particle_count)
