"""How a scene's step time evolves: every `--every` steps prints the device time per step (CUDA events over the window),
the longest candidate list, the fullest cell, candidates per particle, and a histogram of the cell occupancy.

    python tools/regime_probe.py --steps 80 --every 8            # the 1M block under g = -10 (bench.py's default scene)

Used to see what the collapse of the column does to the sweeps (heavy cells, staging capacity).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimulation_experiment_b200 import _lib, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--spacing", type=float, default=0.1)
ap.add_argument("--gravity", type=float, default=-10.0)
ap.add_argument("--steps", type=int, default=80)
ap.add_argument("--every", type=int, default=8)
ap.add_argument("--sweep", default="auto", choices=["auto", "flow", "warp", "team"])
ap.add_argument("--sweep-capacity", type=int, default=0)
ap.add_argument("--bodies", action="store_true")
ap.add_argument("--phases", action="store_true", help="per-phase times (serialises host and device)")
a = ap.parse_args()
flags = {"auto": 0, "flow": _lib.SPH_FLAG_SWEEP_FLOW, "warp": _lib.SPH_FLAG_SWEEP_WARP, "team": _lib.SPH_FLAG_SWEEP_TEAM}[a.sweep]
if a.phases:
    flags |= _lib.SPH_FLAG_PHASE_TIMING
build = scenes.bodies_scene if a.bodies else scenes.block_scene
kw = dict(spacing=a.spacing, flags=flags, sweep_capacity=a.sweep_capacity)
if not a.bodies:
    kw["gravity"] = (0.0, a.gravity)
sim = scenes.fill_block(build(a.nx, **kw))
dt = float(np.float32(1.0) / np.float32(60.0))
n = sim.GetParticleCount()
done = 0
while done < a.steps:
    k = min(a.every, a.steps - done)
    sim.ResetStats()
    sim.Sync()
    sim.mark(0)
    for _ in range(k):
        sim.Update(dt)
    sim.mark(1)
    ms = sim.elapsed_ms(0, 1) / k
    done += k
    st = sim.GetStats()
    cc = sim.cell_counts()
    occ = cc[cc > 0]
    hist = np.bincount(np.minimum(occ, 255) // 8, minlength=8)
    line = (f"step {done:4d}  {ms:7.4f} ms/step  {n / ms / 1e6:7.1f} M p-steps/s  cand/particle {st.pair_candidates / n:6.1f}  longest list {st.max_particle_neighbor_count:5d}"
            f"  fullest cell {occ.max():4d}  occupied cells {len(occ):7d}  cells by occupancy/8 {hist[:12].tolist()}")
    if a.phases:
        ph, _ = sim.phase_ms()
        line += "  " + " ".join(f"{key[:4]} {v:.3f}" for key, v in ph.items())
    print(line, flush=True)
sim.close()
