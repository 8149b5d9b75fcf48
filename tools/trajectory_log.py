"""Steps a block scene on cuda:0 and logs bulk statistics every `--every` steps (aborts if the state
stops being finite or the step time explodes).  Used to pick physically sane benchmark scenes."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimulation_experiment_b200 import SPH_SOLVER_COLORED_GS, SPH_SOLVER_GATHER, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--spacing", type=float, default=0.1)
ap.add_argument("--gravity", type=float, default=None, help="default: -10 scaled to the reference's hydrostatic head")
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--every", type=int, default=50)
ap.add_argument("--solver", default="gs")
ap.add_argument("--relaxation", type=float, default=1.0)
ap.add_argument("--flags", type=int, default=0, help="SPH_FLAG_* (2 = no CUDA graphs)")
ap.add_argument("--bodies", action="store_true", help="BASELINE.json configs[4]: circles + tilted boxes, 10x viscosity")
ap.add_argument("--viscosity-scale", type=float, default=0.0, help="override: multiply the default linear/quadratic viscosity")
a = ap.parse_args()
g = a.gravity if a.gravity is not None else -10.0 * min(1.0, 5.34375 / (a.nx * a.spacing))
make = scenes.bodies_scene if a.bodies else scenes.block_scene
sim = scenes.fill_block(make(a.nx, spacing=a.spacing, gravity=(0.0, g), relaxation=a.relaxation, flags=a.flags,
                             solver=SPH_SOLVER_GATHER if a.solver == "gather" else SPH_SOLVER_COLORED_GS))
if a.viscosity_scale:
    p = sim.GetParams()
    p.linear_viscosity, p.quadratic_viscosity = 0.5 * a.viscosity_scale, 0.3 * a.viscosity_scale
    sim.SetParams(p)
n = sim.GetParticleCount()
dt = float(np.float32(1) / np.float32(60))
print(f"n {n} gravity {g:.4f} solver {a.solver}", flush=True)
for k in range(a.steps // a.every):
    sim.mark(0)
    for _ in range(a.every):
        sim.Update(dt)
    sim.mark(1)
    ms = sim.elapsed_ms(0, 1) / a.every
    st = sim.GetStats()
    p = sim.particles()
    ok = bool(np.isfinite(p).all())
    print(f"step {(k + 1) * a.every}: {ms:.3f} ms/step ({n / ms / 1e6:.3f} G particle-steps/s)  cand/particle {st.pair_candidates / n:.0f} "
          f"[{st.min_particle_neighbor_count}, {st.max_particle_neighbor_count}]  max cell {st.max_cell_particle_count}  |v|max {np.abs(p[:, 6:8]).max():.2f}  "
          f"rho mean {p[:, 8].mean():.1f} max {p[:, 8].max():.1f}  KE {0.5 * (p[:, 6:8].astype(np.float64) ** 2).sum():.4g}  finite {ok}", flush=True)
    if not ok or ms > 25.0:
        print("ABORT: state not finite or step time exploded", flush=True)
        break
