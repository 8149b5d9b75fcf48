"""The 1M-particle dam-break block of BASELINE.json configs[2] under the reference's gravity (0,-10), stepped by the CPU
oracle in the REFERENCE's own semantics (gs_index, multithreaded like demo 4), logging bulk statistics: shows what the
102-unit column does under a fixed dt = 1/60 independently of anything on the GPU ("commit the log that
shows the g = -10 column diverging in the oracle too").

    python tools/oracle_column_log.py --steps 300 --every 10 > profiles/r2_oracle_column_g10.log
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle_lib import MODE_GS_INDEX, build_oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=1024)
ap.add_argument("--gravity", type=float, default=-10.0)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--every", type=int, default=10)
a = ap.parse_args()
build_oracle()
threads = os.cpu_count() or 1
sim = bench.oracle_scene(a.nx, 0.1, (0.0, a.gravity), threads, MODE_GS_INDEX)
n = sim.n
print(f"oracle gs_index, {threads} threads, {a.nx}x{a.nx} = {n} particles, spacing 0.1, gravity (0, {a.gravity}), dt = 1/60", flush=True)
done = 0
while done < a.steps:
    t0 = time.perf_counter()
    sim.advance(bench.DT, a.every)
    secs = time.perf_counter() - t0
    done += a.every
    p = sim.particles()
    counts = sim.cell_counts()
    ok = bool(np.isfinite(p).all())
    speed = np.sqrt((p[:, 6:8].astype(np.float64) ** 2).sum(1))
    print(f"step {done:4d}: {secs / a.every * 1e3:8.1f} ms/step  |v| max {np.nanmax(speed):10.2f} mean {np.nanmean(speed):7.3f}  rho max {np.nanmax(p[:, 8]):10.1f}  "
          f"fullest cell {int(counts.max()):6d}  y min {np.nanmin(p[:, 1]):9.3f}  finite {ok}", flush=True)
    if not ok:
        print("STOP: the state is no longer finite", flush=True)
        break
sim.close()
