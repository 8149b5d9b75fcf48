"""Generates tests/golden/envelope.json from the UNMODIFIED reference solver (oracle/_ref/libsphref.so): the numbers the
trajectory-parity tests hold the coloured Gauss-Seidel solver to (SURVEY.md section 7 hard parts 1-2, section 8d c2).

The reference's demo-4 path is only deterministic single-threaded; its multithreaded mode races (threading.h:111-129,
demo4.cpp:223-255), so "matches the reference after K steps" is only definable as (i) the gap after ONE step from an
identical injected state, measured against the reference's own MT-vs-ST gap from that state, and (ii) aggregate
invariants (kinetic energy, centre of mass, mean density, extent) after K steps within the reference's own MT-vs-ST
spread.  This script measures both sides of (i) and (ii) on the reference itself:

  * per scene 0-3, aggregates of the single-thread run at K in {1, 8, 32, 64, 128, 256} and of three multithreaded runs;
  * from the committed golden state after 8 steps (tests/golden/scene<k>.npz, itself dumped from the reference):
    inject -> NeighborSearch -> Update(1/60) single-threaded and multithreaded (5 runs); max / mean |dx| and |dv|.

Only runnable where /root/reference exists (the multithreaded numbers also depend on the host's core count, recorded);
the output is committed.

    python tools/make_envelope.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import CpuSim, build_oracle, have_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "envelope.json")
DT = float(np.float32(1.0) / np.float32(60.0))
SEED = 1
KS = [1, 8, 32, 64, 128, 256]
MT_RUNS = 3
ONE_STEP_RUNS = 5


def aggregates(p):
    """p: (n, 12) ParticleData rows -> the invariants of SURVEY.md section 7 hard part 2 (iii)"""
    x = p[:, 0:2].astype(np.float64)
    v = p[:, 6:8].astype(np.float64)
    return {
        "n": int(len(p)),
        "ke": float(0.5 * (v ** 2).sum()),
        "com": [float(x[:, 0].mean()), float(x[:, 1].mean())],
        "mean_rho": float(p[:, 8].astype(np.float64).mean()),
        "mean_speed": float(np.sqrt((v ** 2).sum(1)).mean()),
        "extent": [float(x[:, 0].min()), float(x[:, 0].max()), float(x[:, 1].min()), float(x[:, 1].max())],
    }


def run(scene, threads):
    sim = CpuSim("ref", threads=threads)
    sim.load_scenario(scene, SEED)
    out, done = {}, 0
    for k in KS:
        sim.advance(DT, k - done)
        done = k
        out[str(k)] = aggregates(sim.particles())
    workers = sim.worker_threads()
    sim.close()
    return out, workers


def one_step(scene, state, threads):
    sim = CpuSim("ref", threads=threads)
    sim.load_scenario(scene, SEED)
    sim.put_particles(state)
    sim.neighbor_search(DT)
    sim.advance(DT, 1)
    p = sim.particles()
    sim.close()
    return p


def main():
    build_oracle()
    assert have_ref(), "libsphref.so missing: needs /root/reference"
    doc = {"what": "reference demo-4 solver (libsphref.so, unmodified demo4.cpp): aggregates after K steps, single-threaded and multithreaded, "
                   "and the one-step MT-vs-ST gap from the golden state after 8 steps",
           "dt": DT, "seed": SEED, "K": KS, "host_cores": os.cpu_count(), "scenes": {}}
    for scene in (0, 1, 2, 3):
        st, _ = run(scene, 1)
        mts = []
        for _ in range(MT_RUNS):
            mt, workers = run(scene, 8)
            mts.append(mt)
        g = np.load(os.path.join(ROOT, "tests", "golden", f"scene{scene}.npz"))
        state = g["state8"]
        a = one_step(scene, state, 1)
        again = one_step(scene, state, 1)
        assert np.array_equal(a, again), "the single-threaded reference must be deterministic"
        gaps = []
        for _ in range(ONE_STEP_RUNS):
            b = one_step(scene, state, 8)
            dx = np.sqrt(((a[:, 0:2].astype(np.float64) - b[:, 0:2]) ** 2).sum(1))
            dv = np.sqrt(((a[:, 6:8].astype(np.float64) - b[:, 6:8]) ** 2).sum(1))
            gaps.append({"dx_max": float(dx.max()), "dx_mean": float(dx.mean()), "dv_max": float(dv.max()), "dv_mean": float(dv.mean())})
        doc["scenes"][str(scene)] = {"st": st, "mt": mts, "mt_workers": workers, "one_step_from_state8": {"st_aggregates": aggregates(a), "mt_vs_st": gaps}}
        print(f"scene {scene}: n={st['1']['n']} KE(ST) " + " ".join(f"{k}:{st[str(k)]['ke']:.0f}" for k in KS))
        for mt in mts:
            print("          KE(MT) " + " ".join(f"{k}:{mt[str(k)]['ke']:.0f}" for k in KS))
        print("          one step MT vs ST: dx max " + " ".join(f"{q['dx_max']:.2e}" for q in gaps) + " | mean " + " ".join(f"{q['dx_mean']:.2e}" for q in gaps))
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print("->", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
