"""Trajectory parity against the REFERENCE (SURVEY.md section 7 hard parts 1-2, section 8d c2), shared by the CPU test
of the oracle's coloured mode and the GPU test through the C ABI.

The reference's demo-4 path is deterministic only single-threaded; its shipped multithreaded mode races.  So the bar
is the reference's own MT-vs-ST spread, measured on the unmodified reference (libsphref.so) by tools/make_envelope.py and
committed as tests/golden/envelope.json:

  aggregates after K steps (K in 1, 8, 32, 64, 128, 256; scenes 0-3), against the single-thread reference run:
    kinetic energy   |KE/KE_st - 1|      <= max(1.5 x the worst of the three MT runs, 5 %)
    centre of mass   |com - com_st|_inf  <= max(2.5 x the worst MT run, 1e-3 world units)
    mean density     |rho/rho_st - 1|    <= max(2 x the worst MT run, 1 %)
    particle count   equal; every particle inside the walls the reference run stayed inside (+ one collision radius)

  one step from an injected identical state (golden state after 8 steps): coloured sweep vs the reference's index-order
  sweep, per particle:  max |dx| <= 2 x, mean |dx| <= 2.5 x the reference's own MT-vs-ST gap from the same state.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ENVELOPE = os.path.join(HERE, "golden", "envelope.json")
DT = float(np.float32(1.0) / np.float32(60.0))
KE_FACTOR, KE_FLOOR = 1.5, 0.05
COM_FACTOR, COM_FLOOR = 2.5, 1e-3
RHO_FACTOR, RHO_FLOOR = 2.0, 0.01
ONE_STEP_MAX_FACTOR, ONE_STEP_MEAN_FACTOR = 2.0, 2.5


def load():
    with open(ENVELOPE) as f:
        return json.load(f)


def aggregates(p):
    """(n, 12) ParticleData rows -> the same invariants tools/make_envelope.py stores"""
    x = p[:, 0:2].astype(np.float64)
    v = p[:, 6:8].astype(np.float64)
    return {
        "n": int(len(p)),
        "ke": float(0.5 * (v ** 2).sum()),
        "com": [float(x[:, 0].mean()), float(x[:, 1].mean())],
        "mean_rho": float(p[:, 8].astype(np.float64).mean()),
        "extent": [float(x[:, 0].min()), float(x[:, 0].max()), float(x[:, 1].min()), float(x[:, 1].max())],
    }


def check_aggregates(env, scene, k, p):
    """assert the state `p` of `scene` after `k` steps lies inside the reference's envelope; returns a report line"""
    sc = env["scenes"][str(scene)]
    st, mts = sc["st"][str(k)], [m[str(k)] for m in sc["mt"]]
    a = aggregates(p)
    assert a["n"] == st["n"], f"scene {scene} K={k}: {a['n']} particles, the reference has {st['n']}"
    assert np.isfinite(p).all()
    ke_dev = abs(a["ke"] / st["ke"] - 1.0)
    ke_tol = max(KE_FACTOR * max(abs(m["ke"] / st["ke"] - 1.0) for m in mts), KE_FLOOR)
    com_dev = float(np.abs(np.array(a["com"]) - np.array(st["com"])).max())
    com_tol = max(COM_FACTOR * max(float(np.abs(np.array(m["com"]) - np.array(st["com"])).max()) for m in mts), COM_FLOOR)
    rho_dev = abs(a["mean_rho"] / st["mean_rho"] - 1.0)
    rho_tol = max(RHO_FACTOR * max(abs(m["mean_rho"] / st["mean_rho"] - 1.0) for m in mts), RHO_FLOOR)
    line = (f"scene {scene} K={k:3d}: KE {a['ke']:10.1f} (ref {st['ke']:10.1f}, dev {ke_dev:.3f} <= {ke_tol:.3f})  com dev {com_dev:.2e} <= {com_tol:.2e}"
            f"  rho dev {rho_dev:.4f} <= {rho_tol:.4f}")
    assert ke_dev <= ke_tol, line
    assert com_dev <= com_tol, line
    assert rho_dev <= rho_tol, line
    lo_x, hi_x, lo_y, hi_y = st["extent"]
    ex = a["extent"]
    slack = 0.05 + 1e-4  # kSPHParticleCollisionRadius (sph.h:35,38): where the planes park a particle
    assert ex[0] >= min(lo_x, -5.0 + 0.05) - slack and ex[1] <= max(hi_x, 5.0 - 0.05) + slack, line
    assert ex[2] >= min(lo_y, -2.8125 + 0.05) - slack and ex[3] <= max(hi_y, 2.8125 - 0.05) + slack, line
    return line


def check_one_step_gap(env, scene, ours, reference_order):
    """`ours`: state one step after the golden state in the coloured order; `reference_order`: the same step in the
    reference's index order (the oracle's gs_index mode, which equals the reference bit for bit)"""
    ref = env["scenes"][str(scene)]["one_step_from_state8"]["mt_vs_st"]
    dx = np.sqrt(((ours[:, 0:2].astype(np.float64) - reference_order[:, 0:2]) ** 2).sum(1))
    ref_max, ref_mean = max(r["dx_max"] for r in ref), float(np.mean([r["dx_mean"] for r in ref]))
    line = (f"scene {scene}: one step, coloured vs index order: max |dx| {dx.max():.3e} (reference MT vs ST {ref_max:.3e}), "
            f"mean |dx| {dx.mean():.3e} (reference {ref_mean:.3e})")
    assert dx.max() <= ONE_STEP_MAX_FACTOR * ref_max, line
    assert dx.mean() <= ONE_STEP_MEAN_FACTOR * ref_mean, line
    return line
