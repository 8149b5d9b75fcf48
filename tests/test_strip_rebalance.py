"""Strip re-balancing, the host side: the planner that turns the global row histogram into new strip
boundaries.  The C++ planner inside libsphb200.so (sph_plan_strip_bounds, pure host arithmetic, no device) must agree
with its Python mirror (strips.plan_bounds) exactly - every rank plans for itself and all plans must coincide - and
the plan must keep the guarantees the one-exchange migration relies on.  Then the keep/send rule of a re-balancing step
(authority by the OLD rows, keep/send by the NEW ones) is replayed on numpy arrays: ownership stays a partition and
every rank ends up holding its whole new window."""
import ctypes
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from nbodysimulation_experiment_b200 import strips  # noqa: E402


def c_plan(counts, old, halo, max_shift):
    from nbodysimulation_experiment_b200 import build

    lib = ctypes.CDLL(build.build())
    counts = np.ascontiguousarray(counts, np.uint32)
    old = np.ascontiguousarray(old, np.int32)
    out = np.zeros(len(old), np.int32)
    rc = lib.sph_plan_strip_bounds(counts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(len(counts)), old.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_int32(len(old) - 1), ctypes.c_int32(halo), ctypes.c_int32(max_shift), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out.tolist()


def histograms():
    rng = np.random.default_rng(7)
    gy = 383
    block = np.zeros(gy, np.int64)
    block[:171] = 1536  # the 512 x 512 block of tools/mgpu_check.py at rest
    yield "block", block
    collapsed = np.zeros(gy, np.int64)
    collapsed[:60] = np.linspace(9000, 1500, 60).astype(np.int64)  # after the collapse: everything near the floor
    yield "collapsed", collapsed
    yield "random", rng.integers(0, 4000, gy)
    yield "empty", np.zeros(gy, np.int64)
    spike = np.zeros(gy, np.int64)
    spike[200] = 10 ** 6
    yield "one heavy row", spike


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("halo,max_shift", [(7, 2), (7, 8), (3, 1)])
def test_planner_matches_its_mirror_and_keeps_the_guarantees(world, halo, max_shift):
    gy = 383
    for name, counts in histograms():
        old = [round(171 * r / world) for r in range(world)] + [gy]  # the even split of the block's rows
        for it in range(400):  # iterate: re-balancing moves a little every time
            py = strips.plan_bounds(counts, old, halo, max_shift)
            assert py == c_plan(counts, old, halo, max_shift), (name, it, old)
            assert py[0] == 0 and py[-1] == gy
            if py != old:
                heights = np.diff(py)
                assert (heights >= 2 * halo + 4).all(), (name, py)
                for b in range(1, world):
                    assert abs(py[b] - old[b]) <= max_shift
                    # every row of a rank's new window belongs (old ownership) to itself or a direct neighbour
                    assert py[b] - halo >= old[b - 1] and py[b] + halo <= old[b + 1], (name, old, py)
            old = py
        if counts.sum() and name in ("block", "random") and world <= 4:
            loads = [int(counts[old[r]:old[r + 1]].sum()) for r in range(world)]
            assert max(loads) <= counts.sum() / world + counts.max() * (2 * halo + 4 + max_shift), (name, loads)  # converged near the even share


def test_thin_strips_are_left_alone_rather_than_jumped_over():
    """Strips that START thinner than 2*halo + 4 rows: the min-rows fix-up used to run after the shift / halo clamps
    and could push a boundary past a non-neighbour's rows (old = [0, 7, 14, 60] gave [0, 18, 36, 60]: rows 14..17
    moved from rank 2 to rank 0 in one step and their particles vanished).  Such a plan must be refused."""
    halo, max_shift = 7, 2
    for old in ([0, 7, 14, 60], [0, 10, 20, 30, 40, 50, 60, 70, 86], [0, 9, 30, 60]):
        gy = old[-1]
        for counts in (np.full(gy, 100), np.arange(gy) * 10, np.r_[np.full(gy // 2, 1000), np.zeros(gy - gy // 2, np.int64)]):
            py = strips.plan_bounds(counts, old, halo, max_shift)
            assert py == c_plan(counts, old, halo, max_shift)
            for b in range(1, len(old) - 1):
                assert abs(py[b] - old[b]) <= max_shift, (old, py)
                assert py[b] - halo >= old[b - 1] and py[b] + halo <= old[b + 1], (old, py)
    assert strips.plan_bounds(np.full(60, 100), [0, 7, 14, 60], 7, 2) == [0, 7, 14, 60]


def test_rebalancing_step_keeps_ownership_a_partition():
    """Particles scattered over the rows; two ranks; the boundary moves by two rows.  Authority follows the old
    rows, keep / send the new ones (predict_key_kernel with StripDesc::authLo/authHi): afterwards every particle has
    exactly one owner and each rank holds every particle of its new window."""
    rng = np.random.default_rng(3)
    gy, halo, world = 120, 7, 2
    old = [0, 60, gy]
    rows_before = rng.integers(0, 100, 20000)  # row on the previous grid
    rows_after = np.clip(rows_before + rng.integers(-1, 2, len(rows_before)), 0, gy - 1)  # after predict: moved by <= 1 row
    for new in ([0, 58, gy], [0, 62, gy], [0, 60, gy]):
        held, owners = [], np.zeros(len(rows_before), np.int64)
        for rank in range(world):
            auth = strips.owned(rows_before, (old[rank], old[rank + 1]))
            keep, down, up, lost = strips.classify(rows_after[auth], rank, world, (new[rank], new[rank + 1]), halo, gy)
            assert not lost.any()
            held.append({"keep": np.flatnonzero(auth)[keep], "down": np.flatnonzero(auth)[down], "up": np.flatnonzero(auth)[up]})
        for rank in range(world):
            got = [held[rank]["keep"]]
            if rank > 0:
                got.append(held[rank - 1]["up"])
            if rank + 1 < world:
                got.append(held[rank + 1]["down"])
            got = np.concatenate(got)
            wlo, whi = strips.window((new[rank], new[rank + 1]), halo, gy, world)
            inside = got[(rows_after[got] >= wlo) & (rows_after[got] < whi)]  # unpack drops what is outside the window
            want = np.flatnonzero((rows_after >= wlo) & (rows_after < whi))
            assert np.array_equal(np.sort(inside), want), (new, rank)  # the whole new window, nothing twice
            owners[inside[strips.owned(rows_after[inside], (new[rank], new[rank + 1]))]] += 1
        assert (owners == 1).all()
