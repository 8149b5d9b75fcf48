"""The scheduling rule of the one-launch coloured sweep (color_sweep_flow_kernel, csrc/sph_kernels.cuh), modelled on
the CPU: occupied cells of all nine colours form one queue (colour 0's list in row-major order, then colour 1's, ...),
workers draw tickets from it in order, and a cell may start once every occupied cell of a LOWER colour within two
rows/columns has finished.  Checked here with an event-driven simulation of many workers and random cell durations:

  * the result equals the nine sequential colour sweeps for an update that does not commute (so any two cells whose
    3x3 footprints overlap really have to run in colour order);
  * nothing dead-locks, for any number of workers and for workers that hold one extra ticket (the kernel draws its
    next ticket before it publishes the current cell);
  * a dependency radius of one cell is NOT enough (the test fails it on purpose): footprints of cells two apart
    overlap in the cell between them.

No GPU, no oracle: this is host-side logic.
"""
import heapq

import numpy as np
import pytest


def color_of(cx, cy):
    return (cy % 3) * 3 + cx % 3


def build_queue(occ):
    """colour lists in row-major order, concatenated (color_rows_count/scan/fill_kernel)"""
    gy, gx = occ.shape
    lists = [[] for _ in range(9)]
    for cy in range(gy):
        for cx in range(gx):
            if occ[cy, cx]:
                lists[color_of(cx, cy)].append((cx, cy))
    return [c for lst in lists for c in lst]


def sweep_cell(state, cx, cy, stamp):
    """A stand-in for the pair sweep of one cell: reads and rewrites its 3x3 block with an order-sensitive update."""
    gy, gx = state.shape
    y0, y1, x0, x1 = max(cy - 1, 0), min(cy + 2, gy), max(cx - 1, 0), min(cx + 2, gx)
    block = state[y0:y1, x0:x1]
    mix = np.uint64(int(block.sum()) % (1 << 61))
    state[y0:y1, x0:x1] = (block * np.uint64(6364136223846793005) + mix + np.uint64(stamp)) % np.uint64(1 << 61)


def sequential(occ, init):
    state = init.copy()
    for cx, cy in build_queue(occ):  # nine launches, one colour after another = queue order
        sweep_cell(state, cx, cy, cy * occ.shape[1] + cx)
    return state


def dataflow(occ, init, workers, rng, radius=2, tickets_held=1, queue=None):
    """Event-driven run.  Returns (state, cells that had to wait).  Raises on dead-lock."""
    gy, gx = occ.shape
    queue = build_queue(occ) if queue is None else queue
    state = init.copy()
    done = np.zeros_like(occ, dtype=bool)
    next_ticket = 0
    waits = 0
    # a worker: list of tickets it holds (works them in order); event heap of (finish time, worker, cell)
    held = [[] for _ in range(workers)]
    running = []
    idle = list(range(workers))
    now = 0.0

    def deps_done(cx, cy):
        c = color_of(cx, cy)
        for ny in range(max(cy - radius, 0), min(cy + radius + 1, gy)):
            for nx in range(max(cx - radius, 0), min(cx + radius + 1, gx)):
                if (nx, ny) != (cx, cy) and occ[ny, nx] and color_of(nx, ny) < c and not done[ny, nx]:
                    return False
        return True

    waiting = set()
    while True:
        progressed = False
        rng.shuffle(idle)
        for w in list(idle):
            while len(held[w]) < tickets_held and next_ticket < len(queue):
                held[w].append(next_ticket)
                next_ticket += 1
            if not held[w]:
                continue
            cx, cy = queue[held[w][0]]
            if deps_done(cx, cy):
                # the cell reads its block when it starts and writes it when it ends; nothing that overlaps it may run
                # in between - which is exactly what the rule has to guarantee, so the update is applied at the start
                sweep_cell(state, cx, cy, cy * gx + cx)
                heapq.heappush(running, (now + rng.uniform(0.2, 3.0), w, (cx, cy)))
                idle.remove(w)
                waiting.discard(w)
                progressed = True
            elif w not in waiting:
                waiting.add(w)
                waits += 1
        if not running:
            if next_ticket >= len(queue) and not any(held):
                return state, waits
            if not progressed:
                raise RuntimeError("dead-lock: no cell running and no waiting cell can start")
            continue
        now, w, (cx, cy) = heapq.heappop(running)
        done[cy, cx] = True
        held[w].pop(0)
        idle.append(w)


@pytest.mark.parametrize("workers,tickets_held", [(1, 1), (3, 1), (17, 1), (64, 1), (17, 2), (200, 3)])
def test_dataflow_rule_equals_sequential_colours(workers, tickets_held):
    rng = np.random.default_rng(workers * 10 + tickets_held)
    occ = rng.random((14, 19)) < 0.7
    init = rng.integers(1, 1 << 40, occ.shape).astype(np.uint64)
    want = sequential(occ, init)
    for trial in range(4):
        got, _ = dataflow(occ, init, workers, np.random.default_rng(trial), radius=2, tickets_held=tickets_held)
        assert np.array_equal(got, want)


def test_radius_one_is_not_enough():
    """Cells two apart share the cell between them: waiting only for direct neighbours gives wrong results."""
    rng = np.random.default_rng(5)
    occ = np.ones((9, 12), bool)
    init = rng.integers(1, 1 << 40, occ.shape).astype(np.uint64)
    want = sequential(occ, init)
    wrong = 0
    for trial in range(6):
        got, _ = dataflow(occ, init, 40, np.random.default_rng(trial), radius=1)
        wrong += int(not np.array_equal(got, want))
    assert wrong > 0


def shuffled_queue(occ, rng):
    """each colour's list in random order: what lists appended with atomics look like"""
    gy, gx = occ.shape
    lists = [[] for _ in range(9)]
    for cy in range(gy):
        for cx in range(gx):
            if occ[cy, cx]:
                lists[color_of(cx, cy)].append((cx, cy))
    for lst in lists:
        rng.shuffle(lst)
    return [c for lst in lists for c in lst]


def test_row_major_lists_need_no_waiting_when_a_colour_outnumbers_the_workers():
    """With sorted lists the cells a newcomer depends on were handed out a whole colour earlier: as long as a colour's
    list is longer than the cells in flight, (almost) nobody waits.  The same rule on shuffled lists waits all the
    time (measured on the GPU: 12.5 % of the warp time) - and still gives the right result."""
    rng = np.random.default_rng(11)
    occ = np.ones((60, 60), bool)  # 400 cells per colour
    init = rng.integers(1, 1 << 40, occ.shape).astype(np.uint64)
    want = sequential(occ, init)
    got, waits_sorted = dataflow(occ, init, 100, np.random.default_rng(0))
    assert np.array_equal(got, want)
    got, waits_shuffled = dataflow(occ, init, 100, np.random.default_rng(0), queue=shuffled_queue(occ, np.random.default_rng(1)))
    assert np.array_equal(got, want)  # the order inside a colour never matters for the result
    assert waits_sorted < 0.02 * occ.sum()
    assert waits_shuffled > 10 * max(waits_sorted, 1)
