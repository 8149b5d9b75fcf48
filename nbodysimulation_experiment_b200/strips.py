"""Host-side arithmetic of the y-strip decomposition: which grid rows a rank owns,
which particles it keeps or copies to a neighbour after the predict phase, and how the per-rank
results merge back into one creation-ordered set.

The DEVICE applies the keep/send rule inside `predict_key_kernel` (csrc/sph_kernels.cuh); the
functions here are its host mirror, used to plan buffer capacities, by bench.py and the tools, and
by the CPU (gloo) tests of the multi-rank protocol.  Nothing here touches particle physics.
"""
import numpy as np

DEFAULT_HALO_ROWS = 7  # keep in sync with kDefaultHaloRows in csrc/sphb200.cu


def cell_rows(y, half_height, cell, grid_y):
    """Row of SPHComputeCellIndex (sph.h:450-463) for an array of y coordinates, in float32."""
    y = np.asarray(y, np.float32)
    r = ((y + np.float32(half_height)) / np.float32(cell)).astype(np.int64)  # (int) truncates toward zero
    return np.clip(r, 0, grid_y - 1)


def split_rows(occupied_rows, grid_y, world):
    """Even split of the first `occupied_rows` grid rows into `world` strips; the last strip also owns
    the (empty) rows above.  Returns [(row_begin, row_end)] with row_end exclusive."""
    occupied_rows = max(min(int(occupied_rows), grid_y), world)
    cuts = [int(round(occupied_rows * r / world)) for r in range(world)] + [grid_y]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def window(own, halo, grid_y, world):
    """Rows a rank holds locally: its own rows plus `halo` ghost rows per side, clipped to the grid."""
    lo, hi = own
    return (max(0, lo - halo), min(grid_y, hi + halo)) if world > 1 else (0, grid_y)


def classify(rows, rank, world, own, halo, grid_y):
    """The keep/send rule for AUTHORITATIVE particles whose new cell row is `rows`:
    keep   - row inside this rank's window;
    down   - row inside the lower neighbour's window (row < own_lo + halo);
    up     - row inside the upper neighbour's window (row >= own_hi - halo);
    lost   - none of the above (moved farther than a neighbour can take in one step)."""
    rows = np.asarray(rows)
    lo, hi = own
    wlo, whi = window(own, halo, grid_y, world)
    keep = (rows >= wlo) & (rows < whi)
    down = (rows < lo + halo) if rank > 0 else np.zeros(rows.shape, bool)
    up = (rows >= hi - halo) if rank + 1 < world else np.zeros(rows.shape, bool)
    lost = ~(keep | down | up)
    return keep, down, up, lost


def owned(rows, own):
    rows = np.asarray(rows)
    return (rows >= own[0]) & (rows < own[1])


def merge_owned(parts, total):
    """parts = [(ids, records)] from every rank -> records in creation order; checks that ownership
    is a partition of the particles."""
    ids = np.concatenate([np.asarray(p[0], np.int64) for p in parts])
    rec = np.concatenate([np.asarray(p[1]) for p in parts])
    if len(ids) != total or len(np.unique(ids)) != total or (len(ids) and (ids.min() < 0 or ids.max() >= total)):
        raise ValueError(f"ownership is not a partition: {len(ids)} records, {len(np.unique(ids))} distinct ids, expected {total}")
    out = np.zeros((total,) + rec.shape[1:], rec.dtype)
    out[ids] = rec
    return out


def halo_capacity_estimate(n_per_rank, rows_per_rank, halo, safety=3.0):
    """Records one direction of the exchange can carry per step: the particles of `halo` rows at the
    strip's mean density, times a safety factor for compression during the run."""
    per_row = n_per_rank / max(rows_per_rank, 1)
    return int(per_row * halo * safety) + 4096


def plan_bounds(row_counts, old_bounds, halo, max_shift=2):
    """New strip boundaries from the global row histogram - the host mirror of plan_strip_bounds in
    csrc/sphb200.cu, same integer arithmetic.  old_bounds = [B_0 = 0, B_1, ..., B_world = grid_y]; rank r owns rows
    [B_r, B_r+1).  Boundary r goes where the prefix of the counts reaches r/world of the total, but moves at most
    max_shift rows, stays `halo` rows inside the old ranges of the two ranks it separates (every row of a rank's new
    window then belongs to the rank itself or to a direct neighbour: one neighbour exchange moves everything) and
    leaves every strip at least 2*halo + 4 rows tall.  Returns the old bounds when that cannot be met."""
    counts = [int(c) for c in row_counts]
    old = [int(b) for b in old_bounds]
    world, gy = len(old) - 1, len(counts)
    min_rows = 2 * halo + 4
    total = sum(counts)
    nb = list(old)
    if total == 0:
        return nb
    prefix, row = 0, 0
    for b in range(1, world):
        target = total * b // world
        while row < gy and prefix < target:
            prefix += counts[row]
            row += 1
        want = row
        want = max(want, old[b] - max_shift)
        want = min(want, old[b] + max_shift)
        want = max(want, old[b - 1] + halo)
        want = min(want, old[b + 1] - halo)
        nb[b] = want
    for b in range(1, world):
        nb[b] = max(nb[b], nb[b - 1] + min_rows)
    for b in range(world - 1, 0, -1):
        nb[b] = min(nb[b], nb[b + 1] - min_rows)
    if any(nb[b] - nb[b - 1] < min_rows for b in range(1, world + 1)):
        return old
    # the fix-up passes may have pushed a boundary past the shift / halo limits (strips that start thinner than
    # min_rows): never trade a correct split for a better balanced one
    if any(abs(nb[b] - old[b]) > max_shift or nb[b] < old[b - 1] + halo or nb[b] > old[b + 1] - halo for b in range(1, world)):
        return old
    return nb
