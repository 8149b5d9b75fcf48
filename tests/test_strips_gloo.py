"""CPU: the multi-rank (y-strip) host logic, exercised with two real processes over the gloo backend.

The device applies the keep/send rule in predict_key_kernel and ships records with NCCL; here the same
rule (nbodysimulation_experiment_b200.strips) runs on numpy arrays and the records travel over
torch.distributed/gloo, so the protocol invariants are checked without a GPU:

  * after one exchange every particle has exactly one owner, and the owners' union is everything;
  * a rank's ghost set is exactly the neighbour's particles inside its window;
  * the merged per-rank results equal the single-rank array, in creation order;
  * the max-over-ranks timing reduction bench.py uses.

The particles are a reference dump (tests/golden/scene0.npz) displaced by its own velocities, i.e.
the state the device would classify after its predict phase.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from nbodysimulation_experiment_b200 import strips  # noqa: E402

WORLD = 2
GRID_Y, CELL, HALF_H = 18, float(np.float32(6.0) * np.float32(0.05)), 5.625 / 2
HALO = 3  # the 18-row reference grid is too short for the default 7-row halo


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def load_state():
    g = np.load(os.path.join(HERE, "golden", "scene0.npz"))
    p = g["state8"].copy()
    dt = np.float32(1.0) / np.float32(60.0)
    old_rows = strips.cell_rows(p[:, 1], HALF_H, CELL, GRID_Y)
    p[:, 2:4] = p[:, 0:2]
    p[:, 0:2] = p[:, 6:8] * dt + p[:, 0:2]  # predict (demo4.cpp:330-339)
    return p, old_rows


def worker(rank, port, out_dir, WORLD=WORLD):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    p, old_rows = load_state()
    n = len(p)
    ids = np.arange(n)
    rows_occupied = int(old_rows.max()) + 1
    own_all = strips.split_rows(rows_occupied, GRID_Y, WORLD)
    own = own_all[rank]
    # what this rank is authoritative for: the particles it owned on the previous grid
    mine = strips.owned(old_rows, own)
    new_rows = strips.cell_rows(p[mine, 1], HALF_H, CELL, GRID_Y)
    keep, down, up, lost = strips.classify(new_rows, rank, WORLD, own, HALO, GRID_Y)
    assert not lost.any()
    my_ids, my_rec = ids[mine], p[mine]

    def pack(mask):
        return torch.from_numpy(np.concatenate([my_ids[mask, None].astype(np.float64), my_rec[mask].astype(np.float64)], 1))

    # one exchange: counts first (the device puts them in the message header), then the records
    recv = []
    for peer, mask in ((rank - 1, down), (rank + 1, up)):
        if not 0 <= peer < WORLD:
            continue
        send_buf = pack(mask)
        cnt = torch.tensor([len(send_buf)])
        other = torch.zeros(1, dtype=torch.long)
        reqs = [dist.isend(cnt, peer), dist.irecv(other, peer)]
        for r in reqs:
            r.wait()
        got = torch.zeros((int(other.item()), 13), dtype=torch.float64)
        reqs = [dist.isend(send_buf, peer), dist.irecv(got, peer)]
        for r in reqs:
            r.wait()
        recv.append(got.numpy())
    got = np.concatenate(recv) if recv else np.zeros((0, 13))
    # where the device files what arrived (unpack_kernel): behind the kept particles, the lower neighbour's records first,
    # the upper neighbour's behind THEM - an interior strip (two neighbours) is the case two ranks never exercise
    n_keep = int(keep.sum())
    offsets = np.cumsum([n_keep] + [len(r) for r in recv])
    assert offsets[-1] == n_keep + len(got) and len(recv) == (rank > 0) + (rank + 1 < WORLD)
    if 0 < rank < WORLD - 1:
        assert len(recv) == 2 and len(recv[0]) > 0 and len(recv[1]) > 0, "an interior strip must hear from both neighbours"
        lower_rows = strips.cell_rows(recv[0][:, 2].astype(np.float32), HALF_H, CELL, GRID_Y)
        upper_rows = strips.cell_rows(recv[1][:, 2].astype(np.float32), HALF_H, CELL, GRID_Y)
        assert lower_rows.max() < upper_rows.min(), "records of the two neighbours come from opposite ends of the window"
    local_ids = np.concatenate([my_ids[keep], got[:, 0].astype(np.int64)])
    local_rec = np.concatenate([my_rec[keep], got[:, 1:].astype(np.float32)])
    assert len(np.unique(local_ids)) == len(local_ids), "a particle arrived twice"
    local_rows = strips.cell_rows(local_rec[:, 1], HALF_H, CELL, GRID_Y)
    wlo, whi = strips.window(own, HALO, GRID_Y, WORLD)
    assert ((local_rows >= wlo) & (local_rows < whi)).all()
    # ghost set == every particle of the whole scene inside my window that I do not own
    all_rows = strips.cell_rows(p[:, 1], HALF_H, CELL, GRID_Y)
    expect_local = ids[(all_rows >= wlo) & (all_rows < whi)]
    assert np.array_equal(np.sort(local_ids), expect_local)
    own_now = strips.owned(local_rows, own)
    np.save(os.path.join(out_dir, f"ids{rank}.npy"), local_ids[own_now])
    np.save(os.path.join(out_dir, f"rec{rank}.npy"), local_rec[own_now])
    # the timing reduction of bench.py: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(WORLD)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_over_gloo(tmp_path, world):
    """two ranks: every strip has one neighbour; three ranks: the middle strip has two"""
    port = free_port()
    mp.spawn(worker, args=(port, str(tmp_path), world), nprocs=world, join=True)
    p, _ = load_state()
    parts = [(np.load(tmp_path / f"ids{r}.npy"), np.load(tmp_path / f"rec{r}.npy")) for r in range(world)]
    assert all(len(a) > 0 for a, _ in parts)
    merged = strips.merge_owned(parts, len(p))
    assert np.array_equal(merged, p)


def test_split_and_windows():
    s = strips.split_rows(1366, 3072, 8)
    assert s[0][0] == 0 and s[-1][1] == 3072 and all(a[1] == b[0] for a, b in zip(s, s[1:]))
    assert all(hi - lo >= strips.DEFAULT_HALO_ROWS for lo, hi in s)
    assert strips.window((171, 342), 7, 3072, 8) == (164, 349)
    assert strips.window((0, 171), 7, 3072, 8) == (0, 178)
    assert strips.window((0, 3072), 7, 3072, 1) == (0, 3072)
    with pytest.raises(ValueError):
        strips.merge_owned([(np.array([0, 1]), np.zeros((2, 12))), (np.array([1, 2]), np.zeros((2, 12)))], 4)


def test_cell_rows_matches_reference_dump():
    g = np.load(os.path.join(HERE, "golden", "scene1.npz"))
    # after put_particles the reference re-filed the grid from the final positions
    last = g[f"state{int(g['steps'][-1])}"]
    assert np.array_equal(strips.cell_rows(last[:, 1], HALF_H, CELL, GRID_Y), g["refiled_cell_of_particle"][:, 1])


# ---- a re-balancing step (sph_set_rebalance) over gloo ---------------------------------------------
RB_GRID_Y, RB_HALO, RB_N = 120, 3, 30000


def rebalance_state():
    """particles piled up near the floor (rows on the previous grid), each moving at most one row in this step"""
    rng = np.random.default_rng(21)
    before = np.minimum(rng.geometric(0.06, RB_N) - 1, RB_GRID_Y - 1)
    after = np.clip(before + rng.integers(-1, 2, RB_N), 0, RB_GRID_Y - 1)
    return before, after


def rebalance_worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    before, after = rebalance_state()
    ids = np.arange(RB_N)
    old = [0, 60, RB_GRID_Y]
    # what the device does: row counts of the rows I own at their global index + my first row, summed over the ranks
    words = torch.zeros(RB_GRID_Y + WORLD, dtype=torch.int64)
    mine = strips.owned(before, (old[rank], old[rank + 1]))
    words[:RB_GRID_Y] = torch.from_numpy(np.bincount(before[mine], minlength=RB_GRID_Y))
    words[RB_GRID_Y + rank] = old[rank]
    dist.all_reduce(words)
    assert words[:RB_GRID_Y].sum().item() == RB_N and words[RB_GRID_Y:].tolist() == old[:WORLD]
    new = strips.plan_bounds(words[:RB_GRID_Y].tolist(), words[RB_GRID_Y:].tolist() + [RB_GRID_Y], RB_HALO, max_shift=4)
    plans = [None] * WORLD
    dist.all_gather_object(plans, new)
    assert all(p == new for p in plans), plans  # every rank planned the same split
    assert new[1] == old[1] - 4  # the pile is near the floor: the boundary moves down as far as it may
    own = (new[rank], new[rank + 1])
    keep, down, up, lost = strips.classify(after[mine], rank, WORLD, own, RB_HALO, RB_GRID_Y)  # authority: old rows; keep / send: new rows
    assert not lost.any()
    my_ids = ids[mine]
    recv = []
    for peer, mask in ((rank - 1, down), (rank + 1, up)):
        if not 0 <= peer < WORLD:
            continue
        send_buf = torch.from_numpy(my_ids[mask].astype(np.int64))
        cnt, other = torch.tensor([len(send_buf)]), torch.zeros(1, dtype=torch.long)
        for r in [dist.isend(cnt, peer), dist.irecv(other, peer)]:
            r.wait()
        got = torch.zeros(int(other.item()), dtype=torch.int64)
        for r in [dist.isend(send_buf, peer), dist.irecv(got, peer)]:
            r.wait()
        recv.append(got.numpy())
    arrived = np.concatenate(recv) if recv else np.zeros(0, np.int64)
    wlo, whi = strips.window(own, RB_HALO, RB_GRID_Y, WORLD)
    arrived = arrived[(after[arrived] >= wlo) & (after[arrived] < whi)]  # unpack_kernel files only what is inside the window
    local = np.concatenate([my_ids[keep], arrived])
    assert len(np.unique(local)) == len(local)
    assert np.array_equal(np.sort(local), ids[(after >= wlo) & (after < whi)])  # the whole new window
    np.save(os.path.join(out_dir, f"rb_ids{rank}.npy"), local[strips.owned(after[local], own)])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_rebalancing_step(tmp_path):
    """Histogram all-reduce, identical plans, authority by the old rows and keep/send by the new ones, one exchange:
    afterwards ownership is a partition and every rank holds its whole new window."""
    port = free_port()
    mp.spawn(rebalance_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    owned = np.concatenate([np.load(tmp_path / f"rb_ids{r}.npy") for r in range(WORLD)])
    assert len(owned) == RB_N and len(np.unique(owned)) == RB_N
