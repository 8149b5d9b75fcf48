"""Multi-GPU parity check: N y-strips (one process per GPU, NCCL halo + migration exchange) against
the same scene on one GPU, particle by particle, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/mgpu_check.py --nx 256 --steps 40

Every rank builds its strip of the scene; rank 0 additionally runs the whole scene on its own GPU.
After K steps the owned particles of all ranks are gathered and compared by creation id.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimulation_experiment_b200 import SPH_SOLVER_COLORED_GS, SPH_SOLVER_GATHER, ParticleSimulation, _lib, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--spacing", type=float, default=0.1)
    ap.add_argument("--gravity", type=float, default=-10.0)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--solver", default="gs")
    ap.add_argument("--halo-rows", type=int, default=0)
    ap.add_argument("--rebalance", type=int, default=0, help="re-balance the strips every N steps (0 = static strips)")
    ap.add_argument("--max-shift", type=int, default=2)
    ap.add_argument("--ny", type=int, default=0, help="block height in particles (default: nx); tall blocks give many strips enough rows")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="strip exchange: peer-memory mailboxes (default) or NCCL messages")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = SPH_SOLVER_GATHER if a.solver == "gather" else SPH_SOLVER_COLORED_GS
    dt = float(np.float32(1.0) / np.float32(60.0))

    # every strip is given room for the whole scene twice over (what it holds + what arrives behind it in one step) and
    # for whole-scene exchange messages: under g = -10 the column falls into the bottom strips, and the default sizing
    # (3x the even share) overflows from 4 strips on
    ny = a.ny or a.nx
    n_total = a.nx * ny
    flags = _lib.SPH_FLAG_EXCHANGE_NCCL if a.transport == "nccl" else 0
    sim = scenes.block_scene(a.nx, ny=ny, spacing=a.spacing, gravity=(0.0, a.gravity), device=local, rank=rank, world_size=world, solver=solver,
                             halo_rows=a.halo_rows, capacity=2 * n_total + 1024, halo_capacity=n_total, flags=flags)
    uid = [ParticleSimulation.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.comm_init(uid[0])
    strips = scenes.block_strips(sim, world)
    sim.set_strip(*strips[rank])
    if a.rebalance:
        sim.set_rebalance(a.rebalance, a.max_shift)
    scenes.fill_block(sim)
    counts = [None] * world
    dist.all_gather_object(counts, sim.local_particle_count())
    if rank == 0:
        print(f"strips {strips} owned {counts} total {sum(counts)} of {n_total}", flush=True)
    assert sum(counts) == n_total, counts

    err = None
    try:
        for _ in range(a.steps):
            sim.Update(dt)
        mine = sim.read_owned(records=True)
        sim.GetStats()  # raises on overflow flags
        final_strip = sim.get_strip()
    except Exception as e:  # fail fast on every rank instead of hanging the others in a collective
        err = repr(e)
    flag = torch.tensor([0 if err is None else 1], device="cuda")
    dist.all_reduce(flag)
    if flag.item():
        print(f"rank {rank}: step loop failed: {err}", flush=True)
        os._exit(2)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine["ids"], mine["records"]))
    final = [None] * world
    dist.all_gather_object(final, final_strip)
    if rank == 0:
        print("strips after", a.steps, "steps:", final, flush=True)
    ok = True
    if rank == 0:
        ids = np.concatenate([g[0] for g in gathered])
        rec = np.concatenate([g[1] for g in gathered])
        print("owned after", a.steps, "steps:", [len(g[0]) for g in gathered], flush=True)
        assert len(ids) == n_total and len(np.unique(ids)) == n_total, "ownership is not a partition of the particles"
        multi = np.zeros((n_total, 12), np.float32)
        multi[ids] = rec
        one = scenes.fill_block(scenes.block_scene(a.nx, ny=ny, spacing=a.spacing, gravity=(0.0, a.gravity), device=local, solver=solver))
        for _ in range(a.steps):
            one.Update(dt)
        ref = one.particles()
        same = (multi.view(np.uint32) == ref.view(np.uint32)) | ((multi == 0) & (ref == 0))
        bad = np.argwhere(~same.all(1)).ravel()
        print(f"exchange transport: {a.transport}", flush=True)
        print(f"{world}-GPU vs 1-GPU after {a.steps} steps: {len(bad)} of {n_total} particles differ; max abs diff {np.abs(multi - ref).max():.3e}", flush=True)
        if len(bad):
            rows = ((ref[bad, 1] + one.scene['height'] / 2) / scenes.KERNEL_HEIGHT).astype(int)
            print("rows of the differing particles:", np.unique(rows)[:40], flush=True)
            ok = False
        one.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    sim.close()
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)
    if rank == 0:
        print("MGPU PARITY OK", flush=True)


if __name__ == "__main__":
    main()
