"""Per-step device time of the first K steps of a scene, on every rank (every
Update must cost the same from step 1 on - the reference's contract, app.cpp:228-236).

    python tools/step_timeline.py --steps 40                                            # one GPU, the 1M block
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/step_timeline.py --steps 40 --out profiles/r2_timeline_n8.json

Each step is bracketed by two CUDA events on the simulation's stream and read after the step has finished, so the
host cannot run ahead: a step's time includes whatever its launches wait for (graph capture and instantiation, NCCL
connection set-up, a neighbour strip that is late).  Rank 0 prints / writes {rank: [ms per step]}.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimulation_experiment_b200 import ParticleSimulation, _lib, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1024, help="block edge per GPU (weak scaling like bench.py)")
    ap.add_argument("--nx-total", type=int, default=0)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--gravity", type=float, default=-10.0)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--out", default="")
    ap.add_argument("--no-sync", action="store_true", help="enqueue all steps without waiting (as bench.py does); per-step times from events recorded between the steps")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch

    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = a.nx_total or (a.nx if world == 1 else int(round(a.nx * world ** 0.5 / 32.0)) * 32)
    dt = float(np.float32(1.0) / np.float32(60.0))
    flags = _lib.SPH_FLAG_EXCHANGE_NCCL if a.transport == "nccl" else 0
    t0 = time.perf_counter()
    sim = scenes.block_scene(nx, gravity=(0.0, a.gravity), device=local, rank=rank, world_size=world, flags=flags)
    if world > 1:
        uid = [ParticleSimulation.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(uid[0])
        sim.set_strip(*scenes.block_strips(sim, world)[rank])
    scenes.fill_block(sim)
    sim.Sync()
    setup_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    ms, host_ms = [], []
    if a.no_sync:
        stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        evs[0].record(stream)
        for k in range(a.steps):
            h0 = time.perf_counter()
            sim.Update(dt)
            evs[k + 1].record(stream)
            host_ms.append((time.perf_counter() - h0) * 1e3)
        sim.Sync()
        ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(a.steps)]
    else:
        for _ in range(a.steps):
            h0 = time.perf_counter()
            sim.mark(0)
            sim.Update(dt)
            sim.mark(1)
            ms.append(sim.elapsed_ms(0, 1))  # waits for the step
            host_ms.append((time.perf_counter() - h0) * 1e3)
    sim.GetStats()
    sim.close()
    mine = {"rank": rank, "device_ms": [round(x, 4) for x in ms], "host_ms": [round(x, 4) for x in host_ms], "setup_s": round(setup_s, 3)}
    rows = [mine]
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, mine)
        dist.destroy_process_group()
    if rank == 0:
        doc = {"what": f"per-step time of the first {a.steps} steps, {nx}x{nx} = {nx * nx} particles on {world} GPU(s), g = {a.gravity}, exchange {a.transport if world > 1 else '-'}",
               "ranks": rows}
        worst = np.max(np.array([r["device_ms"] for r in rows]), axis=0)
        doc["max_over_ranks_ms"] = [round(float(x), 4) for x in worst]
        doc["steady_ms"] = round(float(np.median(worst[len(worst) // 2:])), 4)
        doc["first_steps_over_steady"] = [round(float(x / doc["steady_ms"]), 2) for x in worst[:8]]
        text = json.dumps(doc)
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            open(a.out, "w").write(text + "\n")
        print(text)


if __name__ == "__main__":
    main()
