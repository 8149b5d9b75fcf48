"""Where a frame of the end-to-end loop on strips goes: per iteration, the host time of Update / wait_render_owned /
render_owned (bench.py's e2e leg, same calls in the same order), on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/e2e_probe.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimulation_experiment_b200 import ParticleSimulation, bind_host_to_gpu, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
bind_host_to_gpu(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx = int(round(1024 * world ** 0.5 / 32.0)) * 32
dt = float(np.float32(1.0) / np.float32(60.0))
sim = scenes.block_scene(nx, device=local, rank=rank, world_size=world)
uid = [ParticleSimulation.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
sim.comm_init(uid[0])
sim.set_strip(*scenes.block_strips(sim, world)[rank])
scenes.fill_block(sim)
for _ in range(5):
    sim.Update(dt)
bufs = [sim.owned_buffers(records=False, render=True, pinned=True) for _ in range(2)]
sim.render_owned(bufs[1])
sim.wait_render_owned()
sim.Sync()
torch.cuda.synchronize()
dist.barrier()
rows = []
t_start = time.perf_counter()
for k in range(24):
    t0 = time.perf_counter()
    sim.Update(dt)
    t1 = time.perf_counter()
    sim.wait_render_owned()
    t2 = time.perf_counter()
    sim.render_owned(bufs[k % 2])
    t3 = time.perf_counter()
    rows.append((round((t1 - t0) * 1e3, 3), round((t2 - t1) * 1e3, 3), round((t3 - t2) * 1e3, 3)))
sim.wait_render_owned()
sim.Sync()
total = (time.perf_counter() - t_start) * 1e3
out = [None] * world
dist.all_gather_object(out, {"rank": rank, "ms_per_frame": round(total / 24, 3), "update_wait_render_ms": rows[4:16]})
if rank == 0:
    for o in out:
        print(o)
sim.close()
dist.destroy_process_group()
