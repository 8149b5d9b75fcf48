"""Device-to-host bandwidth of plain pinned copies with all ranks copying at once (what bounds the end-to-end leg of
bench.py on strips: 28-31 MB per rank and frame).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mb = 30
src = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, device="cuda")
dst = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
gbs = mb * 1024 * 1024 * reps / (time.perf_counter() - t0) / 1e9
out = [gbs]
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, gbs)
    dist.destroy_process_group()
if rank == 0:
    print(f"{world} rank(s) copying {mb} MB device->pinned host at once: per rank {[round(x, 1) for x in out]} GB/s, aggregate {sum(out):.1f} GB/s")
