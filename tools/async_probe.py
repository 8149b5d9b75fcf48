"""Experiment: does enqueueing many steps without a host sync change the device time per step?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if "--torch" in sys.argv:
    import torch
    torch.cuda.set_device(0)
    if "--flush" in sys.argv:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda"); flush.fill_(1); torch.cuda.synchronize()
from nbodysimulation_experiment_b200 import scenes
flags = 2 if "--nographs" in sys.argv else 0
sim = scenes.fill_block(scenes.block_scene(1024, spacing=0.1, gravity=(0.0, -0.5219), flags=flags))
dt = float(np.float32(1) / np.float32(60))
for _ in range(32):
    sim.Update(dt)
sim.Sync()
t0 = time.perf_counter()
for k in range(6):
    sim.mark(k)
    for _ in range(50):
        sim.Update(dt)
    if "--sync" in sys.argv:
        sim.Sync()
sim.mark(6)
host_enqueue = time.perf_counter() - t0
sim.Sync()
print(sys.argv[1:], "host enqueue s %.3f" % host_enqueue, "ms/step per block:", ["%.3f" % (sim.elapsed_ms(k, k + 1) / 50) for k in range(6)], "total %.3f" % (sim.elapsed_ms(0, 6) / 300))
