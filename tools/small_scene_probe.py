"""ms/step of the reference's scenes on cuda:0 (small, latency-bound workloads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nbodysimulation_experiment_b200 import ParticleSimulation
dt = float(np.float32(1) / np.float32(60))
for scene in (0, 1, 2):
    sim = ParticleSimulation(); sim.LoadScenario(scene, seed=1)
    for _ in range(20): sim.Update(dt)
    sim.Sync(); sim.mark(0)
    for _ in range(200): sim.Update(dt)
    sim.mark(1); ms = sim.elapsed_ms(0, 1) / 200
    n = sim.GetParticleCount()
    print(f"scene {scene}: {n} particles, {ms:.4f} ms/step, {n / ms * 1e3:.3e} particle-steps/s")
    sim.close()
