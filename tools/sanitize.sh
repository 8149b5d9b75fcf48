#!/bin/bash
# compute-sanitizer over the hot path: memcheck and synccheck on the reference's scene 0 (block-per-cell
# sweeps), on a 65k block through the one-launch sweep (inter-block acquire / release flags, heavy-cell teams), and on
# three strips sharing the GPU (mailbox protocol).  Logs -> gpurun_out/<tag>_sanitize_*.log
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh <tag>'
tag=${1:-san}
out=gpurun_out
mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from nbodysimulation_experiment_b200 import ParticleSimulation, StripGroup, _lib, scenes
dt = float(np.float32(1.0) / np.float32(60.0))
case = sys.argv[1]
if case == "scene0":
    s = ParticleSimulation(); s.LoadScenario(0, seed=1)
    for _ in range(6): s.Update(dt)
    s.GetStats(); print("scene0 ok", s.GetParticleCount()); s.close()
elif case == "flow":
    s = scenes.fill_block(scenes.block_scene(256, gravity=(0.0, -10.0), flags=_lib.SPH_FLAG_SWEEP_FLOW, sweep_capacity=96))  # cap 96: part of the cells go to the teams
    for _ in range(6): s.Update(dt)
    s.GetStats(); print("flow ok", s.GetParticleCount()); s.close()
elif case == "strips":
    sims = [scenes.block_scene(96, ny=192, rank=r, world_size=3, capacity=40000, halo_capacity=20000) for r in range(3)]
    rows = scenes.block_strips(sims[0], 3)
    for r, s in enumerate(sims): s.set_strip(*rows[r])
    g = StripGroup(sims)
    for s in sims: scenes.fill_block(s)
    for _ in range(6): g.Update(dt)
    for s in sims: s.GetStats()
    print("strips ok", [s.local_particle_count() for s in sims]); g.close()
PY
for tool in memcheck synccheck; do
  for case in scene0 flow strips; do
    timeout 400 $CS --tool $tool --print-limit 20 python /tmp/san_case.py $case > $out/${tag}_sanitize_${tool}_${case}.log 2>&1
    echo "$tool $case rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $out/${tag}_sanitize_${tool}_${case}.log) $(grep 'ERROR SUMMARY' $out/${tag}_sanitize_${tool}_${case}.log | head -1)"
  done
done
