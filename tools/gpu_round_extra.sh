#!/bin/bash
# The other workloads of the results table on one GPU: fast fp mode, the dense 1M block, the
# 4M-particle scene with bodies (BASELINE.json configs[4]), the reference's own scene; plus the GPU tests.
#   gpurun --timeout 900 -- 'bash tools/gpu_round_extra.sh <tag>'
tag=${1:-x}
export TAG=$tag
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "all gpu tests rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --fp fast --no-cpu > $out/${tag}_bench_fast.json 2> $out/${tag}_bench_fast.err; echo "bench fast rc=$?"
timeout 200 python bench.py --workload dambreak_1m_dense --no-cpu --steps 128 > $out/${tag}_bench_dense.json 2> $out/${tag}_bench_dense.err; echo "bench dense rc=$?"
timeout 200 python bench.py --workload dambreak_1m_dense --no-cpu --steps 128 --sweep warp > $out/${tag}_bench_dense_warp.json 2> $out/${tag}_bench_dense_warp.err; echo "bench dense warp rc=$?"
timeout 300 python bench.py --workload bodies_4m --no-cpu --steps 128 --warmup 32 > $out/${tag}_bench_bodies4m.json 2> $out/${tag}_bench_bodies4m.err; echo "bench bodies rc=$?"
timeout 100 python tools/small_scene_probe.py > $out/${tag}_small_scene.log 2>&1; echo "small scene rc=$?"; tail -3 $out/${tag}_small_scene.log
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ["TAG"])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "cand %.0f" % d["config"]["candidates_per_particle_rank0"],
              {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
