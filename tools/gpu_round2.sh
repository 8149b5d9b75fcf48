#!/bin/bash
# Round-2 one-GPU box session: all GPU parity tests, the bench line with the driver's flags and the default ones, the
# per-step timeline, the ncu launch list and a full capture of the pair kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh <tag> [quick]'
tag=${1:-r2}
quick=${2:-}
export TAG=$tag
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
timeout 1100 python -m pytest tests -x -q -m gpu --durations=8 > $out/${tag}_pytest.log 2>&1; echo "all gpu tests rc=$?" | tee -a $out/${tag}_pytest.log
tail -14 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_20.json 2> $out/${tag}_bench_20.err; echo "bench 20/5 rc=$?"
timeout 300 python bench.py --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu --workload dambreak_1m_scaled --steps 256 --warmup 32 > $out/${tag}_bench_scaled.json 2> $out/${tag}_bench_scaled.err; echo "bench scaled rc=$?"
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ["TAG"])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
if [ -z "$quick" ]; then
timeout 200 python tools/step_timeline.py --steps 40 --out $out/${tag}_timeline_n1.json > /dev/null 2> $out/${tag}_timeline.err; echo "timeline rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python tools/profile_step.py --steps 4 > $out/${tag}_p1.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"color_sweep_flow|density|reorder_kernel|predict_key|scan_" -s 14 -c 12 -o $out/${tag}_prof -f python tools/profile_step.py --steps 4 > $out/${tag}_p2.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $out | tail -14
