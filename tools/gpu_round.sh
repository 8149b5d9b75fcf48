#!/bin/bash
# One GPU-box session: parity tests, bench (default + the nine-launch sweep for comparison), ncu launch list and a full
# capture of the sweep kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-x}
export TAG=$tag
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "flow or sweep or colored_block" > $out/${tag}_pytest_flow.log 2>&1; echo "flow tests rc=$?" | tee -a $out/${tag}_pytest_flow.log
timeout 900 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "all gpu tests rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --sweep warp --no-cpu > $out/${tag}_bench_warp.json 2> $out/${tag}_bench_warp.err; echo "bench warp rc=$?"
timeout 200 python bench.py --nx 512 --no-cpu > $out/${tag}_bench512.json 2> $out/${tag}_bench512.err; echo "bench512 rc=$?"
timeout 200 python bench.py --nx 512 --sweep warp --no-cpu > $out/${tag}_bench512_warp.json 2> $out/${tag}_bench512_warp.err; echo "bench512 warp rc=$?"
python - <<'PY'
import json, glob, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("TAG", "")
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ["TAG"])):
    try:
        d = json.loads(open(f).read())
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python tools/profile_step.py --steps 4 --gravity -0.5219 > $out/${tag}_p1.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"color_sweep_flow|density_kernel|reorder_kernel" -s 4 -c 6 -o $out/${tag}_prof -f python tools/profile_step.py --steps 4 --gravity -0.5219 > $out/${tag}_p2.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -12
