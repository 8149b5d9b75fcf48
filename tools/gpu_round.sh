#!/bin/bash
# One GPU-box session: parity tests, bench (default + the nine-launch sweep for comparison), ncu launch list and a full
# capture of the sweep kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "flow or sweep or colored_block" > $out/${tag}_pytest_flow.log 2>&1; echo "flow tests rc=$?" | tee -a $out/${tag}_pytest_flow.log
timeout 900 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "all gpu tests rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --sweep warp --no-cpu > $out/${tag}_bench_warp.json 2> $out/${tag}_bench_warp.err; echo "bench warp rc=$?"
cat $out/${tag}_bench.json | cut -c1-400
cat $out/${tag}_bench_warp.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python tools/profile_step.py --steps 4 --gravity -0.5219 > $out/${tag}_p1.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"color_sweep_flow|density_kernel|reorder_kernel" -s 4 -c 6 -o $out/${tag}_prof -f python tools/profile_step.py --steps 4 --gravity -0.5219 > $out/${tag}_p2.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -12
