#!/bin/bash
# Round-2 eight-GPU box session (lean: 8 GPUs are charged 8x): parity at 8 and 4 ranks, the per-step timeline of the first
# 40 steps at 8 ranks, bench.py exactly as the driver launches it at N=8 (weak-scaling leg + the 16M c4 leg).
tag=${1:-m8}
export TAG=$tag
out=gpurun_out
mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
timeout 200 bash -c "$(declare -f run); run 8 --master-port 29511 tools/mgpu_check.py --nx 512 --steps 100" > $out/${tag}_mgpu_check_8.log 2>&1; echo "mgpu_check 8 rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn\|OMP\|\*\*\*" $out/${tag}_mgpu_check_8.log | tail -5
timeout 200 bash -c "$(declare -f run); run 4 --master-port 29512 tools/mgpu_check.py --nx 512 --steps 100 --rebalance 8 --max-shift 3" > $out/${tag}_mgpu_check_4_rebalance.log 2>&1; echo "mgpu_check 4 rebalance rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn\|OMP\|\*\*\*" $out/${tag}_mgpu_check_4_rebalance.log | tail -5
timeout 200 bash -c "$(declare -f run); run 8 --master-port 29514 tools/step_timeline.py --steps 40 --out $out/${tag}_timeline_n8.json" > /dev/null 2> $out/${tag}_timeline.err; echo "timeline rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_timeline_n8.json')); print('timeline max over ranks', d['max_over_ranks_ms'][:12], 'steady', d['steady_ms'])"
timeout 400 bash -c "$(declare -f run); run 8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5" > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; echo "bench 8 rc=$?"
timeout 300 bash -c "$(declare -f run); run 4 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 --no-c4" > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err; echo "bench 4 rc=$?"
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG", "m"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
        if d.get("c4_value"):
            print("   c4 %.4g ms/step %.4f e2e %.4g" % (d["c4_value"], d["c4_ms_per_step"], d["c4_e2e"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $out/${tag}_bench_n8.err
