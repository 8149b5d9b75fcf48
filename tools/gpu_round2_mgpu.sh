#!/bin/bash
# Round-2 multi-GPU box session: N strips == 1 GPU bit for bit over both transports (tools/mgpu_check.py), the per-step
# timeline of the first 40 steps, and bench.py exactly as the driver launches it (weak-scaling leg + the 16M c4 leg).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round2_mgpu.sh <tag> 8'
tag=${1:-m}
export TAG=$tag
n=${2:-8}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 300 $TR --master-port 29511 tools/mgpu_check.py --nx 256 --ny 1024 --steps 100 > $out/${tag}_mgpu_check_peer.log 2>&1; echo "mgpu_check peer rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn" $out/${tag}_mgpu_check_peer.log | tail -6
timeout 300 $TR --master-port 29512 tools/mgpu_check.py --nx 256 --ny 1024 --steps 100 --transport nccl > $out/${tag}_mgpu_check_nccl.log 2>&1; echo "mgpu_check nccl rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn" $out/${tag}_mgpu_check_nccl.log | tail -4
timeout 300 $TR --master-port 29513 tools/mgpu_check.py --nx 256 --ny 1024 --steps 100 --rebalance 8 --max-shift 3 > $out/${tag}_mgpu_check_rebalance.log 2>&1; echo "mgpu_check rebalance rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn" $out/${tag}_mgpu_check_rebalance.log | tail -4
timeout 200 $TR --master-port 29514 tools/step_timeline.py --steps 40 --out $out/${tag}_timeline_n${n}.json > /dev/null 2> $out/${tag}_timeline.err; echo "timeline rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_timeline_n${n}.json')); print('timeline max over ranks', d['max_over_ranks_ms'][:12], 'steady', d['steady_ms'])"
timeout 600 $TR --master-port 29515 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err; echo "bench rc=$?"
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
tail -3 $out/${tag}_bench_n${n}.err
