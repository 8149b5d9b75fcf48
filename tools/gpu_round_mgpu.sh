#!/bin/bash
# One multi-GPU box session: N strips == 1 GPU bit for bit (tools/mgpu_check.py), then the 16M-particle block of
# BASELINE.json configs[3] on all GPUs.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_mgpu.sh <tag> 8'
tag=${1:-m}
export TAG=$tag
n=${2:-8}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/mgpu_check.py --nx 512 --steps 150 > $out/${tag}_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"
grep -v "^W\|^\[W\|NCCL\|warn" $out/${tag}_mgpu_check.log | tail -6
timeout 400 $TR --master-port 29512 bench.py --gpus $n --nx-total 4096 --steps 256 --warmup 32 > $out/${tag}_bench_n${n}_16m.json 2> $out/${tag}_bench_n${n}_16m.err; echo "bench 16M rc=$?"
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG", "m"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $out/${tag}_bench_n${n}_16m.err
