#!/bin/bash
# The driver's scaling run in small: bench.py at N = 8, 4, 2 exactly as the driver launches it, on one 8-GPU box.
tag=${1:-sc}
export TAG=$tag
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 "$@"; }
timeout 400 bash -c "$(declare -f run); run 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --per-step" > $out/${tag}_bench_n8_a.json 2> $out/${tag}_bench_n8_a.err; echo "bench 8 a rc=$?"
timeout 400 bash -c "$(declare -f run); run 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5" > $out/${tag}_bench_n8_b.json 2> $out/${tag}_bench_n8_b.err; echo "bench 8 b rc=$?"
timeout 300 bash -c "$(declare -f run); run 4 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 --no-c4" > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err; echo "bench 4 rc=$?"
timeout 300 bash -c "$(declare -f run); run 2 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-c4" > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; echo "bench 2 rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench 1 rc=$?"
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG", "m"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
        if d.get("c4_value"):
            print("   c4 %.4g ms/step %.4f e2e %.4g" % (d["c4_value"], d["c4_ms_per_step"], d["c4_e2e"]["value"]))
        if d.get("per_step_ms_rank0"):
            print("   per step", d["per_step_ms_rank0"])
    except Exception as e:
        print(f, "unreadable", e)
PY
