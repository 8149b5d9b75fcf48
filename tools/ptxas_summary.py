"""ptxas resource usage and a SASS opcode histogram per kernel of libsphb200.so, for profiles/ (run in the authoring
container: nvcc cross-compiles, no GPU needed).

    python tools/ptxas_summary.py > profiles/r2_ptxas_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysimulation_experiment_b200 import build  # noqa: E402


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "")).replace("void ", "").replace("sphb200::", "") for o in out]


def main():
    cmd = [build._nvcc()] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-o", "/tmp/_ptxas_probe.so"] + build.SOURCES
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    rows, name = [], None
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            spill = ""
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"stack {m.group(1)} spill st/ld {m.group(2)}/{m.group(3)}"
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            rows.append((name, int(m.group(1)), m.group(3) or "0", spill))
            name = None
    names = demangle([r[0] for r in rows])
    print("# ptxas -v (sm_100a), nvcc", " ".join(build.NVCC_FLAGS[:6]))
    print(f"{'kernel':58s} {'regs':>5s} {'smem':>6s}  spills")
    for (raw, regs, smem, spill), nm in sorted(zip(rows, names), key=lambda x: x[1]):
        print(f"{nm[:58]:58s} {regs:5d} {smem:>6s}  {spill}")
    # SASS opcode histogram of the hot kernels
    sass = subprocess.run(["cuobjdump", "-sass", "/tmp/_ptxas_probe.so"], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    keys = list(hist)
    for raw, nm in zip(keys, demangle(keys)):
        if not re.search(r"color_sweep_flow|density_kernel|reorder_kernel|predict_key|grid_scan|collide_velocity|integrate", nm):
            continue
        h = hist[raw]
        tot = sum(h.values())
        print(f"\n== {nm}: {tot} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
    os.remove("/tmp/_ptxas_probe.so")


if __name__ == "__main__":
    main()
