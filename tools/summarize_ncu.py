"""Turns ncu outputs brought back from the GPU box into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  profiles/rN_launches_<what>.txt  "<command that was profiled>"
    python tools/summarize_ncu.py kernels  gpurun_out/prof.ncu-rep   profiles/rN_kernels_<what>.txt   "<command that was profiled>"

`traffic`: per-kernel DRAM bytes and duration per launch of a `--set full` report as JSON (profiles/rN_traffic*.json, read by
bench.py for roofline.traffic).  `launches`: the CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`; prints the
share of the step every kernel takes.  `kernels`: a `--set full` report; prints, per captured launch,
duration, DRAM bytes, issue-slot utilisation, pipe utilisation, lane efficiency, occupancy limits and the
top warp-stall reasons.
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, what):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {what}\n# ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times are cold-cache and serialised, read the SHARES\n")
        f.write(f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:58s} launches={c:4d} total_us={t:10.1f} avg_us={t / c:8.1f} share={t / tot * 100:5.1f}%\n")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block", "shared memory/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks"),
]


def kernels(src, dst, what):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    with open(dst, "w") as f:
        f.write(f"# {what}\n# ncu --set full --clock-control none --import-source on; values are per launch\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')][:110]}\n")
            for key, label in WANT:
                if key in hdr:
                    f.write(f"   {label:42s} {r[hdr.index(key)]} {units[hdr.index(key)]}\n")
            samples = []
            for h in stall:
                try:
                    samples.append((float(r[hdr.index(h)].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
            tot = sum(s for s, _ in samples) or 1.0
            top = ", ".join(f"{n} {s / tot * 100:.0f}%" for s, n in sorted(samples, reverse=True)[:6])
            f.write(f"   {'warp stall samples (top)':42s} {top}\n")


def traffic(src, dst, what):
    """dram__bytes_read.sum + dram__bytes_write.sum and duration per launch of every captured kernel -> JSON (bench.py's roofline.traffic)"""
    import json

    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("sphb200::", "").replace("(int)", "")
        b = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            b += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        i = hdr.index("gpu__time_duration.sum")
        t = float(r[i].replace(",", "")) * tscale.get(units[i], 1.0)
        inst = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", "")) if "smsp__inst_executed.sum" in hdr else 0.0
        busy = float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")].replace(",", "")) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else 0.0
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += b
        a[2] += t
        a[3] += inst
        a[4] += busy
    out = {k: {"launches_captured": c, "dram_bytes_per_launch": b / c, "duration_us_per_launch": t / c, "warp_instructions_per_launch": i / c,
               "issue_slots_busy_pct": u / c} for k, (c, b, t, i, u) in agg.items()}
    out["_source"] = what
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    what = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "kernels": kernels, "traffic": traffic}[mode](src, dst, what)
    print(open(dst).read())
