#!/usr/bin/env python
"""bench.py — particle-steps/s of the SPH hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, via the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on host cores

A "step" is one Update(1/60) of the whole scene (demo4.cpp:286-451).  Workload at N=1 is BASELINE.json
configs[2]: the 1M-particle dam-break block under the reference's gravity (0,-10) (see WORKLOADS below and
DESIGN.md); with N>1 every rank owns a y-strip of a scene N times larger (weak scaling), and a second leg in the
same process times BASELINE.json configs[3], the 16M-particle block, on the same N GPUs (`c4` in the line).  One
JSON line is printed by rank 0.

Timing: CUDA events on the simulation's own stream (sph_mark / sph_elapsed_ms), W >= 3 warm-up steps, a
barrier + device synchronize on both sides, max over ranks.  The state (~90 B/particle + 8 B/cell, >130 MB at
1M particles... plus the cell arrays) is streamed once per phase, and the L2 is additionally flushed between
the warm-up and the timed region; `config.l2` says so.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = float(np.float32(1.0) / np.float32(60.0))
METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"

# BASELINE.json configs[2] / SURVEY.md 8(d) c3.  nx*nx particles per GPU.
WORKLOADS = {
    # SURVEY.md 8(d) c3 as written: gravity (0,-10).  Default run length 8 + 64 steps: under the reference's g the
    # 102-unit column collapses at ~45 units/s = 2.5 cells per step after ~4 s and the fixed-dt relaxation (the
    # reference's own semantics as much as ours: profiles/r2_oracle_column_g10.log) breaks down around step 250.
    "dambreak_1m": dict(nx=1024, spacing=0.1, gravity_scale=False, relaxation=1.0),
    # the long-run variant: gravity scaled to the hydrostatic head of the reference's own dam (5.34 units, sph.h:310),
    # stable for thousands of steps; use with --steps 256 --warmup 32
    "dambreak_1m_scaled": dict(nx=1024, spacing=0.1, gravity_scale=True, relaxation=1.0),
    "dambreak_1m_dense": dict(nx=1024, spacing=0.05, gravity_scale=True, relaxation=1.0),
    # BASELINE.json configs[4] / SURVEY.md 8(d) c5: 2048 x 2048 dense block, 10x viscosity, circles + boxes.
    # Use --steps 128 --warmup 32: at 10x viscosity the explicit impulses (sph.h:508) diverge after ~250 steps
    "bodies_4m": dict(nx=2048, spacing=0.05, gravity_scale=True, relaxation=1.0, bodies=True),
}
BYTES_PER_PARTICLE_STEP = 176  # SURVEY.md 8(d) / BASELINE.md 4
BYTES_PER_CELL_STEP = 16
# algorithmic HBM bytes per particle of each phase (DESIGN.md "kernels")
PHASE_BYTES = {"integrate": 16, "viscosity": 24, "predict_key": 36, "scan": 0, "reorder": 44, "density": 16, "delta": 24, "collide_velocity": 32}
# integrate, viscosity sweep (1 launch, or 9 with --sweep warp/team), predict_key, scan (tile sums, apply), colour lists (count, fill),
# scatter_ids, reorder, density, delta sweep (1 or 9), collide_velocity
KERNELS_PER_STEP = {"gs": 12, "gs9": 28, "gather": 10}
TRAFFIC_FILE = "r2_traffic.json"  # ncu --set full figures of the dominant kernel on the default workload (tools/summarize_ncu.py traffic)


def scene_gravity(nx, spacing, scaled):
    """Dam break under the reference's gravity (0,-10) scaled so that the hydrostatic head matches the
    reference scene (5.34 units of fluid, sph.h:310): fixed dt = 1/60 and h = 0.3 are only stable when
    |v| dt stays below h ("scene scaling")."""
    height = nx * spacing
    return (0.0, -10.0 * min(1.0, 5.34375 / height)) if scaled else (0.0, -10.0)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (pynvml) polling
    every 20 ms: spawning `nvidia-smi -lms` next to a 0.3 s timed region costs more than it measures (its
    start-up and driver locks stalled the launch stream by ~40 % here).  Falls back to nvidia-smi."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, device):
        self.device = device
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.armed = False  # samples count only while armed (the thread is started ahead of the timed region)
        self.thread = None
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = device
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[device])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_now(self):
        """one sample from the calling thread: bench.py takes it right after the last timed step is enqueued, when the
        host has nothing to do and the GPU is busy with the queued steps - a timed region shorter than the polling period
        still gets a sample under load"""
        if not self.nvml:
            return
        nv = self.nvml
        try:
            clock = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.sm.append(clock)
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag.is_set():
            if self.armed:
                self.sample_now()
            self.stop_flag.wait(0.02)

    def start(self):
        if self.nvml:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "how": "NVML polled every 20 ms inside the timed region, plus once after the last step was enqueued"}
        try:  # one-shot fallback right after the region
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                                  "-i", str(self.device)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [out[2].strip()], "samples": 1, "how": "nvidia-smi once, after the region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"], "samples": 0}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from nbodysimulation_experiment_b200 import (SPH_FLAG_PHASE_TIMING, SPH_FP_EXACT, SPH_FP_FAST, SPH_SOLVER_COLORED_GS, SPH_SOLVER_GATHER,
                                                 ParticleSimulation, bind_host_to_gpu, pinned_empty, scenes)
    from nbodysimulation_experiment_b200 import _lib

    sweep_flags = {"auto": 0, "flow": _lib.SPH_FLAG_SWEEP_FLOW, "warp": _lib.SPH_FLAG_SWEEP_WARP, "team": _lib.SPH_FLAG_SWEEP_TEAM}[args.sweep]
    if args.transport == "nccl":
        sweep_flags |= _lib.SPH_FLAG_EXCHANGE_NCCL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    # host threads and the page-locked readback frames next to this rank's GPU (8 ranks share two sockets)
    numa_bound = bind_host_to_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = WORKLOADS[args.workload]
    spacing = wl["spacing"]
    fp_mode = SPH_FP_FAST if args.fp == "fast" else SPH_FP_EXACT
    solver = SPH_SOLVER_GATHER if args.solver == "gather" else SPH_SOLVER_COLORED_GS
    relaxation = args.relaxation if args.relaxation else wl["relaxation"]
    one_launch = args.sweep == "flow" or (args.sweep == "auto")  # every bench scene has >= 131072 particles per GPU
    kernels_per_step = KERNELS_PER_STEP["gather" if args.solver == "gather" else ("gs" if (one_launch or world > 1) else "gs9")]
    if world > 1:  # publish, wait, one unpack per neighbour (rank 0 has one); no NCCL kernel inside a step with the peer transport
        kernels_per_step += 3
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def all_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def leg(nx, steps, warmup, with_phases):
        """One workload on all ranks: device-timed steps, then the end-to-end frames, then (optionally) the per-phase pass."""
        gravity = scene_gravity(nx, spacing, wl["gravity_scale"] or args.scaled_gravity)

        def make(flags=0):
            build = scenes.bodies_scene if wl.get("bodies") else scenes.block_scene
            sim = build(nx, spacing=spacing, gravity=gravity, fp_mode=fp_mode, flags=flags | sweep_flags, relaxation=relaxation, device=local_rank,
                        solver=solver, sweep_capacity=args.sweep_capacity, rank=rank, world_size=world, halo_rows=args.halo_rows)
            if world > 1:
                uid = [ParticleSimulation.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                sim.comm_init(uid[0])
                sim.set_strip(*scenes.block_strips(sim, world)[rank])
                if args.rebalance:
                    sim.set_rebalance(args.rebalance)
            return scenes.fill_block(sim)

        sim = make()
        n_total = nx * nx
        n_local = sim.local_particle_count()
        gx, gy = sim.grid_dims()

        def barrier():
            sim.Sync()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()

        for _ in range(warmup):
            sim.Update(DT)
        flush.fill_(1)
        # Everything slow goes BEFORE the barrier: on strips a rank that enters the timed loop late keeps its neighbours
        # waiting inside their first step (they need its halo records), and the wait then travels up the strips one rank
        # per step.  (Measured, profiles/bench_r2/r2w_*: NVML initialisation between the barrier and the first step
        # on every rank cost 2-3.5 ms in the first step of half the ranks, 0.70 -> 0.84 ms/step over 20 steps.)
        # Clocks are sampled on rank 0 only (its numbers are the ones reported): eight processes polling NVML inside
        # a 15 ms timed region contend for the driver and cost every rank ~6 %.
        sample_clocks = not args.no_clock_sampler and rank == 0
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sample_clocks:
            sampler.start()  # polls from now on, keeps samples only while armed
        per_step = None
        if args.per_step:  # diagnostic: an event between the steps (recorded on the simulation's stream, read after the region)
            stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local_rank))
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        host_ms = []
        barrier()
        if sample_clocks:
            sampler.armed = True
        sim.mark(0)
        for k in range(steps):
            if args.per_step:
                evs[k].record(stream)
                h0 = time.perf_counter()
            sim.Update(DT)
            if args.per_step:
                host_ms.append(round((time.perf_counter() - h0) * 1e3, 3))
        if args.per_step:
            evs[steps].record(stream)
        sim.mark(1)
        if sample_clocks:
            sampler.sample_now()  # the device is still working through the queued steps
        ms = sim.elapsed_ms(0, 1)
        if args.per_step:
            mine = {"rank": rank, "device_ms": [round(evs[k].elapsed_time(evs[k + 1]), 3) for k in range(steps)], "host_ms": host_ms}
            per_step = [mine]
            if world > 1:
                per_step = [None] * world
                dist.all_gather_object(per_step, mine)
        if sample_clocks:
            sampler.armed = False
        barrier()
        clocks = sampler.stop() if sample_clocks else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler disabled" if rank == 0 else "sampled on rank 0"]}
        ms = all_max(ms)
        stats = sim.GetStats()  # raises if a capacity / lost / timeout flag was set on the device
        out = {"nx": nx, "n_total": n_total, "n_local": n_local, "cells": gx * gy, "grid": (gx, gy), "gravity": gravity, "ms": ms, "steps": steps,
               "value": n_total * steps / (ms * 1e-3), "clocks": clocks, "candidates": stats.pair_candidates / max(n_local, 1), "per_step_ms": per_step}

        # ---- e2e: Update + Render readback (positions + colours) into host memory every step, the per-frame
        # traffic of the reference's app loop (app.cpp:231-233,286-289).  One GPU: creation-order arrays in pinned
        # memory.  Strips: each rank reads back the particles it owns (compacted, with their ids).
        # The SAME steps of the scene as the device-timed leg (a fresh simulation, the same warm-up): the step cost
        # of a collapsing column grows with the step number, and the two numbers are meant to differ by the readback
        # alone. ------------------------------------------------------------------------------------------------
        sim.close()
        sim = make()
        for _ in range(warmup):
            sim.Update(DT)
        e2e_steps = steps
        params_blob = np.zeros(16, np.float32)  # per-step host inputs: dt, gravity, external force
        d2h = 0
        if world == 1:
            frames = [(pinned_empty((n_total, 2), np.float32), pinned_empty((n_total, 4), np.float32)) for _ in range(2)]
            sim.Render(frames[1][0][0], frames[1][1][0], wait=False)  # untimed warm-up frame: the first Render allocates the device-side snapshot
            sim.WaitRender()
        else:  # (and on strips it learns how many particles a frame ships)
            owned_bufs = [sim.owned_buffers(records=False, render=True, pinned=True) for _ in range(2)]
            sim.render_owned(owned_bufs[1])
            sim.wait_render_owned()
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            sim.SetGravity(gravity)
            sim.Update(DT)  # enqueued; runs while the previous frame's copy drains
            if world == 1:
                sim.WaitRender()  # frame k-1 is now complete in host memory
                (pos_host, _), (col_host, _) = frames[k % 2]
                sim.Render(pos_host, col_host, wait=False)  # snapshot + D2H on the copy stream
                d2h = n_total * 24
            else:
                got = sim.wait_render_owned()  # frame k-1 (None before the first one)
                sim.render_owned(owned_bufs[k % 2])  # snapshot + D2H on the copy stream, overlapped with Update k+1
                if got is not None:
                    d2h = len(got["ids"]) * 28
        if world == 1:
            sim.WaitRender()
        else:
            got = sim.wait_render_owned()
            d2h = len(got["ids"]) * 28
            assert np.isfinite(got["positions"]).all() and (got["colors"][:, 3] == 1.0).all()
        barrier()
        e2e_s = all_max(time.perf_counter() - t0)
        out["e2e"] = {"value": n_total * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(params_blob.nbytes), "d2h_bytes_per_step": int(d2h),
                      "steps": e2e_steps, "what": "Update + Render readback (pos float2 + colour float4" + (" + id" if world > 1 else "") + ") to host memory every step"
                      + "; double-buffered pinned frames, the copy of frame k overlaps Update k+1" + ("; host threads and frames bound to the GPU's NUMA node" if numa_bound else "")}
        if world == 1:
            for (p_arr, p_own), (c_arr, c_own) in frames:
                assert np.isfinite(p_arr).all() and (c_arr[:, 3] == 1.0).all()
                p_own.free()
                c_own.free()
        else:
            del owned_bufs
        sim.close()

        # ---- per-phase device times (separate pass: the event brackets serialise host and device) --------
        if with_phases:
            psim = make(flags=SPH_FLAG_PHASE_TIMING)
            for _ in range(3):
                psim.Update(DT)
            psim.ResetStats()
            for _ in range(max(3, min(steps, 20))):
                psim.Update(DT)
            out["phases"], _ = psim.phase_ms()
            out["n_phase_local"] = psim.local_particle_count()
            psim.close()
        return out

    # weak scaling: the block grows with the GPU count so that every strip holds ~nx*nx particles
    nx_one = args.nx or wl["nx"]
    nx = nx_one if world == 1 else int(round(nx_one * world ** 0.5 / 32.0)) * 32
    if args.nx_total:  # builder flag: edge of the whole block, overriding the weak-scaling rule
        nx = args.nx_total
    warmup = max(args.warmup, 3)
    main = leg(nx, args.steps, warmup, True)
    # BASELINE.json configs[3] / SURVEY.md 8(d) c4: the 16M-particle block on the same N GPUs, same flags, same process
    c4 = None
    if world > 1 and not args.no_c4 and not args.nx_total and args.workload == "dambreak_1m":
        c4 = leg(4096, args.steps, warmup, False)

    ms, n_total, n_local, cells, phases, clocks = main["ms"], main["n_total"], main["n_local"], main["cells"], main["phases"], main["clocks"]
    gx, gy = main["grid"]
    roofline = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_phase_local = main["n_phase_local"]
        dom = max((k for k in phases if k != "exchange"), key=lambda k: phases[k])
        swept = solver == SPH_SOLVER_COLORED_GS and dom in ("viscosity", "delta")
        launches = 9 if (swept and not one_launch) else 1
        sweep_kernel = "color_sweep_flow_kernel" if one_launch else ("color_sweep_team_kernel" if args.sweep == "team" else "color_sweep_kernel")
        dom_bytes = PHASE_BYTES[dom] * n_phase_local / launches
        launch_ms = phases[dom] / launches
        achieved = dom_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
        step_bytes = BYTES_PER_PARTICLE_STEP * n_total + BYTES_PER_CELL_STEP * cells
        step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9
        fpn = "Exact" if args.fp == "exact" else "Fast"
        kernel_name = {"viscosity": f"{sweep_kernel}<{fpn}, 1>" if swept else f"viscosity_kernel<{fpn}>", "delta": f"{sweep_kernel}<{fpn}, 0>" if swept else f"delta_kernel<{fpn}>",
                       "density": f"density_kernel<{fpn}>", "reorder": "reorder_kernel", "predict_key": "predict_key_kernel",
                       "collide_velocity": "collide_velocity_kernel"}.get(dom, dom)
        traffic, issue = None, None
        tpath = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
        if os.path.exists(tpath) and args.workload == "dambreak_1m" and world == 1 and args.fp == "exact" and one_launch:
            prof = json.load(open(tpath)).get(kernel_name, {})
            traffic = prof.get("dram_bytes_per_launch")
            if prof.get("warp_instructions_per_launch") and launch_ms > 0 and clocks.get("sm_mhz"):
                # what actually binds the kernel: warp instructions per launch (ncu, same scene at step 2-3) over the launch
                # time measured here, against the issue rate of the chip (148 SMs x 4 schedulers x 1 instruction per clock)
                peak_ginst = 148 * 4 * clocks["sm_mhz"] * 1e6 / 1e9
                ach_ginst = prof["warp_instructions_per_launch"] / (launch_ms * 1e-3) / 1e9
                issue = {"bound": "instruction issue", "achieved": ach_ginst, "peak": peak_ginst, "unit": "G warp-instructions/s", "frac": ach_ginst / peak_ginst,
                         "warp_instructions_per_launch": prof["warp_instructions_per_launch"], "ncu_issue_slots_busy_pct": prof.get("issue_slots_busy_pct"),
                         "source": f"profiles/{TRAFFIC_FILE} (ncu --set full of this kernel) + kernel_ms measured in this run"}
        # the streaming / grid passes are the HBM-bound part of the step: their own fractions
        hbm_passes = {}
        for name, what in (("integrate", "integrate_kernel"), ("predict_key", "predict_key_kernel"), ("reorder", "scatter + reorder"), ("collide_velocity", "collide_velocity_kernel")):
            if phases.get(name, 0) > 0:
                gbs = PHASE_BYTES[name] * n_phase_local / (phases[name] * 1e-3) / 1e9
                hbm_passes[name] = {"kernel": what, "ms": phases[name], "achieved_gbs": gbs, "frac": gbs / peak}
        if phases.get("scan", 0) > 0:
            gbs = BYTES_PER_CELL_STEP * cells / world / (phases["scan"] * 1e-3) / 1e9
            hbm_passes["scan"] = {"kernel": "cell scan + colour lists", "ms": phases["scan"], "achieved_gbs": gbs, "frac": gbs / peak}
        roofline = {
            "bound": "hbm", "kernel": kernel_name + ((" (the whole %s sweep, nine colours in one launch)" if one_launch else " (one of the 9 colour launches of the %s sweep)") % dom if swept else ""),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": launch_ms,
            "step_model": {"bytes_per_step": step_bytes, "achieved_gbs": step_gbs, "frac": step_gbs / (peak * world)},
            "hbm_passes": hbm_passes,
            "issue": issue,
            "note": "the pair passes (viscosity, density, delta) are instruction-issue bound, not HBM bound (ncu summaries under profiles/): their HBM fraction is small by "
                    "construction (BASELINE.md section 4) and `issue` restates the dominant one against the chip's issue rate; `hbm_passes` are the streaming / grid "
                    "passes, which ARE HBM bound; traffic (ncu, cold L2) is the whole kernel's DRAM bytes per launch; rank 0's phases",
        }

    cpu = cpu_baseline(nx_one, spacing, main["gravity"], relaxation) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {nx}x{nx} = {n_total} particles on {world} GPU(s), spacing {spacing}, h = cell = 0.3, dt = 1/60, "
                                   f"grid {gx}x{gy}, gravity {main['gravity'][1]:.4f}, fp_mode {args.fp}, solver {args.solver}, sweep {args.sweep}"
                                   + (f", strip exchange {args.transport}" if world > 1 else "")
                                   + (f", strips re-balanced every {args.rebalance} steps" if (args.rebalance and world > 1) else ""),
                       "particles": n_total, "particles_rank0": n_local, "cells": cells,
                       "candidates_per_particle_rank0": main["candidates"],
                       "l2": "state streamed once per phase (>130 MB/step) and a 256 MiB L2 flush before the timed region",
                       "parallelism": f"ystrip{world}",
                       "c4": (f"dambreak 16M: 4096x4096 = {c4['n_total']} particles on {world} GPU(s) (BASELINE.json configs[3]), same parameters, "
                              f"gravity {c4['gravity'][1]:.4f}, grid {c4['grid'][0]}x{c4['grid'][1]}, {warmup} warm-up + {args.steps} timed steps") if c4 else None},
            "e2e": main["e2e"],
            "gpu_launches": kernels_per_step * args.steps,
            "clocks": clocks,
            "roofline": roofline,
            "phases_ms": phases,
            "cpu_baseline": cpu,
        }
        if main.get("per_step_ms"):
            line["per_step"] = main["per_step_ms"]
        if c4:
            line.update({"c4_value": c4["value"], "c4_ms_per_step": c4["ms"] / c4["steps"], "c4_e2e": c4["e2e"], "c4_particles_rank0": c4["n_local"],
                         "c4_clocks": c4["clocks"]})
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def oracle_scene(nx, spacing, gravity, threads, mode):
    from oracle_lib import CpuSim

    width = 4.0 * nx * spacing
    height = width * 9.0 / 16.0
    cell = float(np.float32(6.0) * np.float32(0.05))
    sim = CpuSim("oracle", width=width, height=height, cell=cell, mode=mode, threads=threads)
    p = sim.params()
    p[2] = spacing
    sim.put_params(p)
    sim.set_gravity(float(gravity[0]), float(gravity[1]))
    for a, b, d in ((0.0, 1.0, -height / 2), (0.0, -1.0, -height / 2), (1.0, 0.0, -width / 2), (-1.0, 0.0, -width / 2)):
        sim.add_plane(a, b, float(np.float32(d)))
    sim.add_volume(-width / 2 + nx * spacing / 2 + 0.05, -height / 2 + nx * spacing / 2 + 0.05, 0.0, 0.0, nx, nx, spacing)
    return sim


def sized_oracle(nx, spacing, gravity, cores, mode, budget_s, steps_wanted):
    """The oracle on the FULL scene; the block is only halved (nx -> nx/2) while one step x steps_wanted would not fit
    the time budget.  -> (sim, sample_nx, seconds of one warm step)"""
    sample_nx = nx
    while True:
        sim = oracle_scene(sample_nx, spacing, gravity, cores, mode)
        sim.advance(DT, 1)  # the first step has no viscosity pass (no lists yet, demo4.cpp:148)
        t = sim.advance_timed(DT, 1)
        if t * steps_wanted <= budget_s or sample_nx <= 32:
            return sim, sample_nx, t
        sim.close()
        sample_nx //= 2


def cpu_baseline(nx, spacing, gravity, relaxation, budget_s=20.0):
    """The oracle in the reference's own multithreaded mode (in-place pair updates, thread-pool split of
    threading.h:111-129) on all host cores, on the same scene; the sample is bounded in STEPS (about budget_s of CPU
    work), the particle count is only reduced if even four steps would not fit."""
    from oracle_lib import MODE_GS_INDEX, build_oracle

    build_oracle()
    cores = os.cpu_count() or 1
    sim, sample_nx, t = sized_oracle(nx, spacing, gravity, cores, MODE_GS_INDEX, budget_s, 4)
    n = sim.n
    steps = int(max(4, min(1024, budget_s / max(t, 1e-3))))
    secs = sim.advance_timed(DT, steps)
    sim.close()
    return {"value": n * steps / secs, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps of the {sample_nx}x{sample_nx} = {n} particle block"
                      + (" (the full scene)" if sample_nx == nx else f" (the scene is {nx}x{nx}: reduced to fit {budget_s:.0f} s)")
                      + ", oracle gs_index mt mode = the reference's demo-4 multithreaded semantics"}


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm for the same config on the box's host cores.
    libsphref.so (the reference's own code) is hard-capped at 10 000 particles and a 10 x 5.625 domain
    (sph.h:18-72), so the 1M scene runs through the oracle port in the reference's multithreaded mode -
    on the FULL scene (1M particles x (W + K) steps is seconds on a server CPU); the block is only reduced
    if the run would not end within ~3 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle_lib import MODE_GS_INDEX, REF_SO, CpuSim, build_oracle

    build_oracle()
    wl = WORKLOADS[args.workload]
    nx_one = args.nx or wl["nx"]
    # the same scene as our arm at this GPU count (weak scaling: the block grows with N, see run_ours)
    world = max(args.gpus, 1)
    nx = nx_one if world == 1 else int(round(nx_one * world ** 0.5 / 32.0)) * 32
    if args.nx_total:
        nx = args.nx_total
    spacing = wl["spacing"]
    gravity = scene_gravity(nx, spacing, wl["gravity_scale"] or args.scaled_gravity)
    cores = os.cpu_count() or 1
    sim, sample_nx, _ = sized_oracle(nx, spacing, gravity, cores, MODE_GS_INDEX, 170.0, args.steps + args.warmup)
    n = sim.n
    sim.advance(DT, max(args.warmup - 2, 0))
    secs = sim.advance_timed(DT, args.steps)
    sim.close()
    value = n * args.steps / secs
    extra = None
    if os.path.exists(REF_SO):  # the reference's own binary on ITS default scene (configs[0]), for orientation
        ref = CpuSim("ref", threads=cores)
        ref.load_scenario(0, 1)
        ref.advance(DT, 8)
        s = ref.advance_timed(DT, 64)
        extra = {"value": ref.n * 64 / s, "unit": UNIT, "what": "libsphref.so (unmodified demo4.cpp) scenario 0, 5300 particles, 64 steps, MT"}
        ref.close()
    sample = (f"{args.steps} steps of the {sample_nx}x{sample_nx} = {n} particle block"
              + (" (the full scene)" if sample_nx == nx else f" (the scene is {nx}x{nx}: reduced to fit the time limit)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: oracle port of demo4 (gs_index, mt) on {cores} host cores; {sample}, spacing {spacing}, h = cell = 0.3, dt = 1/60, "
                               f"gravity {gravity[1]:.4f}", "particles": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_binary_scene0": extra,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)    # 64 = the reference's own benchmark frame count (app.h:24)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dambreak_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=0, help="override the block edge (particles = nx*nx per GPU)")
    ap.add_argument("--fp", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--solver", default="gs", choices=["gs", "gather"], help="coloured Gauss-Seidel sweeps (default) or Jacobi gather")
    ap.add_argument("--relaxation", type=float, default=0.0, help="gather only: omega")
    ap.add_argument("--sweep-capacity", type=int, default=0)
    ap.add_argument("--sweep", default="auto", choices=["auto", "flow", "warp", "team"],
                    help="coloured sweep kernel: one launch with dependency flags (flow), nine launches warp-per-cell (warp) or block-per-cell (team)")
    ap.add_argument("--halo-rows", type=int, default=0)
    ap.add_argument("--rebalance", type=int, default=0, help="strips: re-balance every N steps by the particles per grid row (0 = static split)")
    ap.add_argument("--no-clock-sampler", action="store_true", help="experiment: do not poll NVML during the timed region")
    ap.add_argument("--nx-total", type=int, default=0, help="edge of the whole block, overriding the weak-scaling rule")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the second leg on the 16M-particle block (BASELINE.json configs[3])")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="strip exchange: records stored straight into the neighbour's memory (default) or fixed-size NCCL messages")
    ap.add_argument("--scaled-gravity", action="store_true", help="scale gravity to the reference scene's hydrostatic head")
    ap.add_argument("--per-step", action="store_true", help="diagnostic: also report every rank's device and host time of every timed step (per_step)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
