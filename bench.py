#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: simulated events/sec over N seeds at 1/2/4/8 B200.

One "step" = one pass of the hot path over one batch: every seed of the batch runs the MadRaft-style
5-node leader-election workload (BASELINE config 2: 4096 seeds per GPU, 10 virtual seconds each)
from Runtime construction to block_on's return.  Seeds shard contiguously over ranks with no
data-path collective (weak scaling: 4096 seeds per GPU); the only exchange is the end-of-run
all-reduce of the pass/fail summary (NCCL), which is part of the e2e timing.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

The reference (Rust) cannot be built here: the reference arm times the oracle — a C++ restatement
of the reference's executor (oracle/, "kind": "port") — on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- algorithmic bytes per event (restated in DESIGN.md §Roofline) ----
B_POP = 168           # one ready-queue pop (POLL/DROP)
B_FIRE_BASE = 92      # one timer fire from a heap of n entries: 92 + 40*floor(log2 n)
B_FIRE_PER_LEVEL = 40
MEAN_LOG2_HEAP = 3.8  # measured on the oracle for this workload (mean floor(log2 n) at fire: 3.70 / 3.87)
B_PUSH = 32           # one timer push (>= 32 B)
B_MSG = 40            # one datagram send/deliver


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel at the bench configuration,
# from the `ncu --set full` capture summarised in profiles/r1c_step_kernel_v2_bench_config.json
TRAFFIC_BYTES_PER_LAUNCH = {"raft": 67072}


def algorithmic_bytes(ev_pop, ev_fire, msgs):
    pushes = ev_fire  # every fired timer was pushed once
    return ev_pop * B_POP + ev_fire * (B_FIRE_BASE + B_FIRE_PER_LEVEL * MEAN_LOG2_HEAP) + pushes * B_PUSH + msgs * B_MSG


def effective_cores():
    """Host cores this process may actually use: min(visible CPUs, cgroup CPU quota).  The GPU boxes
    expose 128 logical CPUs but run the container under a 16-core CFS quota; oversubscribing a
    quota throttles the whole process group (measured: 128 threads are slower than 16)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return n


def best_thread_count(o, cfg, wl):
    """all the host threads it can use: calibrate quota, 2x quota (SMT) and take the fastest"""
    cores = effective_cores()
    best, best_rate = cores, 0.0
    for jobs in sorted({cores, min(2 * cores, os.cpu_count() or cores)}):
        n = max(64, 8 * jobs)
        o.run_many(cfg, wl, 7_000_000, min(n, 64), jobs=jobs)
        t = time.perf_counter()
        r = o.run_many(cfg, wl, 7_100_000, n, jobs=jobs)
        rate = (int(r["ev_pop"].sum()) + int(r["ev_fire"].sum())) / (time.perf_counter() - t)
        if rate > best_rate:
            best, best_rate = jobs, rate
    return best, cores


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for i, nm in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_for(args):
    import madsim_b200 as m

    if args.workload == "raft":
        return m.Workload.raft(), "MadRaft-style 5-node leader election, 10 virtual s (BASELINE config 2)"
    if args.workload == "raft_faults":
        return m.Workload.raft(faults=7), "5-node election + partition/clock-jump faults (BASELINE config 3 shard)"
    if args.workload == "pingpong":
        return m.Workload.pingpong(), "2-node ping-pong R=1000 (BASELINE config 1, batched)"
    raise SystemExit(f"unknown workload {args.workload}")


def run_reference_arm(args, rank):
    """The reference's own CPU algorithm on the host cores (oracle port; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as o

    import madsim_b200 as m

    wl, desc = workload_for(args)
    n = args.ref_seeds
    cfg = m.Config()
    cores, quota = best_thread_count(o, cfg, wl)
    for w in range(args.warmup):
        o.run_many(cfg, wl, 10_000_000 + w * n, min(n, 256), jobs=cores)
    t_total, events = 0.0, 0
    for s in range(args.steps):
        t = time.perf_counter()
        res = o.run_many(cfg, wl, s * n, n, jobs=cores)
        t_total += time.perf_counter() - t
        events += int(res["ev_pop"].sum()) + int(res["ev_fire"].sum())
    value = events / t_total
    line = {
        "impl": "reference", "metric": "simulated events/sec", "value": value, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": desc, "seeds_per_step": n, "virtual_seconds": 10,
                   "note": "oracle = C++20 restatement of the reference executor; the Rust crate cannot be built here"},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port",
                         "host": {"logical_cpus": os.cpu_count(), "cgroup_quota_cores": quota},
                         "sample": f"{n} seeds x {args.steps} steps of the same workload, one runtime per task, {cores} threads"},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="msim", choices=["msim", "reference"])
    ap.add_argument("--workload", default="raft", choices=["raft", "raft_faults", "pingpong"])
    ap.add_argument("--seeds-per-gpu", type=int, default=4096)
    ap.add_argument("--ref-seeds", type=int, default=4096, help="reference arm: seeds per step (bounded sample)")
    ap.add_argument("--cpu-baseline-seeds", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seeds-per-cta", type=int, default=0)
    ap.add_argument("--lanes-per-warp", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import madsim_b200 as m
    from madsim_b200 import dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: madsim_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl, desc = workload_for(args)
    cfg = m.Config()
    n_local = args.seeds_per_gpu
    n_total = n_local * world

    run = D.ShardedRun(cfg, wl, local_rank)
    run.batch.set_geometry(args.seeds_per_cta, args.lanes_per_warp)
    stream = run.stream
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(step_idx):
        """device-timed: the step kernel only; inputs (seed range + config) already resident"""
        first = step_idx * n_total + rank * n_local
        run.batch.run(first, n_local, sync=False)

    # ---------------- warm-up (both paths: kernels, NCCL communicator, pinned staging) ----------------
    for w_i in range(max(args.warmup, 3)):
        step_resident(1_000_000 + w_i)
        run.batch.sync()
        run.run((2_000 + w_i) * n_total, n_total, sync=False)
        run.reduce()
        run.batch.results()
    barrier()

    # ---------------- device-timed steps (value) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    kernel_ms, ev_pop, ev_fire, msgs, launches = [], 0, 0, 0, 0
    for s in range(args.steps):
        flush.add_(1)  # L2 flush between timed iterations (state is on-chip; nothing is cached across steps anyway)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            step_resident(s)
            ev1.record(stream)
        stream.synchronize()
        barrier()
        kernel_ms.append(ev0.elapsed_time(ev1))
        res, bm, st = run.batch.results()
        ev_pop += st["ev_pop"]
        ev_fire += st["ev_fire"]
        msgs += int(res["user"][:, 3].sum()) if wl.workload == m.WL_RAFT else 0
        launches += 1
        assert st["n_unfinished"] == 0
    local_ms = float(sum(kernel_ms))
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([ev_pop, ev_fire, msgs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms = float(t.item())
    g_pop, g_fire, g_msgs = (int(x) for x in tot.tolist())
    value = (g_pop + g_fire) / (max_ms / 1e3)

    # ---------------- end-to-end steps through the public API (e2e) ----------------
    # each step: host config/seed range → C-ABI run → device summary → (NCCL all-reduce) → host
    # results buffers (64 B per seed + bitmap + counters)
    e2e_ms = []
    h2d = 512  # kernel arguments (Params digest + seed range) uploaded per launch
    d2h = n_local * 64 + ((n_local + 31) // 32) * 4 + 64
    for s in range(args.steps):
        flush.add_(1)
        barrier()
        t0 = time.perf_counter()
        run.run((100 + s) * n_total, n_total, sync=False)
        out, gbitmap = run.reduce()
        res, bm, st = run.batch.results()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_ms.append(dt * 1e3)
        launches += 2
        assert out["n_seeds"] == n_total
    te = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_events_per_step = out["ev_pop"] + out["ev_fire"]
    e2e_value = e2e_events_per_step * args.steps / (float(te.item()) / 1e3)
    sampler.stop()
    clocks = sampler.summary()

    # ---------------- roofline of the dominant kernel (msim_step_kernel<WL_Raft>) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    alg_bytes = algorithmic_bytes(ev_pop, ev_fire, msgs if world == 1 else msgs)  # this rank's launches
    achieved = alg_bytes / (local_ms / 1e3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(args.workload), "kernel": "msim_step_kernel<WL_Raft>", "peak_source": peak_src,
        "bytes_per_launch": alg_bytes / args.steps, "launch_ms": local_ms / args.steps,
        "note": "algorithmic bytes (SURVEY §8d model: 168 B/pop, 92+40*3.8 B/fire, 32 B/push, 40 B/msg). The kernel keeps "
                "all per-seed state in shared memory: measured DRAM traffic is ~0.07 MB per launch (profiles/r1l), so the "
                "kernel is instruction-issue/latency bound, not HBM bound; frac is reported against HBM as the contract asks.",
    }

    line = {
        "metric": "simulated events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": desc, "seeds_per_gpu": n_local, "seeds_total": n_total, "virtual_seconds": 10,
                   "events_per_step": (g_pop + g_fire) // args.steps, "parallelism": f"seed-sharded x{world}",
                   "l2": "flushed (256 MiB write) between timed iterations; per-seed state is on-chip",
                   "geometry": {k: st[k] for k in ("seeds_per_cta", "lanes_per_warp", "ctas", "smem_bytes")}},
        "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(te.item()) / args.steps,
                "includes": "C-ABI run + device summary + NCCL all-reduce of the pass/fail summary + D2H of per-seed results"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }

    # ---------------- CPU baseline on this box's host cores (rank 0, N=1 only) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as o

        cores, quota = best_thread_count(o, cfg, wl)
        nb = args.cpu_baseline_seeds
        o.run_many(cfg, wl, 5_000_000, min(nb, 256), jobs=cores)
        t0 = time.perf_counter()
        ref = o.run_many(cfg, wl, 0, nb, jobs=cores)
        dt = time.perf_counter() - t0
        ev = int(ref["ev_pop"].sum()) + int(ref["ev_fire"].sum())
        # the same seeds were run on the GPU in timed step 0: parity is asserted in the same run
        run.batch.run(0, min(nb, n_local))
        got, _, _ = run.batch.results()
        k = min(nb, n_local)
        parity = bool(np.array_equal(got["trace_hash"][:k], ref["trace_hash"][:k]))
        t1 = time.perf_counter()
        one = o.run_many(cfg, wl, 0, 64, jobs=1)
        dt1 = time.perf_counter() - t1
        line["cpu_baseline"] = {
            "value": ev / dt, "unit": "events/s", "cores": cores, "kind": "port",
            "host": {"logical_cpus": os.cpu_count(), "cgroup_quota_cores": quota},
            "sample": f"{nb} seeds of the same workload (one full step), one runtime per task on {cores} threads",
            "single_core_events_per_s": (int(one["ev_pop"].sum()) + int(one["ev_fire"].sum())) / dt1,
            "parity_same_run": parity,
        }
        if not parity:
            raise SystemExit("PARITY FAILURE: device trace hashes differ from the oracle")
    if rank == 0:
        print(json.dumps(line), flush=True)
    run.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
