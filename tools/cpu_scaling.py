"""How the CPU oracle scales with threads on this box (for the cpu_baseline / reference arm)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as o, madsim_b200 as m
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: pass
wl = m.Workload.raft()
for jobs in (1, 2, 4, 8, 16, 32, 64, 128):
    n = max(64, jobs * 16)
    t = time.time(); r = o.run_many(m.Config(), wl, 0, n, jobs=jobs); dt = time.time() - t
    print(jobs, n, "%.1f Mev/s" % ((int(r["ev_pop"].sum()) + int(r["ev_fire"].sum())) / dt / 1e6), flush=True)
