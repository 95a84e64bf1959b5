"""One Raft batch for profiling under ncu (not a benchmark)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import madsim_b200 as m
n = int(os.environ.get("N", 4096)); S = int(os.environ.get("S", 0)); L = int(os.environ.get("L", 0))
dur = int(os.environ.get("DUR", 10000))
with m.Batch(m.Config(), 0) as b:
    b.load(m.Workload.raft(duration_ms=dur, faults=int(os.environ.get("FAULTS", 0)))).set_geometry(S, L)
    b.run(0, n)
    res, bm, st = b.results()
    print(st)
