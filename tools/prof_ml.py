"""One batch of any program at a chosen geometry, for profiling under ncu (not a benchmark).
  WL=stream N=16384 S=112 L=4 python tools/prof_ml.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import madsim_b200 as m
from gpu_ml_sweep import WL
cfg, wl = WL[os.environ.get("WL", "stream")]()
n = int(os.environ.get("N", 16384)); S = int(os.environ.get("S", 0)); L = int(os.environ.get("L", 0))
with m.Batch(cfg, 0) as b:
    b.load(wl).set_geometry(S, L)
    b.run(0, n)
    res, bm, st = b.results()
    print(st)
