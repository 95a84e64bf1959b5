"""Random (config, program) pairs on the DEVICE through the C-ABI against the oracle (exploration / evidence
script; the same generator as tests/fuzz_parity.py).  usage: python tools/gpu_fuzz.py <rng seed> <seconds>"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import fuzz_parity as fz
import madsim_b200 as m
import oracle_py as o
from madsim_b200.api import SEC


def main():
    rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60
    t0, cases, compared, skipped, escalated = time.time(), 0, 0, 0, 0
    by = {}
    while time.time() - t0 < seconds:
        cfg, wl = fz.rand_cfg(rnd), fz.rand_wl(rnd)
        if wl.workload == m.WL_TONIC and wl.p[1] == 0 and cfg.time_limit_ns == 0:
            cfg.time_limit_ns = 20 * SEC
        if wl.workload == m.WL_STREAM and cfg.packet_loss_rate >= 0.5 and cfg.time_limit_ns == 0:
            cfg.time_limit_ns = 30 * SEC
        base, n = rnd.randint(0, 10**6), 40
        ref = o.run_many(cfg, wl, base, n)
        with m.Batch(cfg, 0) as b:
            b.load(wl).run(base, n)
            got, _, st = b.results()
        ok = got["status"] != m.ST_FAIL_CAPACITY
        for i in range(n):
            if ok[i] and not all(np.array_equal(ref[f][i], got[f][i]) for f in ref.dtype.names):
                print("MISMATCH", wl, cfg, base + i, ref[i], got[i])
                sys.exit(1)
        cases += 1
        compared += int(ok.sum())
        skipped += int((~ok).sum())
        escalated += int(st["escalated"])
        by[wl.name()] = by.get(wl.name(), 0) + 1
    print(f"gpu fuzz: cases {cases}, seeds compared {compared}, skipped (capacity, after escalation) {skipped}, escalated {escalated}, by program {by}")


if __name__ == "__main__":
    main()
