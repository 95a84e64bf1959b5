"""GPU exploration script (not part of the product): geometry sweep for the step kernel."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import madsim_b200 as m


def main():
    n = int(os.environ.get("N", 4096))
    faults = int(os.environ.get("FAULTS", 0))
    kind = os.environ.get("WL", "raft")
    wl = m.Workload.raft(faults=faults) if kind == "raft" else getattr(m.Workload, kind)()
    out = []
    geos = [tuple(int(x) for x in g.split(":")) for g in os.environ.get("GEOS", "0:1,0:2,0:4,0:8").split(",")]
    for S, L in geos:
        with m.Batch(m.Config(), 0) as b:
            b.load(wl).set_geometry(S, L)
            b.run(0, n)  # warm
            t = time.time()
            b.run(0, n)
            res, bitmap, st = b.results()
            dt = time.time() - t
        ev = int(st["ev_pop"] + st["ev_fire"])
        rec = dict(S=st["seeds_per_cta"], L=st["lanes_per_warp"], ctas=st["ctas"], smem=st["smem_bytes"],
                   kernel_ms=st["kernel_ms"], e2e_ms=dt * 1e3, events=ev, mev_s=ev / st["kernel_ms"] / 1e3,
                   n_fail=st["n_fail"], unfinished=st["n_unfinished"],
                   status=np.bincount(res["status"], minlength=6).tolist())
        rec["wl"] = kind
        print(json.dumps(rec), flush=True)
        out.append(rec)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{kind}_n{n}_f{faults}_{os.environ.get('TAG','x')}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
