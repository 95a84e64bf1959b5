"""GPU exploration script (not part of the product): single-lane batch kernel against the multi-lane (lock-step
phase) kernel over (seeds per CTA, lanes per warp), each geometry checked against the oracle on a seed window.

  CASES="stream:16384,chaos:16384" GEOS="28:1,56:2,112:4" python tools/gpu_ml_sweep.py
Writes gpurun_out/ml_sweep_<TAG>.json and a text table to stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import madsim_b200 as m
import oracle_py as o

WL = {
    "stream": lambda: (m.Config(), m.Workload.stream()),
    "chaos": lambda: (m.Config(), m.Workload.chaos()),
    "signal": lambda: (m.Config(), m.Workload.signal()),
    "pingpong": lambda: (m.Config(), m.Workload.pingpong(rounds=200)),
    "sleep": lambda: (m.Config(), m.Workload.sleep()),
    "etcd": lambda: (m.Config(), m.Workload.etcd()),
    "etcd30": lambda: (m.Config(packet_loss_rate=0.3), m.Workload.etcd()),
    "raft": lambda: (m.Config(), m.Workload.raft()),
    "raft2s": lambda: (m.Config(), m.Workload.raft(duration_ms=2000)),
    "raftf": lambda: (m.Config(), m.Workload.raft(faults=7)),
    "tonic": lambda: (m.Config(), m.Workload.tonic()),
    "election": lambda: (m.Config(), m.Workload.election()),
}
FIELDS = ("trace_hash", "elapsed_ns", "draws", "ev_pop", "ev_fire", "status", "fail_code")


def main():
    cases = [c.split(":") for c in os.environ.get("CASES", "stream:16384,chaos:16384,signal:16384").split(",")]
    # "S:L" lets the L lanes of a warp run free (the batch kernel); "S:Lk" sends them through the lock-step kernel
    geos = []
    for g in os.environ.get("GEOS", "28:1,56:2,112:4").split(","):
        S, L = g.split(":")
        geos.append((int(S), int(L.rstrip("k")) | (m.GEOMETRY_LOCKSTEP if L.endswith("k") else 0)))
    check = int(os.environ.get("CHECK", 256))
    out = []
    print(f"# {'wl':8s} {'seeds':>6s} {'S':>4s} {'L':>3s} {'warps':>5s} {'kernel_ms':>10s} {'Mev/s':>9s}  parity")
    for name, n in cases:
        n = int(n)
        cfg, wl = WL[name]()
        ref = o.run_many(cfg, wl, 0, min(check, n)) if check else None
        for S, L in geos:
            try:
                with m.Batch(cfg, 0) as b:
                    b.load(wl).set_geometry(S, L)
                    b.run(0, n)  # warm
                    best = 1e30
                    for _ in range(3):
                        b.run(0, n)
                        res, bitmap, st = b.results()
                        best = min(best, st["kernel_ms"])
            except Exception as e:  # geometry does not fit shared memory
                print(f"  {name:8s} {n:6d} {S:4d} {L & 0xff:3d}  -- {e}")
                continue
            ev = int(st["ev_pop"] + st["ev_fire"])
            ok = None
            if ref is not None:
                k = len(ref)
                ok = all(np.array_equal(res[f][:k], ref[f]) for f in FIELDS)
            rec = dict(wl=name, seeds=n, S=st["seeds_per_cta"], L=st["lanes_per_warp"], lockstep=bool(L & m.GEOMETRY_LOCKSTEP),
                       ctas=st["ctas"],
                       smem=st["smem_bytes"], kernel_ms=best, events=ev, mev_s=ev / best / 1e3, parity=ok,
                       n_cap=int((res["status"] == m.ST_FAIL_CAPACITY).sum()))
            warps = (rec["S"] + rec["L"] - 1) // rec["L"]
            tag = "k" if rec["lockstep"] else " "
            print(f"  {name:8s} {n:6d} {rec['S']:4d} {rec['L']:3d}{tag} {warps:4d} {best:10.3f} {rec['mev_s']:9.1f}  {ok}  cap={rec['n_cap']}",
                  flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"ml_sweep_{os.environ.get('TAG', 'x')}.json"), "w"), indent=1)
    if any(r["parity"] is False for r in out):
        raise SystemExit("PARITY FAILURE in the sweep")


if __name__ == "__main__":
    main()
