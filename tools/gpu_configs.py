"""Runs every BASELINE.json config at its full size on ONE B200 and records device time and
events/s (exploration / evidence script, not part of the product; bench.py stays the contract).
Writes gpurun_out/configs_<TAG>.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import madsim_b200 as m

CONFIGS = [
    ("config1 pingpong x1 seed", m.Config(), m.Workload.pingpong(rounds=1000), 1),
    ("config2 raft 4096 seeds", m.Config(), m.Workload.raft(), 4096),
    ("config3 raft+faults, 8192-seed shard of 65536", m.Config(), m.Workload.raft(faults=7), 8192),
    ("config3 raft+faults, all 65536 seeds on one GPU", m.Config(), m.Workload.raft(faults=7), 65536),
    ("config4 etcd 16384 seeds loss 0%", m.Config(), m.Workload.etcd(), 16384),
    ("config4 etcd 16384 seeds loss 30%", m.Config(packet_loss_rate=0.3), m.Workload.etcd(), 16384),
    ("config5 tonic 32768 seeds", m.Config(), m.Workload.tonic(), 32768),
    ("stream 16384 seeds", m.Config(), m.Workload.stream(), 16384),
    ("chaos 16384 seeds", m.Config(), m.Workload.chaos(), 16384),
    ("signal 16384 seeds", m.Config(), m.Workload.signal(), 16384),
]


def main():
    out = []
    for name, cfg, wl, n in CONFIGS:
        with m.Batch(cfg, 0) as b:
            b.load(wl)
            b.run(0, n)  # warm-up
            b.run(0, n)
            res, bitmap, st = b.results()
        ev = int(st["ev_pop"] + st["ev_fire"])
        rec = dict(config=name, seeds=n, state_bytes=m.workload_state_bytes(wl.workload) if hasattr(m, "workload_state_bytes") else None,
                   seeds_per_cta=st["seeds_per_cta"], ctas=st["ctas"], kernel_ms=st["kernel_ms"], events=ev,
                   events_per_s=ev / (st["kernel_ms"] * 1e-3), n_pass=int(st["n_pass"]), n_fail=int(st["n_fail"]),
                   n_capacity=int((res["status"] == m.ST_FAIL_CAPACITY).sum()), unfinished=int(st["n_unfinished"]))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"configs_{os.environ.get('TAG', 'x')}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
