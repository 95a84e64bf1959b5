"""TEST INFRASTRUCTURE: random (config, program) pairs — the device engine's state machines (compiled for the
host, tests/enginecheck) against the oracle, per seed, every output field.  Seeds that the bounded
device tables refuse (MSIM_ST_FAIL_CAPACITY — loud by design) are skipped, everything else must be
identical.  `python tests/fuzz_parity.py <rng seed> <seconds>` runs it stand-alone; tests/test_fuzz_parity.py
runs a short fixed slice."""
import os
import random

import numpy as np
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import madsim_b200 as m
import oracle_py as o
from madsim_b200.api import MS, SEC, Config, Workload


def rand_cfg(rnd):
    lo = rnd.choice([1, 1, 1, 50, 999]) * MS
    return Config(packet_loss_rate=rnd.choice([0, 0, 0, 0.01, 0.05, 0.2, 0.5]), send_latency_ns=(lo, lo + rnd.choice([1, 9, 9, 200, 1500]) * MS),
                  buggify=rnd.random() < 0.2, time_limit_ns=rnd.choice([0, 0, 0, 3 * SEC, 30 * SEC]))


def rand_wl(rnd):
    k = rnd.choice(["pingpong", "raft", "raft", "chaos", "stream", "etcd", "etcd", "tonic", "tonic", "signal", "election"])
    if k == "election":
        return Workload.election(hold_s=rnd.randint(20, 50), timeout_rate=rnd.choice([0, 0, 0.05, 0.3, 1.0]))
    if k == "pingpong":
        return Workload.pingpong(rounds=rnd.randint(0, 60), think_ns=rnd.choice([0, MS, 10 * MS]))
    if k == "raft":
        n = rnd.choice([1, 2, 3, 5])
        faults = rnd.randint(0, 7) & (7 if n >= 2 else 5)  # link faults need two nodes
        return Workload.raft(n_nodes=n, duration_ms=rnd.choice([500, 2000, 4000]), faults=faults, bug=rnd.randint(0, 1),
                             rpc_timeout_ms=rnd.choice([20, 40, 80]), heartbeat_ms=rnd.choice([30, 50]),
                             election_ms=rnd.choice([(80, 160), (150, 300)]))
    if k == "chaos":
        panics = rnd.choice([0, 0, 1, 3])
        return Workload.chaos(rounds=rnd.randint(1, 12), period_ms=rnd.choice([100, 700, 2000]), panics=panics,
                              matching=rnd.choice([0, 0, panics, max(panics - 1, 0)]))
    if k == "stream":
        return Workload.stream(messages=rnd.randint(0, 120), think_ns=rnd.choice([0, MS, 5 * MS]), faults=rnd.randint(0, 1))
    if k == "etcd":
        return Workload.etcd(clients=rnd.randint(1, 3), ops=rnd.randint(1, 40), think_ns=rnd.choice([0, 5 * MS, 50 * MS]),
                             timeout_rate=rnd.choice([0, 0, 0.1, 0.5, 1.0]), bug=rnd.randint(0, 3), lease_ttl_s=rnd.randint(0, 3),
                             lease_admin=rnd.random() < 0.5)
    if k == "tonic" and rnd.random() < 0.25:
        return Workload.tonic(clients=rnd.randint(1, 6), rounds=rnd.randint(0, 12), think_ns=rnd.choice([0, 10 * MS, 100 * MS]),
                              restart_only=rnd.random() < 0.3, server_crash=rnd.random() < 0.4, interceptors=True)
    if k == "tonic":
        return Workload.tonic(clients=rnd.randint(1, 6), rounds=rnd.randint(0, 12), think_ns=rnd.choice([0, 10 * MS, 100 * MS]),
                              restart_only=rnd.random() < 0.3, streaming=rnd.random() < 0.5, server_crash=rnd.random() < 0.4, timeouts=rnd.random() < 0.3)
    return Workload.signal(rounds=rnd.randint(0, 10), period_ms=rnd.choice([1, 10, 100]), behavior=rnd.randint(0, 2), catch=rnd.random() < 0.6)


def run(rng_seed: int, seconds: float, seeds_per_case: int = 12):
    """→ (cases, seeds compared, seeds skipped for capacity, seeds escalated to the large-capacity variant);
    raises AssertionError on the first mismatch"""
    rnd = random.Random(rng_seed)
    t0, cases, compared, skipped, escalated = time.time(), 0, 0, 0, 0
    while time.time() - t0 < seconds:
        cfg, wl = rand_cfg(rnd), rand_wl(rnd)
        # programs that cannot end under this config need a time limit (an unreachable peer retries forever)
        if wl.workload == m.WL_TONIC and wl.p[1] == 0 and cfg.time_limit_ns == 0:
            cfg.time_limit_ns = 20 * SEC
        if wl.workload == m.WL_STREAM and cfg.packet_loss_rate >= 0.5 and cfg.time_limit_ns == 0:
            cfg.time_limit_ns = 30 * SEC
        base = rnd.randint(0, 10**6)
        ref = o.run_many(cfg, wl, base, seeds_per_case)
        got, _ = o.engine_host(cfg, wl, base, seeds_per_case)
        for i in np.nonzero(got["status"] == m.ST_FAIL_CAPACITY)[0]:  # what capacity escalation does on the device
            got[i] = o.engine_host_big(cfg, wl, base + int(i), 1)[0]
            escalated += 1
        ok = got["status"] != m.ST_FAIL_CAPACITY
        skipped += int((~ok).sum())
        compared += int(ok.sum())
        for i in range(seeds_per_case):
            assert not ok[i] or ref[i] == got[i], f"{wl} {cfg} seed {base + i}:\n oracle {ref[i]}\n engine {got[i]}"
        cases += 1
    return cases, compared, skipped, escalated


if __name__ == "__main__":
    print("cases %d, seeds compared %d, skipped (capacity, after escalation) %d, escalated %d" % run(int(sys.argv[1]) if len(sys.argv) > 1 else 1,
                                                                    float(sys.argv[2]) if len(sys.argv) > 2 else 60))
