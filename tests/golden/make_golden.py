"""Generates tests/golden/golden_v1.json from the ORACLE (oracle/liboracle.so).

The reference holds no golden vectors for this path and cannot be built here (no Rust toolchain,
SURVEY.md §8c), so these vectors pin the oracle + engine against regressions; they are NOT outputs
of the reference itself (parity vs the reference stays "unpinned").
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_py as o

import madsim_b200 as m

CASES = {
    "sleep_1s": (dict(), dict(kind="sleep", duration_ns=10**9), [0, 1, 2, 2333]),
    "pingpong_r1000": (dict(), dict(kind="pingpong", rounds=1000), [0, 1, 2333]),  # BASELINE config 1
    "pingpong_r50_loss10": (dict(packet_loss_rate=0.1), dict(kind="pingpong", rounds=50), list(range(8))),
    "raft_cfg2": (dict(), dict(kind="raft"), list(range(16)) + [4095, 65535]),  # BASELINE config 2
    "raft_cfg3": (dict(), dict(kind="raft", faults=7), list(range(16)) + [4095, 65535]),  # config 3
    "raft_cfg2_bug": (dict(), dict(kind="raft", bug=1), [88, 89, 90]),
    "stream_partitions": (dict(), dict(kind="stream"), list(range(6))),
    "stream_loss20": (dict(packet_loss_rate=0.2), dict(kind="stream", messages=500), [0, 1]),
    "chaos": (dict(), dict(kind="chaos"), list(range(6))),
    "chaos_panics": (dict(), dict(kind="chaos", rounds=30, period_ms=700, panics=3), list(range(6))),
    "etcd_cfg4": (dict(), dict(kind="etcd"), list(range(8)) + [16383]),  # BASELINE config 4
    "etcd_cfg4_loss30": (dict(packet_loss_rate=0.3), dict(kind="etcd"), list(range(4))),
    "etcd_cfg4_timeouts_bug": (dict(packet_loss_rate=0.1), dict(kind="etcd", timeout_rate=0.125, bug=3), list(range(8))),
    "tonic_cfg5": (dict(), dict(kind="tonic"), list(range(6)) + [32767]),  # BASELINE config 5
    "tonic_cfg5_restart_only": (dict(), dict(kind="tonic", clients=1, restart_only=True), list(range(4))),
    "signal_burst": (dict(), dict(kind="signal", behavior=0), list(range(4))),
    "signal_delay": (dict(), dict(kind="signal", behavior=1), list(range(4))),
    "signal_skip_no_handler": (dict(), dict(kind="signal", behavior=2, catch=False), list(range(4))),
    "tonic_cfg5_streaming": (dict(), dict(kind="tonic", streaming=True), list(range(6))),
    "etcd_cfg4_lease_admin": (dict(packet_loss_rate=0.1), dict(kind="etcd", lease_admin=True, timeout_rate=0.125), list(range(6))),
    "tonic_server_crash": (dict(), dict(kind="tonic", streaming=True, server_crash=True), list(range(6))),
    "tonic_request_timeouts": (dict(), dict(kind="tonic", streaming=True, timeouts=True), list(range(4))),
    "raft_cfg3_loss5": (dict(packet_loss_rate=0.05), dict(kind="raft", faults=7), list(range(4))),
    "tonic_interceptors": (dict(), dict(kind="tonic", interceptors=True), list(range(4))),
    "election": (dict(), dict(kind="election"), list(range(6))),
    "election_timeouts": (dict(), dict(kind="election", timeout_rate=0.125, hold_s=25), list(range(8))),
    # seeds 5.. of this leaking configuration overflow the default mailbox / message tables about every other time: the device
    # answers them from the large-capacity variant (capacity escalation)
    "raft_leaky_300ms_escalated": (dict(), dict(kind="raft", duration_ms=300, rpc_timeout_ms=3, heartbeat_ms=20, election_ms=(40, 80), faults=4),
                                   list(range(5, 21))),
    "tonic_server_crash_slow_links_escalated": (dict(packet_loss_rate=0.05, send_latency_ns=(999000000, 2499000000), time_limit_ns=30 * 10**9),
                                                dict(kind="tonic", clients=4, think_ns=100 * 10**6, server_crash=True, timeouts=True), [626266]),
}


def make_wl(d):
    d = dict(d)
    kind = d.pop("kind")
    return getattr(m.Workload, kind)(**d)


def main():
    out = {"format": 1, "fields": ["trace_hash", "elapsed_ns", "draws", "ev_pop", "ev_fire", "status", "fail_code",
                                   "user"], "cases": {}}
    for name, (cfg, wl, seeds) in CASES.items():
        rows = []
        for s in seeds:
            r, _, _, _ = o.run_seed(m.Config(**cfg), make_wl(wl), s)
            rows.append([s, str(int(r["trace_hash"])), int(r["elapsed_ns"]), int(r["draws"]), int(r["ev_pop"]),
                         int(r["ev_fire"]), int(r["status"]), int(r["fail_code"]), [int(x) for x in r["user"]]])
        out["cases"][name] = {"config": cfg, "workload": wl, "rows": rows}
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
