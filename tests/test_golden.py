"""Committed golden vectors (tests/golden/golden_v1.json, made by tests/golden/make_golden.py):
regression pins for the oracle and for the engine's state machines (host-compiled here; the GPU
twin of this test lives in test_gpu_parity.py)."""
import numpy as np
import pytest

import madsim_b200 as m


def make_wl(d):
    d = dict(d)
    kind = d.pop("kind")
    return getattr(m.Workload, kind)(**d)


def rows_of(case):
    return [(r[0], int(r[1]), r[2], r[3], r[4], r[5], r[6], r[7], tuple(r[8])) for r in case["rows"]]


def rec_tuple(seed, r):
    return (seed, int(r["trace_hash"]), int(r["elapsed_ns"]), int(r["draws"]), int(r["ev_pop"]), int(r["ev_fire"]),
            int(r["status"]), int(r["fail_code"]), tuple(int(x) for x in r["user"]))


def test_golden_has_baseline_configs(golden):
    assert {"pingpong_r1000", "raft_cfg2", "raft_cfg3"} <= set(golden["cases"])


def test_oracle_matches_golden(oracle, golden):
    for name, case in golden["cases"].items():
        cfg, wl = m.Config(**case["config"]), make_wl(case["workload"])
        for want in rows_of(case):
            r, _, _, _ = oracle.run_seed(cfg, wl, want[0])
            assert rec_tuple(want[0], r) == want, name


def test_engine_host_matches_golden(oracle, golden):
    for name, case in golden["cases"].items():
        cfg, wl = m.Config(**case["config"]), make_wl(case["workload"])
        for want in rows_of(case):
            got, _ = oracle.engine_host(cfg, wl, want[0], 1)
            if got["status"][0] == m.ST_FAIL_CAPACITY:  # what capacity escalation does on the device
                got = oracle.engine_host_big(cfg, wl, want[0], 1)
            assert rec_tuple(want[0], got[0]) == want, name


def test_pingpong_is_about_14_events_per_round(golden):
    # SURVEY §8d "Expect ≈ 14 events per round" for BASELINE config 1
    for r in rows_of(golden["cases"]["pingpong_r1000"]):
        assert 13.5 < (r[4] + r[5]) / 1000 < 14.5 and r[8][0] == 1000 and r[6] == m.ST_PASS
