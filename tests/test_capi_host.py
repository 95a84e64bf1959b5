"""The C-ABI library loads and exports every symbol include/msim_b200.h declares; host-side
logic (parameter digests, Builder env parsing, sharding) — no compute, no GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import madsim_b200 as m
from madsim_b200 import _native as N
from madsim_b200 import dist as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "msim_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(msim_[a-z_0-9]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    lib = C.CDLL(N.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"libmsim_b200.so does not export {sym}"
    assert declared == set(N.PROTOTYPES), "python prototypes and the header disagree"
    assert N.load().msim_version() == 0x00010000


def test_struct_layouts_match_header():
    assert C.sizeof(N.MsimSeedResult) == 64 and C.sizeof(N.MsimTraceRec) == 24
    assert C.sizeof(N.MsimConfig) == 40 and C.sizeof(N.MsimWorkloadDesc) == 72
    assert m.SEED_RESULT_DTYPE.itemsize == 64 and m.TRACE_DTYPE.itemsize == 24


def test_defaults_and_validation_need_no_device():
    lib = N.load()
    cfg = N.MsimConfig()
    assert lib.msim_config_default(C.byref(cfg)) == 0
    assert (cfg.packet_loss_rate, cfg.send_latency_lo_ns, cfg.send_latency_hi_ns) == (0.0, 1_000_000, 10_000_000)
    d = N.MsimWorkloadDesc()
    assert lib.msim_workload_default(N.WL_RAFT, C.byref(d)) == 0
    assert (d.n_nodes, list(d.p)[:7]) == (5, [10000, 0, 0, 40, 50, 150, 300])
    assert lib.msim_workload_default(99, C.byref(d)) == N.EINVAL
    # Raft must keep 28 seeds per 227 KB CTA (4096 seeds = 147 CTAs = one wave on 148 SMs)
    assert 1000 < lib.msim_workload_state_bytes(N.WL_RAFT) <= (232448 - 128) // 28
    assert lib.msim_workload_state_bytes(77) == N.EINVAL
    assert lib.msim_workload_default(N.WL_CHAOS, C.byref(d)) == 0 and list(d.p)[:2] == [8, 2000]
    assert lib.msim_workload_default(N.WL_ETCD, C.byref(d)) == 0
    assert (d.n_nodes, list(d.p)[:5]) == (3, [32, 50_000_000, 0, 0, 1])
    assert m.Workload.etcd().p == list(d.p) and m.Workload.etcd().n_nodes == 3
    assert m.Workload.etcd(timeout_rate=0.5).p[2] == 0x3F000000  # SimServer::timeout_rate is an f32
    assert lib.msim_workload_default(N.WL_SIGNAL, C.byref(d)) == 0 and m.Workload.signal().p == list(d.p)
    assert lib.msim_workload_default(N.WL_TONIC, C.byref(d)) == 0
    assert (d.n_nodes, list(d.p)[:3]) == (7, [10, 100_000_000, 0]) and m.Workload.tonic().p == list(d.p)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(m.MsimError) as e:
        m.device_count()
    assert e.value.code == N.ENODEV
    with pytest.raises(m.MsimError):
        m.Batch()
    with pytest.raises(m.MsimError):
        m.Builder(count=4).run(m.Workload.pingpong(rounds=1))


def test_builder_from_env_reads_the_reference_variables(monkeypatch):
    # runtime/builder.rs:26-63
    monkeypatch.setenv("MADSIM_TEST_SEED", "42")
    monkeypatch.setenv("MADSIM_TEST_NUM", "1000")
    monkeypatch.setenv("MADSIM_TEST_JOBS", "8")
    monkeypatch.setenv("MADSIM_TEST_TIME_LIMIT", "2.5")
    monkeypatch.setenv("MADSIM_TEST_CHECK_DETERMINISM", "1")
    b = m.Builder.from_env()
    assert (b.seed, b.count, b.jobs, b.time_limit_s, b.check) == (42, 1000, 8, 2.5, True)
    monkeypatch.delenv("MADSIM_TEST_CHECK_DETERMINISM")
    assert m.Builder.from_env().check is False


def test_builder_from_env_defaults_follow_the_reference(monkeypatch, tmp_path):
    # builder.rs:64-118: unset seed = the current time in ns; MADSIM_TEST_CONFIG names a TOML file; count >= 2 under the check
    import time

    for v in ("MADSIM_TEST_SEED", "MADSIM_TEST_NUM", "MADSIM_TEST_JOBS", "MADSIM_TEST_TIME_LIMIT", "MADSIM_TEST_CONFIG",
              "MADSIM_TEST_CHECK_DETERMINISM"):
        monkeypatch.delenv(v, raising=False)
    t0 = time.time_ns()
    b = m.Builder.from_env()
    assert t0 <= b.seed <= time.time_ns() and (b.count, b.jobs, b.check, b.time_limit_s) == (1, 1, False, None)
    assert b.config == m.Config()
    monkeypatch.setenv("MADSIM_TEST_CHECK_DETERMINISM", "")
    assert m.Builder.from_env().count == 2
    monkeypatch.setenv("MADSIM_TEST_NUM", "7")
    assert m.Builder.from_env().count == 7
    cfg = tmp_path / "sim.toml"
    cfg.write_text("[net]\npacket_loss_rate = 0.25\nsend_latency = { start = { secs = 0, nanos = 2000000 }, "
                   "end = { secs = 1, nanos = 5 } }\n\n[tcp]\n")
    monkeypatch.setenv("MADSIM_TEST_CONFIG", str(cfg))
    b = m.Builder.from_env()
    assert b.config.packet_loss_rate == 0.25 and tuple(b.config.send_latency_ns) == (2_000_000, 1_000_000_005)


def test_config_from_toml_matches_the_reference_parse_test():
    # the reference's own parse test (config.rs:52-75): inline tables, an empty [tcp]
    c = m.Config.from_toml("""
        [net]
        packet_loss_rate = 0.1
        send_latency = { start = { secs = 0, nanos = 1000000 }, end = { secs = 0, nanos = 10000000 } }

        [tcp]
        """)
    assert c == m.Config(packet_loss_rate=0.1, send_latency_ns=(1 * m.MS, 10 * m.MS))
    # what `impl Display for Config` prints (toml::to_string_pretty): dotted tables
    c = m.Config.from_toml("[net]\npacket_loss_rate = 0.0\n\n[net.send_latency.start]\nsecs = 0\nnanos = 1000000\n\n"
                           "[net.send_latency.end]\nsecs = 0\nnanos = 10000000\n\n[tcp]\n")
    assert c == m.Config()
    assert m.Config.from_toml("") == m.Config()  # #[serde(default)] everywhere
    assert m.Config.from_toml("[net]\npacket_loss_rate = 1 # all lost\nfuture_knob = \"x\"\n").packet_loss_rate == 1.0
    for bad in ("[net\n", "[net]\nsend_latency = { start = { secs = 0 } }\n", "[net]\npacket_loss_rate = \n", "x = [1, 2]\n"):
        with pytest.raises(m.MsimError) as e:
            m.Config.from_toml(bad)
        assert e.value.code == N.EINVAL and "config" in str(e.value)


def test_simulation_failure_message_matches_reference_note():
    e = m.SimulationFailure(1234, m.ST_FAIL_ASSERT, 1)
    assert "MADSIM_TEST_SEED=1234" in str(e)  # runtime/mod.rs:205-210


def test_shard_covers_range_without_overlap():
    for base, n, world in ((0, 4096, 8), (7, 65536, 8), (100, 10, 4), (0, 3, 8), (5, 1, 2)):
        parts = [D.shard(base, n, r, world) for r in range(world)]
        assert sum(c for _, c in parts) == n
        cur = base
        for first, c in parts:
            assert first == cur
            cur += c


def test_summary_from_results_matches_manual():
    res = np.zeros(70, dtype=m.SEED_RESULT_DTYPE)
    res["status"] = m.ST_PASS
    res["status"][[3, 40, 69]] = m.ST_FAIL_ASSERT
    res["status"][5] = m.ST_RUNNING
    res["ev_pop"], res["ev_fire"], res["draws"] = 10, 20, 30
    c, bm = D.summary_from_results(res, 1000)
    assert list(c[:6]) == [700, 1400, 2100, 66, 3, 1] and int(c[6]) == 1003 and int(c[7]) == 70
    assert bm[0] == (1 << 3) and bm[1] == (1 << 8) and bm[2] == (1 << 5)


def test_cpp_host_header_compiles_and_fails_loudly_without_device():
    """The C++ host mirror builds against the C-ABI; without a GPU it must refuse to run."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "cpp", "host_api_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-L" + os.path.join(ROOT, "madsim_b200"),
                    "-lmsim_b200", "-Wl,-rpath," + os.path.join(ROOT, "madsim_b200")], check=True)
    if not os.path.exists("/dev/nvidia0"):
        p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert p.returncode == 2 and "no CPU fallback" in p.stdout
