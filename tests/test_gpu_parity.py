"""-m gpu: the parity tests proper.  Everything goes through the C-ABI of libmsim_b200.so
(madsim_b200.Batch is a ctypes veneer) and is compared with the oracle on the same seeds, with
the committed golden vectors, and — at BASELINE.json's full sizes — through size-independent
properties (determinism, shard/offset independence, trace ↔ hash consistency).
Integer work: the bar is bit-exact."""
import numpy as np
import pytest

import madsim_b200 as m
from madsim_b200 import dist as D

pytestmark = pytest.mark.gpu

FIELDS = ("trace_hash", "elapsed_ns", "draws", "ev_pop", "ev_fire", "status", "fail_code", "user", "peak_timers",
          "peak_tasks")


def run_gpu(cfg, wl, base, n, S=0, L=0, max_events=0):
    with m.Batch(cfg, 0) as b:
        b.load(wl).set_geometry(S, L).run(base, n, max_events)
        launches = 1
        while max_events:
            res, bm, st = b.results()
            if st["n_unfinished"] == 0:
                break
            b.resume(max_events)
            launches += 1
            assert launches < 100000
        res, bm, st = b.results()
    return res, bm, st


def host_threads():
    """host threads this process may use (affinity and cgroup quota)"""
    import os

    n = len(os.sched_getaffinity(0))
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return n


def assert_same(ref, got, what=""):
    for f in FIELDS:
        if not np.array_equal(ref[f], got[f]):
            bad = int(np.nonzero((ref[f] != got[f]).reshape(len(ref), -1).any(axis=1))[0][0])
            pytest.fail(f"{what}: {f} differs at index {bad}: oracle {ref[f][bad]} device {got[f][bad]}")


def make_wl(d):
    d = dict(d)
    return getattr(m.Workload, d.pop("kind"))(**d)


def test_device_present():
    assert m.device_count() >= 1


def test_golden_vectors_on_device(golden):
    for name, case in golden["cases"].items():
        cfg, wl = m.Config(**case["config"]), make_wl(case["workload"])
        for r in case["rows"]:
            got, _, _ = run_gpu(cfg, wl, r[0], 1)
            g = got[0]
            assert (int(g["trace_hash"]), int(g["elapsed_ns"]), int(g["draws"]), int(g["ev_pop"]), int(g["ev_fire"]),
                    int(g["status"]), int(g["fail_code"]), [int(x) for x in g["user"]]) == (
                int(r[1]), r[2], r[3], r[4], r[5], r[6], r[7], r[8]), f"{name} seed {r[0]}"


CASES = [
    ("config1-pingpong-seed0", m.Config(), m.Workload.pingpong(rounds=1000), 0, 1),  # BASELINE config 1
    ("sleep", m.Config(), m.Workload.sleep(10**9), 0, 300),
    ("pingpong", m.Config(), m.Workload.pingpong(rounds=60), 0, 500),
    ("pingpong-loss-deadlocks", m.Config(packet_loss_rate=0.1), m.Workload.pingpong(rounds=40), 100, 200),
    ("pingpong-loss-all", m.Config(packet_loss_rate=1.0), m.Workload.pingpong(rounds=3), 0, 33),
    ("pingpong-buggify-medium-latency", m.Config(send_latency_ns=(500 * m.MS, 2500 * m.MS), buggify=True),
     m.Workload.pingpong(rounds=20), 0, 100),
    ("pingpong-small-carry-latency", m.Config(send_latency_ns=(1500 * m.MS, 2200 * m.MS)),
     m.Workload.pingpong(rounds=20), 0, 64),
    ("pingpong-timelimit", m.Config(time_limit_ns=2 * m.SEC), m.Workload.pingpong(rounds=1000), 0, 40),
    ("pingpong-zero-rounds", m.Config(), m.Workload.pingpong(rounds=0), 0, 9),
    ("raft", m.Config(), m.Workload.raft(duration_ms=4000), 0, 300),
    ("raft-faults", m.Config(), m.Workload.raft(duration_ms=5000, faults=7), 1000, 300),
    ("raft-bug", m.Config(), m.Workload.raft(duration_ms=5000, bug=1), 0, 300),
    ("raft-faults-bug-loss", m.Config(packet_loss_rate=0.05), m.Workload.raft(duration_ms=4000, faults=7, bug=1), 7, 200),
    ("raft-3node", m.Config(), m.Workload.raft(n_nodes=3, duration_ms=3000, faults=5), 0, 100),
    ("raft-1node", m.Config(), m.Workload.raft(n_nodes=1, duration_ms=1500), 0, 10),
    ("raft-fast-timeouts", m.Config(), m.Workload.raft(duration_ms=1500, rpc_timeout_ms=25, heartbeat_ms=30,
                                                        election_ms=(80, 160), faults=4), 0, 100),
    ("raft-huge-seeds", m.Config(), m.Workload.raft(duration_ms=1500, faults=7), 2**64 - 50, 50),
    ("stream-reliable-channel", m.Config(), m.Workload.stream(faults=0), 0, 300),
    ("stream-partitions", m.Config(), m.Workload.stream(faults=1), 0, 1000),
    ("stream-partitions-loss20", m.Config(packet_loss_rate=0.2), m.Workload.stream(messages=500, faults=1), 0, 300),
    ("stream-unreachable-timelimit", m.Config(packet_loss_rate=1.0, time_limit_ns=5 * m.SEC),
     m.Workload.stream(messages=3, faults=0), 0, 40),
    ("etcd-kv", m.Config(), m.Workload.etcd(), 0, 500),
    ("etcd-kv-loss30", m.Config(packet_loss_rate=0.3), m.Workload.etcd(), 0, 500),
    ("etcd-timeouts-3clients", m.Config(packet_loss_rate=0.1), m.Workload.etcd(clients=3, ops=42, timeout_rate=0.15), 50, 300),
    ("etcd-stale-read-bug", m.Config(packet_loss_rate=0.05), m.Workload.etcd(bug=1), 0, 300),
    ("etcd-reference-no-lease-panic", m.Config(), m.Workload.etcd(bug=2), 0, 300),
    ("etcd-lease-admin-timeouts", m.Config(packet_loss_rate=0.1), m.Workload.etcd(lease_admin=True, timeout_rate=0.1), 0, 400),
    ("etcd-lease-admin-3clients-bugs", m.Config(), m.Workload.etcd(lease_admin=True, clients=3, ops=30, bug=3, lease_ttl_s=2), 0, 300),
    ("etcd-always-timeout-buggify", m.Config(buggify=True), m.Workload.etcd(ops=3, timeout_rate=1.0), 0, 64),
    ("election-reference-test", m.Config(), m.Workload.election(), 0, 500),
    ("election-hold20-slow-links", m.Config(send_latency_ns=(1 * m.MS, 300 * m.MS)), m.Workload.election(hold_s=20), 10, 300),
    ("election-server-timeouts", m.Config(), m.Workload.election(timeout_rate=0.1), 0, 500),
    ("election-timeouts-loss-buggify", m.Config(packet_loss_rate=0.05, buggify=True), m.Workload.election(timeout_rate=0.3, hold_s=40), 0, 300),
    ("election-always-timeout", m.Config(), m.Workload.election(timeout_rate=1.0), 0, 40),
    ("tonic-kill-restart-chaos", m.Config(), m.Workload.tonic(), 0, 300),
    ("tonic-restart-only-1client", m.Config(), m.Workload.tonic(clients=1, restart_only=True), 0, 200),
    ("tonic-loss1-client-panics", m.Config(packet_loss_rate=0.01), m.Workload.tonic(), 0, 200),
    ("tonic-server-streaming-mix", m.Config(), m.Workload.tonic(streaming=True), 0, 400),
    ("tonic-server-streaming-slow-links", m.Config(send_latency_ns=(1 * m.MS, 900 * m.MS)), m.Workload.tonic(streaming=True, think_ns=0), 0, 300),
    ("tonic-server-crash-streaming", m.Config(), m.Workload.tonic(streaming=True, server_crash=True), 0, 400),
    ("tonic-server-crash-buggify", m.Config(buggify=True), m.Workload.tonic(server_crash=True, rounds=8, think_ns=0), 0, 200),
    ("tonic-request-timeouts", m.Config(), m.Workload.tonic(timeouts=True), 0, 200),
    ("tonic-request-timeouts-everything", m.Config(buggify=True), m.Workload.tonic(streaming=True, timeouts=True, server_crash=True, rounds=8), 0, 200),
    ("tonic-buggify-slow-server", m.Config(buggify=True), m.Workload.tonic(clients=3, rounds=5), 0, 200),
    ("tonic-interceptors-unknown-service", m.Config(), m.Workload.tonic(interceptors=True), 0, 300),
    ("tonic-interceptors-server-crash-buggify", m.Config(buggify=True), m.Workload.tonic(interceptors=True, server_crash=True, rounds=8, think_ns=0), 0, 200),
    ("tonic-16-rounds-slow-links", m.Config(send_latency_ns=(1 * m.MS, 400 * m.MS)), m.Workload.tonic(rounds=16, think_ns=0), 0, 200),
    ("signal-interval-burst-ctrlc-caught", m.Config(), m.Workload.signal(behavior=0, rounds=12), 0, 500),
    ("signal-interval-delay-ctrlc-caught", m.Config(), m.Workload.signal(behavior=1, rounds=12), 0, 500),
    ("signal-interval-skip-ctrlc-caught", m.Config(), m.Workload.signal(behavior=2, rounds=12), 0, 500),
    ("signal-interval-skip-ctrlc-kills", m.Config(), m.Workload.signal(behavior=2, rounds=12, catch=False), 0, 500),
    ("signal-interval-1ms-period", m.Config(), m.Workload.signal(behavior=0, rounds=6, period_ms=1), 0, 300),
    ("chaos-kill-restart-pause", m.Config(), m.Workload.chaos(), 0, 500),
    ("chaos-restart-on-panic", m.Config(), m.Workload.chaos(rounds=30, period_ms=700, panics=3), 0, 500),
    ("chaos-always-panicking", m.Config(), m.Workload.chaos(rounds=4, period_ms=3000, panics=50), 0, 200),
    ("chaos-panic-matching-all", m.Config(), m.Workload.chaos(rounds=10, period_ms=700, panics=3, matching=3), 0, 300),
    ("chaos-panic-matching-third-unmatched", m.Config(), m.Workload.chaos(rounds=10, period_ms=700, panics=3, matching=2), 0, 200),
]


@pytest.mark.parametrize("name,cfg,wl,base,n", CASES, ids=[c[0] for c in CASES])
def test_device_matches_oracle(oracle, name, cfg, wl, base, n):
    got, _, st = run_gpu(cfg, wl, base, n)
    ref = oracle.run_many(cfg, wl, base, n)
    assert_same(ref, got, name)
    assert st["n_unfinished"] == 0


def test_config2_full_size_4096_seeds_bit_exact(oracle):
    # BASELINE config 2: 5-node election, 4096 seeds on one B200 — every per-seed trace hash
    cfg, wl = m.Config(), m.Workload.raft()
    got, bm, st = run_gpu(cfg, wl, 0, 4096)
    ref = oracle.run_many(cfg, wl, 0, 4096)
    assert_same(ref, got, "config 2")
    assert st["n_pass"] == 4096 and st["n_fail"] == 0 and not bm.any()
    assert (got["user"][:, 0] >= 1).all(), "every fault-free run ends with a leader"


def test_config3_one_gpu_share_8192_seeds_bit_exact(oracle):
    # BASELINE config 3 shards 65536 seeds over 8 GPUs: one GPU's share, partitions + clock jumps
    cfg, wl = m.Config(), m.Workload.raft(faults=7)
    got, _, st = run_gpu(cfg, wl, 8192, 8192)
    ref = oracle.run_many(cfg, wl, 8192, 8192)
    assert_same(ref, got, "config 3 shard")


def test_config3_full_65536_seeds_properties():
    """Size-independent properties at the full config-3 size on ONE GPU: determinism (same seeds
    → same hashes on a different geometry), shard independence (8 shards == 1 run), and safety
    (no correct-Raft seed ever shows two leaders in a term)."""
    cfg, wl = m.Config(), m.Workload.raft(faults=7)
    full, bm, st = run_gpu(cfg, wl, 0, 65536)
    assert st["n_unfinished"] == 0
    assert (full["status"] == m.ST_FAIL_CAPACITY).sum() == 0, "no bounded table may overflow at the BASELINE size"
    assert (full["status"] == m.ST_PASS).all()
    if host_threads() >= 16:  # the whole configuration against the oracle (9.6e8 events: ~10 s of host time)
        import oracle_py

        assert_same(oracle_py.run_many(cfg, wl, 0, 65536), full, "config 3, all 65536 seeds")
    again, _, _ = run_gpu(cfg, wl, 0, 65536, S=14, L=2)
    assert np.array_equal(full, again)
    parts = [run_gpu(cfg, wl, r * 8192, 8192)[0] for r in range(8)]
    assert np.array_equal(full, np.concatenate(parts))
    assert len(np.unique(full["trace_hash"])) == 65536, "different seed ⇒ different trace (task/mod.rs:1018-1041)"


def test_config4_etcd_16384_seeds_packet_drop_sweep(oracle):
    """BASELINE config 4: madsim-etcd-client KV linearizability, 3 nodes, 16384 seeds, packet-drop
    sweep 0-30 %.  Every rate: a 512-seed window bit-exact against the oracle; the whole batch
    finishes, every history is linearizable (the reliable channel turns loss into delay, refused
    connects are errors the client records), and loss only ever lowers the number of applied ops."""
    wl = m.Workload.etcd()
    prev_ok = None
    for i, rate in enumerate((0.0, 0.05, 0.10, 0.15, 0.20, 0.25, 0.30)):
        cfg = m.Config(packet_loss_rate=rate)
        got, bm, st = run_gpu(cfg, wl, 0, 16384)
        assert st["n_unfinished"] == 0 and st["n_pass"] == 16384 and not bm.any(), rate
        # the WHOLE batch against the oracle where the host can afford it (16384 etcd seeds are ~1.4e7 events: a
        # second on >= 16 threads), else a 512-seed window that moves with the rate
        if host_threads() >= 16:
            lo, k = 0, 16384
        else:
            lo, k = 2048 * i, 512
        ref = oracle.run_many(cfg, wl, lo, k)
        assert_same(ref, got[lo:lo + k], f"config 4 loss {rate}")
        ok_ops = int(got["user"][:, 0].sum())
        assert (got["user"][:, 0] + got["user"][:, 3] >= 64).all()  # every op ended as ok or error (+ lease grants)
        assert prev_ok is None or ok_ops < prev_ok
        prev_ok = ok_ops


def test_config4_seeds_to_first_bug(oracle):
    # an injected stale-read bug in the server: the first failing seed is the oracle's
    cfg, wl = m.Config(packet_loss_rate=0.1), m.Workload.etcd(bug=1)
    got, bm, st = run_gpu(cfg, wl, 0, 16384)
    assert st["n_fail"] > 0 and (got["fail_code"][got["status"] == m.ST_FAIL_ASSERT] == 61).all()
    first = int(st["first_fail_seed"])
    ref = oracle.run_many(cfg, wl, 0, first + 1)
    assert (ref["status"][:first] == m.ST_PASS).all() and ref["status"][first] == m.ST_FAIL_ASSERT


def test_config5_tonic_32768_seeds_kill_restart_chaos(oracle):
    """BASELINE config 5: tonic-example Greeter under madsim-tonic, 7 nodes, 32768 seeds, kill/restart
    chaos.  A 384-seed window bit-exact against the oracle; at full size: every seed passes (no reply
    ever mismatches, the server survives every client crash), every client start is accounted for
    (6 builds + one per restart), and the batch is deterministic across launch geometries."""
    cfg, wl = m.Config(), m.Workload.tonic()
    got, bm, st = run_gpu(cfg, wl, 0, 32768)
    assert st["n_unfinished"] == 0 and st["n_pass"] == 32768 and not bm.any()
    # against the oracle: ALL 32768 seeds (7.4e8 events, ~10 s) where the host has >= 16 threads, else a 384-seed window
    lo, k = (0, 32768) if host_threads() >= 16 else (20000, 384)
    ref = oracle.run_many(cfg, wl, lo, k)
    assert_same(ref, got[lo:lo + k], "config 5")
    starts, replies, handled = got["user"][:, 1], got["user"][:, 0], got["user"][:, 2]
    assert (starts >= 6).all() and (starts <= 16).all() and (handled >= replies).all() and (replies > 0).all()
    again, _, _ = run_gpu(cfg, wl, 0, 32768, S=9, L=2)
    assert np.array_equal(got, again)


@pytest.mark.parametrize("wl", [m.Workload.etcd(timeout_rate=0.1), m.Workload.tonic(rounds=4), m.Workload.signal(behavior=2),
                                m.Workload.stream(messages=60), m.Workload.election(timeout_rate=0.05)],
                         ids=["etcd", "tonic", "signal", "stream", "election"])
def test_burst_checkpoint_covers_channel_and_signal_state(wl):
    # static (burst) launches through the HBM checkpoint == the dynamic run-to-completion launch
    cfg = m.Config(packet_loss_rate=0.05) if wl.workload != m.WL_TONIC else m.Config()
    ref, _, _ = run_gpu(cfg, wl, 0, 600)
    for k in (97, 2000):
        got, _, st = run_gpu(cfg, wl, 0, 600, max_events=k)
        assert st["n_unfinished"] == 0
        assert np.array_equal(ref, got)


def test_burst_checkpoint_resume_equals_run_to_completion():
    # bursts go through the HBM checkpoint: smem → TMA bulk store → TMA bulk load → smem
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=3000, faults=7)
    ref, _, _ = run_gpu(cfg, wl, 0, 500)
    for k in (1000, 4093):
        got, _, st = run_gpu(cfg, wl, 0, 500, max_events=k)
        assert np.array_equal(ref, got), f"burst of {k} events"
        assert st["launches"] >= 1
    got, _, _ = run_gpu(cfg, m.Workload.pingpong(rounds=30), 0, 100, max_events=17)
    ref2, _, _ = run_gpu(cfg, m.Workload.pingpong(rounds=30), 0, 100)
    assert np.array_equal(ref2, got)


@pytest.mark.parametrize("S,L", [(1, 1), (7, 1), (28, 1), (28, 4), (28, 32), (14, 2), (5, 3), (0, 0)])
def test_geometry_invariance(S, L):
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=2500, faults=7)
    ref, _, _ = run_gpu(cfg, wl, 0, 333)
    got, _, st = run_gpu(cfg, wl, 0, 333, S=S, L=L)
    assert np.array_equal(ref, got)


def test_device_trace_replay_equals_oracle_trace_and_hash(oracle):
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=3000, faults=7)
    with m.Batch(cfg, 0) as b:
        b.load(wl).run(0, 64)
        res, _, _ = b.results()
        for seed in (0, 17, 63):
            tr, rec, n = b.trace(seed)
            r, tr_o, _, _ = oracle.run_seed(cfg, wl, seed, trace=True)
            assert n == len(tr_o) and oracle.first_trace_diff(tr, tr_o) is None
            assert rec["trace_hash"] == res["trace_hash"][seed] == r["trace_hash"]
            # checksum of checksums: recompute the hash from the records
            h = 0xCBF29CE484222325
            M = (1 << 64) - 1
            rotl = lambda v, k: ((v << k) | (v >> (64 - k))) & M
            for t in tr:
                kd = (int(t["kind_id"]) << 32) | int(t["draws"])
                h ^= int(t["now_ns"]) ^ rotl(int(t["aux"]), 17) ^ rotl(kd, 41)
                h = (h * 0x9E3779B97F4A7C15) & M
                h ^= h >> 32
            assert h == int(rec["trace_hash"])


def test_summary_bitmap_and_first_failing_seed(oracle):
    # seeds-to-first-bug with the injected double-vote bug
    cfg, wl = m.Config(), m.Workload.raft(bug=1)
    res, bm, st = run_gpu(cfg, wl, 0, 512)
    ref = oracle.run_many(cfg, wl, 0, 512)
    c, rbm = D.summary_from_results(ref, 0)
    assert st["n_fail"] == int(c[4]) > 0 and st["first_fail_seed"] == int(c[6])
    assert np.array_equal(bm, rbm)
    assert st["ev_pop"] == int(c[0]) and st["ev_fire"] == int(c[1]) and st["draws"] == int(c[2])
    with pytest.raises(m.SimulationFailure) as e:
        m.Builder(seed=0, count=512).run(wl)
    assert e.value.seed == int(c[6]) and f"MADSIM_TEST_SEED={int(c[6])}" in str(e.value)
    # reproducing that one seed fails the same way (runtime/mod.rs:205-210)
    with pytest.raises(m.SimulationFailure):
        m.Builder(seed=e.value.seed, count=1).run(wl)


def test_builder_check_determinism_and_pass(oracle):
    # builder.rs:126-128: under the check only `seed` runs, twice, through Runtime::check_determinism
    rec, st = m.Builder(seed=5, count=64, check=True).run(m.Workload.pingpong(rounds=10))
    assert st is None and rec["status"] == m.ST_PASS
    assert rec["trace_hash"] == oracle.run_many(m.Config(), m.Workload.pingpong(rounds=10), 5, 1)["trace_hash"][0]
    res, st = m.Builder(seed=5, count=64).run(m.Workload.pingpong(rounds=10))
    assert st["n_pass"] == 64 and len(res) == 64
    # a failing seed under the check is the reference's panic_with_info
    with pytest.raises(m.SimulationFailure):
        m.check_determinism(0, m.Config(packet_loss_rate=1.0), m.Workload.pingpong(rounds=3))


DLOG_CASES = [c for c in CASES if c[0] in ("sleep", "pingpong-buggify-medium-latency", "pingpong-loss-deadlocks", "raft-faults",
                                            "stream-partitions-loss20", "etcd-timeouts-3clients", "etcd-lease-admin-timeouts",
                                            "tonic-request-timeouts-everything", "signal-interval-skip-ctrlc-kills",
                                            "chaos-restart-on-panic", "election-server-timeouts")]


@pytest.mark.parametrize("name,cfg,wl,base,n", DLOG_CASES, ids=[c[0] for c in DLOG_CASES])
def test_device_determinism_log_matches_oracle(oracle, name, cfg, wl, base, n):
    """Runtime::check_determinism's per-draw log (rand.rs:64-89) from the replay kernel: byte for byte the oracle's —
    one byte per GlobalRng::with call, (next u8 of a clone of the rng) ^ fold(elapsed) — and the oracle's check
    mode accepts it (and rejects a corrupted copy, naming a draw)."""
    with m.Batch(cfg, 0) as b:
        b.load(wl)
        for seed in (base, base + 5):
            log, rec = b.determinism_log(seed)
            ref, r = oracle.determinism_log(cfg, wl, seed, "oracle")
            assert len(log) == len(ref) and np.array_equal(log, ref), f"{name} seed {seed}"
            assert rec["trace_hash"] == r["trace_hash"] and rec["draws"] == r["draws"]
            assert oracle.check_against_log(cfg, wl, seed, log)
            small, _ = b.determinism_log(seed, cap=7)  # a short buffer is filled, the count still says how many
            assert np.array_equal(small, ref[: len(small)]) or len(ref) <= 7


def test_device_summary_sharded_layout_equals_host_restatement(oracle):
    """msim_batch_summarize_sharded: counters, the own first-fail slot and the bits at their GLOBAL position — for a
    shard that starts at a non-multiple of 32, seeds above 2^63, and an empty shard."""
    import torch

    cfg, wl = m.Config(), m.Workload.raft(duration_ms=1500, bug=1, faults=7)
    for seed_base, n_total, rank, world in ((40, 150, 1, 2), (2**64 - 200, 70, 0, 2), (0, 333, 2, 3)):
        first, count = D.shard(seed_base, n_total, rank, world)
        n64 = sum(D.sharded_layout(n_total, world))
        buf = torch.full((n64,), 0x5A5A5A5A, dtype=torch.int64, device="cuda")
        with m.Batch(cfg, 0) as b:
            b.load(wl).run(first, count)
            b.summarize_sharded_into(buf.data_ptr(), n64 * 8, rank, world, seed_base, n_total)
            res, _, _ = b.results()
        want = D.sharded_summary_from_results(res, first, rank, world, seed_base, n_total)
        assert np.array_equal(buf.cpu().numpy(), want), (seed_base, n_total, rank, world)
        assert np.array_equal(res["trace_hash"], oracle.run_many(cfg, wl, first, count)["trace_hash"])
    # a rank whose shard is empty never ran: zeros + "no failure" in its slot
    with m.Batch(cfg, 0) as b:
        b.load(wl)
        buf = torch.full((sum(D.sharded_layout(2, 3)),), 7, dtype=torch.int64, device="cuda")
        b.summarize_sharded_into(buf.data_ptr(), buf.numel() * 8, 2, 3, 41, 2)
        h = buf.cpu().numpy()
        assert h[8 + 2] == -1 and (np.delete(h, 10) == 0).all()


def test_resume_needs_a_checkpoint_of_the_current_run():
    # a burst run leaves a checkpoint; a later run-to-completion launch (other size, other program) must not let
    # msim_batch_resume reload the stale image
    from madsim_b200 import _native as N

    with m.Batch(m.Config(), 0) as b:
        b.load(m.Workload.raft(duration_ms=1000)).run(0, 64, max_events_per_seed=500)
        b.resume(500)
        b.run(0, 300)  # run to completion: no checkpoint
        ref, _, _ = b.results()
        with pytest.raises(m.MsimError) as e:
            b.resume(100)
        assert e.value.code == N.ESTATE
        b.load(m.Workload.pingpong(rounds=5))
        with pytest.raises(m.MsimError) as e:
            b.resume(100)
        assert e.value.code == N.ESTATE
        got, _, _ = b.run(0, 300).results()
        assert (got["status"] == m.ST_PASS).all()


def test_builder_jobs_shards_over_the_visible_gpus(oracle):
    # MADSIM_TEST_JOBS (builder.rs:134-150): here the GPUs of the box; one GPU = one shard, same results either way
    wl = m.Workload.raft(duration_ms=1500, faults=7)
    res, st = m.Builder(seed=11, count=257, jobs=8).run(wl)
    assert st["n_seeds"] == 257 and st["n_pass"] == 257
    assert np.array_equal(res["trace_hash"], oracle.run_many(m.Config(), wl, 11, 257)["trace_hash"])


@pytest.mark.parametrize("wl,S,L", [(m.Workload.stream(), 112, 4), (m.Workload.chaos(rounds=12, panics=2), 140, 5),
                                    (m.Workload.raft(duration_ms=2500, faults=7), 28, 7), (m.Workload.etcd(timeout_rate=0.1), 48, 2),
                                    (m.Workload.tonic(rounds=4), 18, 3), (m.Workload.signal(behavior=1), 100, 32),
                                    (m.Workload.election(timeout_rate=0.05), 60, 4)],
                         ids=["stream", "chaos", "raft", "etcd", "tonic", "signal", "election"])
def test_lockstep_multilane_kernel_matches_oracle(oracle, wl, S, L):
    # msim_step_kernel_ml: a warp's lanes walk the phases of the event step together (votes between the phases);
    # per seed the result must not depend on how lanes are grouped
    cfg = m.Config(packet_loss_rate=0.02)
    got, _, st = run_gpu(cfg, wl, 3, 700, S=S, L=L | m.GEOMETRY_LOCKSTEP)
    assert_same(oracle.run_many(cfg, wl, 3, 700), got, "lock-step kernel")
    assert st["lanes_per_warp"] == L


def test_capacity_overflow_is_reported_not_silent():
    cfg = m.Config()
    wl = m.Workload.raft(duration_ms=3000, rpc_timeout_ms=3, heartbeat_ms=20, election_ms=(40, 80), faults=4)
    res, bm, st = run_gpu(cfg, wl, 0, 32)
    # MSIM_CAP_MAILBOX (5) or MSIM_CAP_MSGS (4): the leaked responses exhaust one of the two tables first — of the
    # default variant AND of the large-capacity one every seed was escalated to (3 s of leaking)
    assert (res["status"] == m.ST_FAIL_CAPACITY).all() and np.isin(res["fail_code"], [4, 5]).all()
    assert st["n_fail"] == 32 and bm[0] == 0xFFFFFFFF and st["escalated"] == 32


LEAKY_300MS = dict(kind="raft", duration_ms=300, rpc_timeout_ms=3, heartbeat_ms=20, election_ms=(40, 80), faults=4)


def test_capacity_escalation_reruns_overflowed_seeds(oracle):
    # 300 ms of the leaking configuration: about half the seeds overflow the default mailbox / message tables and none
    # the large-capacity variant's.  Escalation off: those seeds say FAIL_CAPACITY.  On (the default): they run again
    # on the large tables and EVERY seed of the batch agrees with the oracle; nothing else changes.
    cfg, wl, n = m.Config(), make_wl(LEAKY_300MS), 96
    ref = oracle.run_many(cfg, wl, 5, n)
    with m.Batch(cfg, 0) as b:
        b.load(wl).set_capacity_escalation(False).run(5, n)
        off, _, st_off = b.results()
    over = off["status"] == m.ST_FAIL_CAPACITY
    assert 0 < over.sum() < n and st_off["escalated"] == 0
    assert_same(ref[~over], off[~over], "escalation off, seeds the default tables hold")
    got, bm, st = run_gpu(cfg, wl, 5, n)
    assert st["escalated"] == int(over.sum()) and st["launches"] == 2
    assert_same(ref, got, "escalation on")
    assert st["n_pass"] == n and st["n_fail"] == 0 and not bm.any()
    assert st["ev_pop"] == int(ref["ev_pop"].sum()) and st["draws"] == int(ref["draws"].sum())


@pytest.mark.parametrize("name,cfg,wl,n", [
    ("sleep", m.Config(), m.Workload.sleep(), 64), ("pingpong", m.Config(packet_loss_rate=0.05), m.Workload.pingpong(rounds=60), 300),
    ("raft", m.Config(), m.Workload.raft(duration_ms=3000, faults=7), 200), ("chaos", m.Config(), m.Workload.chaos(rounds=10, panics=2), 300),
    ("stream", m.Config(packet_loss_rate=0.2), m.Workload.stream(), 300), ("etcd", m.Config(packet_loss_rate=0.1), m.Workload.etcd(timeout_rate=0.1), 200),
    ("tonic", m.Config(), m.Workload.tonic(), 120), ("tonic-streaming", m.Config(), m.Workload.tonic(streaming=True, server_crash=True), 60),
    ("tonic-timeouts", m.Config(), m.Workload.tonic(timeouts=True), 24), ("signal", m.Config(), m.Workload.signal(), 300),
    ("election", m.Config(), m.Workload.election(timeout_rate=0.05), 200)])
def test_large_capacity_variant_of_every_program_on_the_device(oracle, name, cfg, wl, n):
    # mode 2: the whole batch on BigWL<program> (the kernels capacity escalation launches) — same records as the oracle
    with m.Batch(cfg, 0) as b:
        b.load(wl).set_capacity_escalation(2).run(11, n)
        got, _, st = b.results()
        small = m.state_bytes(wl.workload)
    assert_same(oracle.run_many(cfg, wl, 11, n), got, f"large variant of {name}")
    assert st["escalated"] == 0 and st["smem_bytes"] > small * st["seeds_per_cta"]


def test_capacity_escalation_before_the_device_summary(oracle):
    # the async path of the multi-GPU sweep: run_async, then the device-side summary without reading results first
    import torch

    cfg, wl, n = m.Config(), make_wl(LEAKY_300MS), 64
    ref = oracle.run_many(cfg, wl, 0, n)
    with m.Batch(cfg, 0) as b:
        b.load(wl).run(0, n, sync=False)
        out = torch.zeros(sum(D.sharded_layout(n, 1)), dtype=torch.int64, device="cuda:0")
        b.summarize_sharded_into(out.data_ptr(), out.numel() * 8, 0, 1, 0, n)
        b.sync()
        torch.cuda.synchronize()
        s, bitmap = D.finish_summary(out.cpu().numpy(), n, 1)
        res, _, st = b.results()
    assert st["escalated"] > 0 and not bitmap.any()
    assert s["n_pass"] == n and s["n_fail"] == 0 and s["ev_pop"] == int(ref["ev_pop"].sum())
    assert_same(ref, res, "async run + sharded summary")


def test_replay_of_an_overflowed_seed_comes_from_the_large_variant(oracle):
    cfg, wl = m.Config(), make_wl(LEAKY_300MS)
    with m.Batch(cfg, 0) as b:
        b.load(wl).set_capacity_escalation(False).run(5, 96)
        off, _, _ = b.results()
        seed = 5 + int(np.nonzero(off["status"] == m.ST_FAIL_CAPACITY)[0][0])
        tr_off, rec_off, _ = b.trace(seed)
        assert rec_off["status"] == m.ST_FAIL_CAPACITY
        b.set_capacity_escalation(True)
        tr, rec, n_tr = b.trace(seed)
        log, rec_l = b.determinism_log(seed)
    res_o, tr_o, _, _ = oracle.run_seed(cfg, wl, seed, trace=True)
    assert rec["status"] == m.ST_PASS and rec["trace_hash"] == res_o["trace_hash"]
    assert n_tr == len(tr_o) and oracle.first_trace_diff(tr_o, tr) is None
    assert len(tr_off) < len(tr_o) and np.array_equal(tr_o[: len(tr_off)], tr_off)  # the refused run is a prefix
    log_o, _ = oracle.determinism_log(cfg, wl, seed, "oracle")
    assert np.array_equal(log, log_o) and rec_l["status"] == m.ST_PASS


def test_error_behaviour():
    from madsim_b200 import _native as N

    with m.Batch(m.Config(), 0) as b:
        with pytest.raises(m.MsimError) as e:
            b.run(0, 4)  # no workload loaded
        assert e.value.code == N.ESTATE
        b.load(m.Workload.pingpong(rounds=1))
        with pytest.raises(m.MsimError) as e:
            b.results()
        assert e.value.code == N.ESTATE
        with pytest.raises(m.MsimError) as e:
            b.run(0, 0)
        assert e.value.code == N.EINVAL
        with pytest.raises(m.MsimError) as e:
            b.load(m.Workload(99))
        assert e.value.code == N.EINVAL
        with pytest.raises(m.MsimError) as e:
            b.load(m.Workload.raft(n_nodes=9))
        assert e.value.code == N.EINVAL
        b.load(m.Workload.raft()).set_geometry(1000, 1)
        with pytest.raises(m.MsimError) as e:
            b.run(0, 4096)
        assert e.value.code == N.EINVAL
    with pytest.raises(m.MsimError):
        m.Batch(m.Config(packet_loss_rate=2.0), 0).load(m.Workload.pingpong())


def test_sharded_run_single_rank_and_device_summary():
    import torch

    run = D.ShardedRun(m.Config(), m.Workload.raft(duration_ms=2000, bug=1), 0)
    out, bitmap = run.run(0, 1000).reduce()
    res, bm, st = run.batch.results()
    assert out["n_seeds"] == 1000 and out["n_fail"] == st["n_fail"] and out["ev_pop"] == st["ev_pop"]
    assert out["first_fail_seed"] == (None if st["n_fail"] == 0 else st["first_fail_seed"])
    assert np.array_equal(bitmap, bm)
    run.close()


def test_cpp_host_api_program():
    """include/msim_b200.hpp (C++ mirror of madsim::runtime::{Runtime, Builder}) driving the C-ABI:
    the reference's unit tests for the path, written as they read in Rust (tests/cpp/host_api_test.cpp)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "host_api_test")
    src = os.path.join(root, "tests", "cpp", "host_api_test.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, src, "-L" + os.path.join(root, "madsim_b200"),
                        "-lmsim_b200", "-Wl,-rpath," + os.path.join(root, "madsim_b200")], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "HOST API TEST PASSED" in p.stdout, p.stdout + p.stderr
