"""Host logic: the device engine's state machines, compiled for the host by tests/enginecheck
(TEST-ONLY — the product runs them on the GPU exclusively), diffed against the oracle.  This is
what lets the state machines be developed and regression-tested where there is no GPU."""
import numpy as np
import pytest

import madsim_b200 as m

FIELDS = ("trace_hash", "elapsed_ns", "draws", "ev_pop", "ev_fire", "status", "fail_code", "user", "peak_timers",
          "peak_tasks")


def escalate_host(oracle, cfg, wl, seed_base, got):
    """What capacity escalation does on the device (msim_b200.cu escalate): seeds that overflowed a bounded table of
    the default variant run again on the large-capacity variant (engine.h BigWL).  → number of seeds re-run"""
    over = np.nonzero(got["status"] == m.ST_FAIL_CAPACITY)[0]
    for i in over:
        got[i] = oracle.engine_host_big(cfg, wl, seed_base + int(i), 1)[0]
    return len(over)


def _assert_same(oracle, cfg, wl, seed_base, n, **kw):
    ref = oracle.run_many(cfg, wl, seed_base, n)
    got, _ = oracle.engine_host(cfg, wl, seed_base, n, **kw)
    escalate_host(oracle, cfg, wl, seed_base, got)
    for f in FIELDS:
        if not np.array_equal(ref[f], got[f]):
            bad = int(np.nonzero((ref[f] != got[f]).reshape(n, -1).any(axis=1))[0][0])
            _, tr_o, _, _ = oracle.run_seed(cfg, wl, seed_base + bad, trace=True)
            _, tr_e = oracle.engine_host(cfg, wl, seed_base + bad, 1, trace_seed=seed_base + bad)
            d = oracle.first_trace_diff(tr_o, tr_e)
            pytest.fail(f"{f} differs at seed {seed_base + bad}: oracle {ref[f][bad]} engine {got[f][bad]}; "
                        f"first trace divergence at record {d}")


CASES = [
    ("sleep", m.Config(), m.Workload.sleep(10**9), 0, 64),
    ("pingpong", m.Config(), m.Workload.pingpong(rounds=40), 0, 64),
    ("pingpong-loss", m.Config(packet_loss_rate=0.1), m.Workload.pingpong(rounds=40), 100, 64),  # deadlocks
    ("pingpong-lossall", m.Config(packet_loss_rate=1.0), m.Workload.pingpong(rounds=3), 0, 8),
    ("pingpong-buggify-medium", m.Config(send_latency_ns=(500 * m.MS, 2500 * m.MS), buggify=True),
     m.Workload.pingpong(rounds=20), 0, 32),
    ("pingpong-small-carry", m.Config(send_latency_ns=(1500 * m.MS, 2200 * m.MS)), m.Workload.pingpong(rounds=20), 0, 16),
    ("pingpong-timelimit", m.Config(time_limit_ns=2 * m.SEC), m.Workload.pingpong(rounds=1000), 0, 8),
    ("pingpong-zero-rounds", m.Config(), m.Workload.pingpong(rounds=0), 0, 8),
    ("raft", m.Config(), m.Workload.raft(duration_ms=3000), 0, 48),
    ("raft-faults", m.Config(), m.Workload.raft(duration_ms=4000, faults=7), 1000, 48),
    ("raft-bug", m.Config(), m.Workload.raft(duration_ms=4000, bug=1), 0, 120),
    ("raft-faults-bug-loss", m.Config(packet_loss_rate=0.05), m.Workload.raft(duration_ms=4000, faults=7, bug=1), 7, 48),
    ("raft-3node", m.Config(), m.Workload.raft(n_nodes=3, duration_ms=3000, faults=5), 0, 32),
    ("raft-1node", m.Config(), m.Workload.raft(n_nodes=1, duration_ms=1500), 0, 8),
    ("raft-fast-timeouts", m.Config(), m.Workload.raft(duration_ms=1500, rpc_timeout_ms=25, heartbeat_ms=30,
                                                        election_ms=(80, 160), faults=4), 0, 32),
    ("raft-zero-duration", m.Config(), m.Workload.raft(duration_ms=0), 0, 4),
    # reliable channels: connect1 / accept1 / Sender / Receiver (net/mod.rs:337-430), SURVEY §8(f) row 1
    ("stream", m.Config(), m.Workload.stream(faults=0), 0, 100),
    ("stream-partitions", m.Config(), m.Workload.stream(faults=1), 0, 200),
    ("stream-partitions-loss20", m.Config(packet_loss_rate=0.2), m.Workload.stream(messages=500, faults=1), 0, 100),
    ("stream-unreachable-timelimit", m.Config(packet_loss_rate=1.0, time_limit_ns=5 * m.SEC),
     m.Workload.stream(messages=3, faults=0), 0, 16),
    # madsim-etcd-client SimServer + one-connection-per-request clients (BASELINE config 4), SURVEY §8(f) row 3
    ("etcd", m.Config(), m.Workload.etcd(), 0, 64),
    ("etcd-loss30", m.Config(packet_loss_rate=0.3), m.Workload.etcd(), 0, 64),
    ("etcd-timeouts-3clients", m.Config(packet_loss_rate=0.1), m.Workload.etcd(clients=3, ops=42, timeout_rate=0.15), 50, 48),
    ("etcd-stale-read-bug", m.Config(packet_loss_rate=0.05), m.Workload.etcd(bug=1), 0, 64),
    ("etcd-reference-no-lease-panic", m.Config(), m.Workload.etcd(bug=2), 0, 64),
    ("etcd-lease-admin-timeouts", m.Config(packet_loss_rate=0.1), m.Workload.etcd(lease_admin=True, timeout_rate=0.1), 0, 96),
    ("etcd-lease-admin-3clients-bugs", m.Config(), m.Workload.etcd(lease_admin=True, clients=3, ops=30, bug=3, lease_ttl_s=2), 0, 64),
    ("etcd-always-timeout", m.Config(), m.Workload.etcd(ops=3, timeout_rate=1.0), 0, 16),
    # madsim-etcd-client election API over the watcher: the reference's election() test (tests/test.rs:129-232), SURVEY §8(f) row 3
    ("election", m.Config(), m.Workload.election(), 0, 64),
    ("election-hold20-slow-links", m.Config(send_latency_ns=(1 * m.MS, 300 * m.MS)), m.Workload.election(hold_s=20), 10, 64),
    ("election-server-timeouts", m.Config(), m.Workload.election(timeout_rate=0.1), 0, 96),
    ("election-timeouts-loss-buggify", m.Config(packet_loss_rate=0.05, buggify=True), m.Workload.election(timeout_rate=0.3, hold_s=40), 0, 96),
    ("election-always-timeout", m.Config(), m.Workload.election(timeout_rate=1.0), 0, 16),
    # tonic-example Greeter under madsim-tonic + kill/restart chaos (BASELINE config 5)
    ("tonic", m.Config(), m.Workload.tonic(), 0, 32),
    ("tonic-restart-only-1client", m.Config(), m.Workload.tonic(clients=1, restart_only=True), 0, 32),
    ("tonic-loss1-client-panics", m.Config(packet_loss_rate=0.01), m.Workload.tonic(), 0, 32),
    ("tonic-server-streaming-mix", m.Config(), m.Workload.tonic(streaming=True), 0, 64),
    ("tonic-server-streaming-restart-only", m.Config(), m.Workload.tonic(streaming=True, restart_only=True, clients=2, rounds=16), 0, 64),
    ("tonic-server-crash-streaming", m.Config(), m.Workload.tonic(streaming=True, server_crash=True), 0, 96),
    ("tonic-server-crash-buggify", m.Config(buggify=True), m.Workload.tonic(server_crash=True, rounds=8, think_ns=0), 0, 64),
    ("tonic-request-timeouts", m.Config(), m.Workload.tonic(timeouts=True), 0, 64),
    ("tonic-request-timeouts-everything", m.Config(buggify=True), m.Workload.tonic(streaming=True, timeouts=True, server_crash=True, rounds=8), 0, 64),
    ("tonic-buggify-slow-server", m.Config(buggify=True), m.Workload.tonic(clients=3, rounds=5), 0, 160),
    ("tonic-interceptors-unknown-service", m.Config(), m.Workload.tonic(interceptors=True), 0, 64),
    ("tonic-interceptors-server-crash-buggify", m.Config(buggify=True), m.Workload.tonic(interceptors=True, server_crash=True, rounds=8, think_ns=0), 0, 64),
    # time::interval + MissedTickBehavior, signal::ctrl_c, Handle::send_ctrl_c
    ("signal-burst-catch", m.Config(), m.Workload.signal(behavior=0, rounds=12), 0, 100),
    ("signal-delay-catch", m.Config(), m.Workload.signal(behavior=1, rounds=12), 0, 100),
    ("signal-skip-catch", m.Config(), m.Workload.signal(behavior=2, rounds=12), 0, 100),
    ("signal-skip-ctrlc-kills", m.Config(), m.Workload.signal(behavior=2, rounds=12, catch=False), 0, 100),
    # Handle::{kill,restart,pause,resume} + restart_on_panic (task/mod.rs:356-424,289-317)
    ("chaos", m.Config(), m.Workload.chaos(), 0, 200),
    ("chaos-long-panics", m.Config(), m.Workload.chaos(rounds=30, period_ms=700, panics=3), 0, 200),
    ("chaos-always-panicking", m.Config(), m.Workload.chaos(rounds=4, period_ms=3000, panics=50), 0, 100),
    # restart_on_panic_matching (task/mod.rs:298-303,965-983): every panic matched / the third one is not ("should panic with 2")
    ("chaos-panic-matching-all", m.Config(), m.Workload.chaos(rounds=10, period_ms=700, panics=3, matching=3), 0, 100),
    ("chaos-panic-matching-third-unmatched", m.Config(), m.Workload.chaos(rounds=10, period_ms=700, panics=3, matching=2), 0, 60),
]


@pytest.mark.parametrize("name,cfg,wl,base,n", CASES, ids=[c[0] for c in CASES])
def test_engine_matches_oracle(oracle, name, cfg, wl, base, n):
    _assert_same(oracle, cfg, wl, base, n, seeds_per_block=5)


@pytest.mark.parametrize("name,cfg,wl,base,n", CASES, ids=[c[0] for c in CASES])
def test_multilane_phase_loop_matches_oracle(oracle, name, cfg, wl, base, n):
    # the multi-lane batch kernel's lock-step phase loop (host emulation, 3 and 8 lanes; the seed queue hands lanes new
    # seeds as old ones finish): per seed it must execute exactly Engine::step()'s sequence
    ref = oracle.run_many(cfg, wl, base, n)
    for lanes in (3, 8):
        got = oracle.engine_host_ml(cfg, wl, base, n, lanes=lanes)
        escalate_host(oracle, cfg, wl, base, got)
        for f in FIELDS:
            assert np.array_equal(ref[f], got[f]), f"{name}: {f} differs with {lanes} lanes"


DLOG_CASES = [c for c in CASES if c[0] in ("sleep", "pingpong-buggify-medium", "pingpong-loss", "raft-faults", "raft-faults-bug-loss",
                                            "stream-partitions-loss20", "etcd-timeouts-3clients", "etcd-lease-admin-timeouts",
                                            "tonic-request-timeouts-everything", "signal-skip-ctrlc-kills", "chaos-long-panics",
                                            "election-server-timeouts")]


@pytest.mark.parametrize("name,cfg,wl,base,n", DLOG_CASES, ids=[c[0] for c in DLOG_CASES])
def test_determinism_log_matches_oracle(oracle, name, cfg, wl, base, n):
    # Runtime::check_determinism's log (rand.rs:64-89): one byte per GlobalRng::with call.  The replay kernel's hook
    # (Engine::log_point) must put a byte exactly where the reference does — every gen_range / gen_bool / random()
    # of the runtime AND of the program — and the oracle's check mode must accept the engine's log
    for seed in (base, base + 3):
        a, ra = oracle.determinism_log(cfg, wl, seed, "oracle")
        b, rb = oracle.determinism_log(cfg, wl, seed, "engine")
        assert len(a) == len(b) and np.array_equal(a, b), f"{name} seed {seed}: logs differ"
        assert ra["trace_hash"] == rb["trace_hash"] and len(a) <= ra["draws"]
        assert oracle.check_against_log(cfg, wl, seed, b)
        if len(b) > 10 and ra["status"] == m.ST_PASS:
            flipped = b.copy()
            flipped[len(b) // 2] ^= 0x10
            assert not oracle.check_against_log(cfg, wl, seed, flipped), "a corrupted log must be detected"


def test_burst_checkpoint_equals_run_to_completion(oracle):
    # registers ↔ state-block spill every k events must not change anything
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=2000, faults=7)
    a, _ = oracle.engine_host(cfg, wl, 0, 16, seeds_per_block=4)
    for k in (1, 7, 1000):
        b, _ = oracle.engine_host(cfg, wl, 0, 16, seeds_per_block=3, max_events=k)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("wl", [m.Workload.etcd(timeout_rate=0.1), m.Workload.tonic(rounds=4), m.Workload.signal(behavior=2),
                                m.Workload.stream(messages=60), m.Workload.election(timeout_rate=0.05)],
                         ids=["etcd", "tonic", "signal", "stream", "election"])
def test_burst_checkpoint_covers_channel_and_signal_state(oracle, wl):
    # connections, accept queues, leases, ctrl-c waiters all live in the state block: k-event bursts == one run
    cfg = m.Config(packet_loss_rate=0.05) if wl.workload != m.WL_TONIC else m.Config()
    a, _ = oracle.engine_host(cfg, wl, 0, 12, seeds_per_block=4)
    assert np.array_equal(a, oracle.run_many(cfg, wl, 0, 12))
    for k in (1, 13, 500):
        b, _ = oracle.engine_host(cfg, wl, 0, 12, seeds_per_block=3, max_events=k)
        assert np.array_equal(a, b)


def test_full_trace_equal(oracle):
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=2500, faults=7)
    for seed in (0, 5, 123456789012345):
        _, tr_o, _, _ = oracle.run_seed(cfg, wl, seed, trace=True)
        _, tr_e = oracle.engine_host(cfg, wl, seed, 1, trace_seed=seed)
        assert len(tr_o) == len(tr_e) and oracle.first_trace_diff(tr_o, tr_e) is None


def test_capacity_overflow_is_reported_not_silent(oracle):
    # 20 virtual minutes of clock-jump faults accumulate stale responses / timers until a bounded
    # table overflows on SOME seed; such seeds must say FAIL_CAPACITY, everything else must match
    cfg, wl = m.Config(), m.Workload.raft(duration_ms=120_000, faults=7)
    got, _ = oracle.engine_host(cfg, wl, 0, 12, seeds_per_block=4)
    ref = oracle.run_many(cfg, wl, 0, 12)
    ok = got["status"] != m.ST_FAIL_CAPACITY
    assert np.array_equal(got["trace_hash"][ok], ref["trace_hash"][ok])
    assert set(np.unique(got["status"])) <= {m.ST_PASS, m.ST_FAIL_CAPACITY}
    assert (got["fail_code"][~ok] > 0).all()


def test_large_capacity_variant_takes_the_overflowed_seeds(oracle):
    # capacity escalation (engine.h BigWL): the seeds the default tables refuse run on tables 2-4x the size and must
    # then agree with the oracle field for field; seeds the default tables hold give the same records on both variants
    # (the case: rpc timeout below the round trip leaks every response into the mailbox, 300 ms of it)
    cfg = m.Config()
    wl = m.Workload.raft(duration_ms=300, rpc_timeout_ms=3, heartbeat_ms=20, election_ms=(40, 80), faults=4)
    got, _ = oracle.engine_host(cfg, wl, 0, 16, seeds_per_block=4)
    big = oracle.engine_host_big(cfg, wl, 0, 16)
    ref = oracle.run_many(cfg, wl, 0, 16)
    over = got["status"] == m.ST_FAIL_CAPACITY
    assert over.any() and not over.all(), "the case is meant to overflow the default tables on SOME seeds"
    for f in FIELDS:
        assert np.array_equal(big[f], ref[f]), f
        assert np.array_equal(got[f][~over], big[f][~over]), f


@pytest.mark.parametrize("name,cfg,wl,n", [
    ("pingpong", m.Config(), m.Workload.pingpong(rounds=50), 16),
    ("chaos", m.Config(), m.Workload.chaos(), 16),
    ("stream", m.Config(packet_loss_rate=0.2), m.Workload.stream(), 16),
    ("etcd", m.Config(packet_loss_rate=0.1), m.Workload.etcd(), 8),
    ("election", m.Config(), m.Workload.election(), 8),
    ("tonic", m.Config(), m.Workload.tonic(), 4),
    ("signal", m.Config(), m.Workload.signal(), 16),
    ("sleep", m.Config(), m.Workload.sleep(), 4),
])
def test_large_capacity_variant_is_the_same_program(oracle, name, cfg, wl, n):
    big = oracle.engine_host_big(cfg, wl, 7, n)
    ref = oracle.run_many(cfg, wl, 7, n)
    for f in FIELDS:
        assert np.array_equal(big[f], ref[f]), (name, f)


def test_leaky_config_overflows_loudly(oracle):
    # rpc timeout (3 ms) below the round trip: EVERY response arrives after its caller gave up and
    # stays in the mailbox forever (the reference leaks them too, endpoint.rs:335-350) → the bounded
    # device mailbox must report MSIM_CAP_MAILBOX, and the trace must agree up to that point
    cfg = m.Config()
    wl = m.Workload.raft(duration_ms=3000, rpc_timeout_ms=3, heartbeat_ms=20, election_ms=(40, 80), faults=4)
    got, tr_e = oracle.engine_host(cfg, wl, 0, 1, trace_seed=0)
    assert got["status"][0] == m.ST_FAIL_CAPACITY and got["fail_code"][0] == 5
    _, tr_o, _, _ = oracle.run_seed(cfg, wl, 0, trace=True)
    assert len(tr_e) < len(tr_o) and np.array_equal(tr_o[: len(tr_e)], tr_e)


def test_bug_mode_finds_a_failing_seed(oracle):
    # seeds-to-first-bug: the injected double-vote bug must be caught as "two leaders in one term"
    cfg, wl = m.Config(), m.Workload.raft(bug=1)
    got, _ = oracle.engine_host(cfg, wl, 0, 128, seeds_per_block=8)
    fails = np.nonzero(got["status"] == m.ST_FAIL_ASSERT)[0]
    assert len(fails) > 0 and (got["fail_code"][fails] == 1).all()
    ref = oracle.run_many(cfg, wl, 0, 128)
    assert np.array_equal(got["status"], ref["status"]) and int(fails[0]) == int(np.nonzero(ref["status"] == 2)[0][0])
