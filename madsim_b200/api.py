"""Host-side mirror of the reference's runtime interface for the hot path.

Reference (Rust)                                   here
-------------------------------------------------  ------------------------------------------
madsim::Config{net{packet_loss_rate,send_latency}}  Config            (config.rs:11-35, network.rs:66-89)
Runtime::with_seed_and_config / block_on            Batch.run         (runtime/mod.rs:53-130) — many seeds at once
Builder::from_env().run(f)                          Builder.from_env().run(workload)  (runtime/builder.rs:64-162)
panic "note: run with `MADSIM_TEST_SEED=..`"        SimulationFailure (runtime/mod.rs:205-210)
`f` (the async test body)                           Workload          (a registered program, see include/msim_b200.h)

Everything executes on the GPU through the C-ABI (libmsim_b200.so); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _native as N

SEED_RESULT_DTYPE = np.dtype(
    [
        ("trace_hash", "<u8"),
        ("elapsed_ns", "<u8"),
        ("draws", "<u8"),
        ("ev_pop", "<u4"),
        ("ev_fire", "<u4"),
        ("status", "<u4"),
        ("fail_code", "<u4"),
        ("user", "<u4", (4,)),
        ("peak_timers", "<u4"),
        ("peak_tasks", "<u4"),
    ]
)
TRACE_DTYPE = np.dtype([("now_ns", "<u8"), ("aux", "<u8"), ("kind_id", "<u4"), ("draws", "<u4")])
assert SEED_RESULT_DTYPE.itemsize == 64 and TRACE_DTYPE.itemsize == 24
MS = 1_000_000
SEC = 1_000_000_000


@dataclass
class Config:
    """madsim::Config (config.rs:11-21) + Runtime knobs (builder.rs:26-63)."""

    packet_loss_rate: float = 0.0
    send_latency_ns: tuple = (1 * MS, 10 * MS)
    time_limit_ns: int = 0
    buggify: bool = False

    @staticmethod
    def from_toml(text: str) -> "Config":
        """`impl FromStr for Config` (config.rs:28-35): the TOML of a MADSIM_TEST_CONFIG file, parsed by the library
        (msim_config_from_toml) — `[net] packet_loss_rate, send_latency = { start = {secs, nanos}, end = {..} }`."""
        c = N.MsimConfig()
        N.check(N.load().msim_config_from_toml(text.encode(), C.byref(c)))
        return Config(c.packet_loss_rate, (c.send_latency_lo_ns, c.send_latency_hi_ns), 0, False)

    def to_c(self) -> N.MsimConfig:
        return N.MsimConfig(
            float(self.packet_loss_rate),
            int(self.send_latency_ns[0]),
            int(self.send_latency_ns[1]),
            int(self.time_limit_ns),
            1 if self.buggify else 0,
            0,
        )


@dataclass
class Workload:
    """A registered program: what `f` is in `Builder::run(f)`."""

    workload: int
    n_nodes: int = 0
    p: list = field(default_factory=lambda: [0] * 8)

    @staticmethod
    def sleep(duration_ns: int = SEC) -> "Workload":
        return Workload(N.WL_SLEEP, 0, [duration_ns, 0, 0, 0, 0, 0, 0, 0])

    @staticmethod
    def pingpong(rounds: int = 1000, think_ns: int = 10 * MS) -> "Workload":
        return Workload(N.WL_PINGPONG, 2, [rounds, think_ns, 0, 0, 0, 0, 0, 0])

    @staticmethod
    def raft(
        n_nodes: int = 5,
        duration_ms: int = 10_000,
        faults: int = 0,
        bug: int = 0,
        rpc_timeout_ms: int = 40,
        heartbeat_ms: int = 50,
        election_ms: tuple = (150, 300),
    ) -> "Workload":
        return Workload(
            N.WL_RAFT,
            n_nodes,
            [duration_ms, faults, bug, rpc_timeout_ms, heartbeat_ms, election_ms[0], election_ms[1], 0],
        )

    @staticmethod
    def chaos(rounds: int = 8, period_ms: int = 2000, panics: int = 0, matching: int = 0) -> "Workload":
        """kill / restart / pause / resume supervisor schedule (task/mod.rs:858-1016); with
        `panics` > 0 the worker panics on its first starts under restart_on_panic (:939-963); with `matching` = k the node
        is built with restart_on_panic_matching("0") .. ("k-1") instead and the worker panics with its start count as
        the message (:965-983): the panic number k (if there is one) is not matched and fails the seed."""
        return Workload(N.WL_CHAOS, 1, [rounds, period_ms, panics, matching, 0, 0, 0, 0])

    @staticmethod
    def stream(messages: int = 200, think_ns: int = 5 * MS, faults: int = 1) -> "Workload":
        """reliable-channel echo (connect1/accept1/Sender/Receiver, net/mod.rs:337-430) under partitions."""
        return Workload(N.WL_STREAM, 2, [messages, think_ns, faults, 0, 0, 0, 0, 0])

    @staticmethod
    def etcd(clients: int = 2, ops: int = 32, think_ns: int = 50 * MS, timeout_rate: float = 0.0, bug: int = 0,
             lease_ttl_s: int = 1, lease_admin: bool = False) -> "Workload":
        """madsim-etcd-client SimServer + `clients` KV clients (one connection per request) with a
        linearizability check of the history (BASELINE config 4); `timeout_rate` is
        SimServer::timeout_rate (an f32 in the reference: passed as its bit pattern)."""
        import struct

        bits = struct.unpack("<I", struct.pack("<f", timeout_rate))[0]
        return Workload(N.WL_ETCD, clients + 1, [ops, think_ns, bits, bug, lease_ttl_s, 0, 0, int(lease_admin)])

    @staticmethod
    def tonic(clients: int = 6, rounds: int = 10, think_ns: int = 100 * MS, restart_only: bool = False,
              streaming: bool = False, server_crash: bool = False, timeouts: bool = False, interceptors: bool = False) -> "Workload":
        """tonic-example Greeter service under madsim-tonic (a fresh endpoint + handshake + connection per
        unary call) while a supervisor kills / restarts the client nodes (BASELINE config 5); with `streaming`
        the clients also make server-streaming (lots_of_replies, sometimes dropping the response stream) and
        client-streaming (lots_of_greetings) and bidi (bidi_hello, sometimes dropped after one reply) calls; with
        `interceptors` both sides run a with_interceptor closure that turns every second request down and some calls
        name a service the server does not have (tests/test.rs:280-363)."""
        return Workload(N.WL_TONIC, clients + 1, [rounds, think_ns, int(restart_only) | 2 * int(streaming) | 4 * int(server_crash) | 8 * int(timeouts) | 16 * int(interceptors), 0, 0, 0, 0, 0])

    @staticmethod
    def signal(rounds: int = 8, period_ms: int = 100, behavior: int = 0, catch: bool = True) -> "Workload":
        """time::interval with a MissedTickBehavior (0 Burst, 1 Delay, 2 Skip) on a worker node that a
        supervisor sends ctrl-c to / restarts (time/interval.rs, signal.rs, Handle::send_ctrl_c)."""
        return Workload(N.WL_SIGNAL, 1, [rounds, period_ms, behavior, int(catch), 0, 0, 0, 0])

    @staticmethod
    def election(hold_s: int = 30, timeout_rate: float = 0.0) -> "Workload":
        """madsim-etcd-client's election API — campaign / proclaim / leader / observe / resign over the server's
        watcher — as the reference's own `election()` test drives it (tests/test.rs:129-232): a first leader that
        holds `hold_s` seconds before it revokes its lease, a second candidate waiting in campaign(), an observer."""
        import struct

        bits = struct.unpack("<I", struct.pack("<f", timeout_rate))[0]
        return Workload(N.WL_ELECTION, 4, [hold_s, bits, 0, 0, 0, 0, 0, 0])

    def to_c(self) -> N.MsimWorkloadDesc:
        return N.MsimWorkloadDesc(self.workload, self.n_nodes, (C.c_uint64 * 8)(*[int(x) for x in self.p]))

    def name(self) -> str:
        return {0: "sleep", 1: "pingpong", 2: "raft", 3: "chaos", 4: "stream", 5: "etcd", 6: "tonic", 7: "signal", 8: "election"}.get(self.workload, "?")


class SimulationFailure(AssertionError):
    """What a panicking seed is in the reference (runtime/mod.rs:205-210)."""

    def __init__(self, seed: int, status: int, fail_code: int):
        self.seed, self.status, self.fail_code = seed, status, fail_code
        super().__init__(
            f"seed {seed}: {N.STATUS_NAMES.get(status, status)} (code {fail_code})\n"
            f"note: run with `MADSIM_TEST_SEED={seed}` environment variable to reproduce this error"
        )


class Batch:
    """Many `Runtime`s that share a Config, advanced together on one GPU."""

    def __init__(self, config: Config | None = None, device: int = 0):
        self._lib = N.load()
        self._h = C.c_void_p()
        self.config = config or Config()
        cfg = self.config.to_c()
        N.check(self._lib.msim_batch_create(C.byref(cfg), device, C.byref(self._h)))
        self.device = device
        self.n_seeds = 0
        self.seed_base = 0
        self.workload: Workload | None = None

    def close(self):
        if self._h:
            self._lib.msim_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def load(self, workload: Workload) -> "Batch":
        d = workload.to_c()
        N.check(self._lib.msim_workload_load(self._h, C.byref(d)))
        self.workload = workload
        return self

    def set_geometry(self, seeds_per_cta: int = 0, lanes_per_warp: int = 0) -> "Batch":
        N.check(self._lib.msim_batch_set_geometry(self._h, seeds_per_cta, lanes_per_warp))
        return self

    def set_capacity_escalation(self, on=True) -> "Batch":
        """Re-run seeds that overflow a bounded table (ST_FAIL_CAPACITY) on the program's large-capacity variant
        (default on; stats["escalated"] counts them).  False / 0: such seeds are reported and never re-run.
        2: every seed of a run-to-completion batch goes to the large variant straight away."""
        N.check(self._lib.msim_batch_set_capacity_escalation(self._h, int(on)))
        return self

    def set_stream(self, cuda_stream: int) -> "Batch":
        N.check(self._lib.msim_batch_set_stream(self._h, C.c_void_p(cuda_stream)))
        return self

    def run(self, seed_base: int, n_seeds: int, max_events_per_seed: int = 0, sync: bool = True) -> "Batch":
        fn = self._lib.msim_batch_run if sync else self._lib.msim_batch_run_async
        N.check(fn(self._h, seed_base, n_seeds, max_events_per_seed))
        self.seed_base, self.n_seeds = seed_base, n_seeds
        return self

    def resume(self, max_events_per_seed: int = 0) -> "Batch":
        N.check(self._lib.msim_batch_resume(self._h, max_events_per_seed))
        return self

    def sync(self) -> "Batch":
        N.check(self._lib.msim_batch_sync(self._h))
        return self

    def results(self, want_results: bool = True, want_summary: bool = True):
        """→ (results ndarray[n] of SEED_RESULT_DTYPE or None, fail_bitmap uint32[ceil(n/32)], stats dict).
        want_summary=False skips the device summary (bitmap and stats come back None): for callers that already
        hold the reduced summary of a sharded run."""
        n = self.n_seeds
        res = np.empty(n, dtype=SEED_RESULT_DTYPE) if want_results else None
        bitmap = np.zeros((n + 31) // 32, dtype=np.uint32) if want_summary else None
        st = N.MsimStats()
        N.check(
            self._lib.msim_batch_results(
                self._h,
                res.ctypes.data_as(C.c_void_p) if want_results else None,
                bitmap.ctypes.data_as(C.c_void_p) if want_summary else None,
                C.byref(st) if want_summary else None,
            )
        )
        stats = {k: getattr(st, k) for k, _ in N.MsimStats._fields_} if want_summary else None
        return res, bitmap, stats

    def trace_hash(self) -> np.ndarray:
        h = np.empty(self.n_seeds, dtype=np.uint64)
        N.check(self._lib.msim_batch_trace_hash(self._h, h.ctypes.data_as(C.c_void_p)))
        return h

    def trace(self, seed: int, cap: int = 1 << 20):
        """Full canonical trace of one seed (device replay).  → (records ndarray, result record)"""
        buf = np.empty(cap, dtype=TRACE_DTYPE)
        n = C.c_size_t(0)
        r = N.MsimSeedResult()
        N.check(self._lib.msim_batch_trace(self._h, seed, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(r)))
        rec = np.frombuffer(bytes(r), dtype=SEED_RESULT_DTYPE)[0]
        return buf[: min(n.value, cap)].copy(), rec, n.value

    def determinism_log(self, seed: int, cap: int = 1 << 22):
        """The log Runtime::check_determinism records (rand.rs:64-89): one byte per GlobalRng::with call of one seed
        (device replay).  → (uint8 ndarray, result record)"""
        buf = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        r = N.MsimSeedResult()
        N.check(self._lib.msim_batch_determinism_log(self._h, seed, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(r)))
        if n.value > cap:
            return self.determinism_log(seed, n.value)
        return buf[: n.value].copy(), np.frombuffer(bytes(r), dtype=SEED_RESULT_DTYPE)[0]

    def summarize_sharded_into(self, device_ptr: int, nbytes: int, rank: int, world: int, seed_base: int, n_total: int):
        """This shard's summary in the one-all-reduce layout (msim_batch_summarize_sharded)."""
        N.check(self._lib.msim_batch_summarize_sharded(self._h, C.c_void_p(device_ptr), nbytes, rank, world, seed_base, n_total))

    @staticmethod
    def sharded_summary_nbytes(n_total: int, world: int) -> int:
        return 8 * (8 + world) + 4 * ((n_total + 31) // 32)

    def summarize_into(self, device_ptr: int, nbytes: int):
        """Device-side summary (counters + fail bitmap) into caller-owned device memory."""
        N.check(self._lib.msim_batch_summarize_device(self._h, C.c_void_p(device_ptr), nbytes))

    def results_device(self):
        p, b = C.c_void_p(), C.c_size_t()
        N.check(self._lib.msim_batch_results_device(self._h, C.byref(p), C.byref(b)))
        return p.value, b.value

    @staticmethod
    def summary_nbytes(n_seeds: int) -> int:
        return 64 + 4 * ((n_seeds + 31) // 32)


def state_bytes(workload: int) -> int:
    return N.load().msim_workload_state_bytes(workload)


def device_count() -> int:
    n = N.load().msim_device_count()
    if n < 0:
        raise N.MsimError(n, N.load().msim_last_error().decode())
    return n


class NonDeterminism(AssertionError):
    """The panic of GlobalRng's check mode (rand.rs:79-84)."""


def check_determinism(seed: int, config: Config, workload: Workload, device: int = 0):
    """Runtime::check_determinism (runtime/mod.rs:178-202): run `seed` twice; the first run records one byte per
    GlobalRng::with call — (next u8 of a clone of the rng) ^ fold(elapsed) —, the second run must reproduce every
    byte.  Raises NonDeterminism naming the first diverging draw; a failing seed raises SimulationFailure."""
    with Batch(config, device) as b:
        b.load(workload)
        log, r1 = b.determinism_log(seed)
        if r1["status"] != N.ST_PASS:
            raise SimulationFailure(seed, int(r1["status"]), int(r1["fail_code"]))
        chk, r2 = b.determinism_log(seed)
    n = min(len(log), len(chk))
    bad = np.nonzero(log[:n] != chk[:n])[0]
    if len(bad) or len(log) != len(chk):
        i = int(bad[0]) if len(bad) else n
        raise NonDeterminism(f"non-determinism detected at draw {i} of seed {seed}")
    return r2


@dataclass
class Builder:
    """runtime/builder.rs:7-162: the many-seed test driver (`#[madsim::test]`)."""

    seed: int = 0
    count: int = 1
    jobs: int = 1  # MADSIM_TEST_JOBS: the reference's worker threads; here the number of GPUs the sweep is sharded over
    config: Config = field(default_factory=Config)
    time_limit_s: float | None = None
    check: bool = False

    @staticmethod
    def from_env() -> "Builder":
        """The environment variables, defaults and rules of builder.rs:64-118: an unset MADSIM_TEST_SEED is the
        current time in nanoseconds, MADSIM_TEST_CONFIG names a TOML file, and under
        MADSIM_TEST_CHECK_DETERMINISM the count is at least 2."""
        import time

        b = Builder()
        env = os.environ
        b.seed = int(env["MADSIM_TEST_SEED"]) if "MADSIM_TEST_SEED" in env else time.time_ns() & ((1 << 64) - 1)
        if "MADSIM_TEST_JOBS" in env:
            b.jobs = int(env["MADSIM_TEST_JOBS"])
        if "MADSIM_TEST_CONFIG" in env:
            with open(env["MADSIM_TEST_CONFIG"]) as f:
                b.config = Config.from_toml(f.read())
        if "MADSIM_TEST_NUM" in env:
            b.count = int(env["MADSIM_TEST_NUM"])
        if "MADSIM_TEST_TIME_LIMIT" in env:
            b.time_limit_s = float(env["MADSIM_TEST_TIME_LIMIT"])
        b.check = "MADSIM_TEST_CHECK_DETERMINISM" in env
        if b.check:
            b.count = max(b.count, 2)
        return b

    def run(self, workload: Workload, device: int = 0):
        """Run seeds seed..seed+count; raise SimulationFailure for the first failing seed (builder.rs:129-159).
        With `check`, Runtime::check_determinism(seed) instead — the first seed only, exactly as builder.rs:126-128.
        `jobs` > 1 shards the range contiguously over that many GPUs (all launches are enqueued before the first
        result is read; the reference's jobs are OS threads, runtime/builder.rs:134-150)."""
        cfg = self.config
        if self.check:
            return check_determinism(self.seed, cfg, workload, device), None
        if self.time_limit_s is not None:
            cfg = Config(cfg.packet_loss_rate, cfg.send_latency_ns, int(self.time_limit_s * 1e9), cfg.buggify)
        from .dist import shard

        jobs = max(1, min(int(self.jobs), device_count(), self.count))
        batches = []
        try:
            for j in range(jobs):
                first, n = shard(self.seed, self.count, j, jobs)
                b = Batch(cfg, device + j if jobs > 1 else device).load(workload)
                batches.append((first, b))
                b.run(first, n, sync=False)
            parts, stats = [], None
            for first, b in batches:
                res, bitmap, st = b.results()
                if st["n_fail"]:
                    i = int(st["first_fail_seed"] - first)
                    raise SimulationFailure(first + i, int(res["status"][i]), int(res["fail_code"][i]))
                parts.append(res)
                if stats is None:
                    stats = dict(st)
                else:
                    for k in ("n_seeds", "ev_pop", "ev_fire", "draws", "n_pass", "n_fail", "n_unfinished", "launches", "escalated"):
                        stats[k] += st[k]
                    stats["kernel_ms"] = max(stats["kernel_ms"], st["kernel_ms"])
        finally:
            for _, b in batches:
                b.close()
        return np.concatenate(parts), stats
