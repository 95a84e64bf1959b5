/*
 * msim_b200.h — C-ABI of libmsim_b200.so: a B200-native many-seed executor for
 * madsim-style deterministic simulation.
 *
 * The reference (madsim-rs/madsim) has NO C-ABI or FFI seam for its executor
 * (SURVEY.md §0.3, §8b): the interface is the Rust API
 *   madsim::runtime::{Runtime, Handle, NodeBuilder, Builder}   madsim/src/sim/runtime/mod.rs:34-442, builder.rs:7-162
 *   madsim::task::{spawn, JoinHandle}                           madsim/src/sim/task/mod.rs:590-672
 *   madsim::time::{sleep, sleep_until, timeout, advance}        madsim/src/sim/time/mod.rs:108-198
 *   madsim::net::{Endpoint, NetSim}, net::rpc                   madsim/src/sim/net/{mod,endpoint,rpc}.rs
 * so every entry point below is NEW; each one cites the reference interface it
 * stands in for.  A Rust shim maps the names above onto these.
 *
 * Conventions
 *  - plain C types only; the caller allocates and owns every output buffer;
 *    the library owns the msim_batch and all device memory behind it.
 *  - functions return 0 (MSIM_OK) or a negative MSIM_E* code; the message is in
 *    msim_last_error() (thread-local).
 *  - per-seed outcomes are DATA (msim_seed_result.status), not errors — the
 *    reference turns them into a panic that carries MADSIM_TEST_SEED
 *    (runtime/mod.rs:205-210); the host shim does the same from the result.
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with MSIM_ENODEV.
 */
#ifndef MSIM_B200_H
#define MSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSIM_VERSION 0x00010000 /* 0.1.0 */

/* ---- return codes ------------------------------------------------------ */
#define MSIM_OK 0
#define MSIM_EINVAL (-1)   /* bad argument / unknown workload / bad parameter */
#define MSIM_ENODEV (-2)   /* no CUDA device (there is NO CPU fallback)        */
#define MSIM_ECUDA (-3)    /* CUDA runtime error, see msim_last_error()        */
#define MSIM_ENOMEM (-4)
#define MSIM_ESTATE (-5)   /* call order violated (e.g. results before run)   */
#define MSIM_ERANGE (-6)   /* seed not part of the batch / buffer too small    */

/* ---- per-seed outcome (what a panic / normal return is in the reference) */
#define MSIM_ST_RUNNING 0          /* event budget of the launch exhausted, resumable          */
#define MSIM_ST_PASS 1             /* main future returned (task/mod.rs:241-243)               */
#define MSIM_ST_FAIL_ASSERT 2      /* workload assertion panicked; fail_code says which        */
#define MSIM_ST_FAIL_DEADLOCK 3    /* "no events, all tasks will block forever" task/mod.rs:250 */
#define MSIM_ST_FAIL_TIME_LIMIT 4  /* "time limit exceeded" task/mod.rs:253-258                 */
#define MSIM_ST_FAIL_CAPACITY 5    /* a bounded per-seed table overflowed; never silent         */

/* fail_code values for MSIM_ST_FAIL_CAPACITY */
#define MSIM_CAP_TIMERS 1
#define MSIM_CAP_TASKS 2
#define MSIM_CAP_READY 3
#define MSIM_CAP_MSGS 4
#define MSIM_CAP_MAILBOX 5
#define MSIM_CAP_REGISTERED 6
#define MSIM_CAP_SOCKETS 7
#define MSIM_CAP_PARKED 8

/* ---- canonical trace ---------------------------------------------------- */
/* An "event" (SURVEY.md §3.2): one ready-queue pop or one timer fire.          */
#define MSIM_EV_POLL 1     /* pop → runnable.run()            task/mod.rs:279-283 */
#define MSIM_EV_DROP 2     /* pop → drop (cancelled / killed) task/mod.rs:269-273 */
#define MSIM_EV_PARK 3     /* pop → parked (node paused)      task/mod.rs:274-277 */
#define MSIM_EV_SKIP 4     /* pop → future already dropped    task/mod.rs:265-268 */
#define MSIM_EV_WAKE 5     /* timer → waker.wake()            time/sleep.rs:51-52 */
#define MSIM_EV_DELIVER 6  /* timer → socket.deliver()        net/mod.rs:323-330  */
#define MSIM_EV_RESTART 7  /* timer → Handle::restart(node)   task/mod.rs:311-313 */

typedef struct msim_trace_rec {
  uint64_t now_ns; /* virtual clock the event observes (pops: at pop time; fires: clock seen by the callback) */
  uint64_t aux;    /* pops: (ready_len_before << 32) | popped_index;  fires: timer deadline in ns             */
  uint32_t kind_id; /* (MSIM_EV_* << 24) | id24: task spawn index, message serial or node id                   */
  uint32_t draws;  /* xoshiro256++ steps consumed so far (low 32 bits)                                       */
} msim_trace_rec;

/* Per-seed trace hash: h0 = MSIM_HASH_INIT; for each record, in order,
 *   w = now_ns ^ rotl64(aux, 17) ^ rotl64(((uint64_t)kind_id << 32) | draws, 41)
 *   h = (h ^ w) * MSIM_HASH_MUL;  h ^= h >> 32;
 * (one 64-bit multiply per event: the hash is on the device's critical path)  */
#define MSIM_HASH_INIT 0xcbf29ce484222325ull
#define MSIM_HASH_MUL 0x9e3779b97f4a7c15ull

/* ---- configuration ------------------------------------------------------ */
/* Mirrors madsim::Config{ net: { packet_loss_rate, send_latency } } (config.rs:11-35,
 * network.rs:66-89) plus the Runtime knobs of builder.rs:64-118.                */
typedef struct msim_config {
  double packet_loss_rate;      /* default 0.0                                   */
  uint64_t send_latency_lo_ns;  /* default 1 ms                                  */
  uint64_t send_latency_hi_ns;  /* default 10 ms (exclusive)                     */
  uint64_t time_limit_ns;       /* 0 = none; Runtime::set_time_limit             */
  uint32_t buggify;             /* 0/1; rand.rs:119-134                          */
  uint32_t reserved0;
} msim_config;

/* ---- workloads ---------------------------------------------------------- */
/* Arbitrary Rust futures cannot run on a GPU; a workload is a registered
 * program (state machines in the device library, coroutines in the oracle).     */
#define MSIM_WL_SLEEP 0     /* block_on(sleep(p[0] ns)) — SURVEY Appendix B                    */
#define MSIM_WL_PINGPONG 1  /* 2-node ping-pong, BASELINE config 1 (endpoint.rs:550-584)       */
#define MSIM_WL_RAFT 2      /* 5-node Raft leader election, BASELINE configs 2 and 3           */
#define MSIM_WL_CHAOS 3     /* kill/restart/pause supervisor schedule (task/mod.rs:852-1016)   */
#define MSIM_WL_STREAM 4    /* reliable connect1/accept1 echo stream under partitions (net/mod.rs:337-430) */
#define MSIM_WL_TONIC 6     /* tonic-example Greeter under madsim-tonic + kill/restart chaos, BASELINE config 5 */
#define MSIM_WL_SIGNAL 7    /* time::interval + MissedTickBehavior, signal::ctrl_c, Handle::send_ctrl_c (interval.rs, signal.rs) */
#define MSIM_WL_ETCD 5      /* madsim-etcd-client SimServer + KV clients + linearizability check, BASELINE config 4 */
#define MSIM_WL_ELECTION 8  /* madsim-etcd-client election API: campaign / proclaim / leader / observe / resign over the watcher
                               (election.rs, service.rs:494-585, server.rs:71-95) — the reference's election() test            */

/* parameter slots p[] per workload
 *  SLEEP    p0 = duration ns
 *  PINGPONG p0 = rounds R (default 1000), p1 = think time ns between rounds (default 10 ms)
 *  RAFT     p0 = virtual duration ms (default 10000), p1 = fault mask (MSIM_RAFT_FAULT_*),
 *           p2 = bug mode (0 none, 1 double-vote), p3 = rpc timeout ms (default 40),
 *           p4 = heartbeat ms (default 50), p5 = election timeout lo ms (150), p6 = hi ms (300)
 *  STREAM   p0 = messages, p1 = think time ns, p2 = faults (1 = partition schedule)
 *  ETCD     n_nodes = 1 server + (n_nodes-1) clients (default 3); p0 = operations per client (default 32, clients x p0 <= 128),
 *           p1 = think time ns (default 50 ms), p2 = SimServer::timeout_rate as IEEE f32 bits (default 0),
 *           p3 = bug bits (1 = server read cache that delete does not invalidate,
 *           2 = clients also delete the leased key: reaches the crate's own expect("no lease")), p4 = lease ttl s (default 1),
 *           p7 = 1: the clients also use the rest of LeaseClient (revoke, keep_alive, time_to_live, leases)
 *  TONIC    n_nodes = 1 server + (n_nodes-1) client nodes (default 7); p0 = chaos rounds (default 10),
 *           p1 = think time ns between unary calls (default 100 ms), p2 = flags (1 = restart only, never kill;
 *           2 = clients also make server-streaming calls, sometimes dropping the response stream at once,
 *           client-streaming calls and bidi calls whose response stream is sometimes dropped early;
 *           4 = halfway through the supervisor kills the SERVER: calls in flight break, later ones are refused,
 *           the clients count failures (user[3]) instead of panicking;
 *           8 = unary calls run inside time::timeout (channel timeout 1 s) and some call `delay`, which must end in
 *           DeadlineExceeded after 1 s / 5 s;
 *           16 = GreeterServer::with_interceptor and every GreeterClient::with_interceptor turn every second request down
 *           (Status::unavailable; the client's before anything reaches the network) and some calls name a service the
 *           server does not have (Unimplemented) — tests/test.rs:280-363; not together with 2 or 8)
 *  SIGNAL   p0 = supervisor rounds (default 8), p1 = interval period ms (default 100),
 *           p2 = MissedTickBehavior (0 Burst, 1 Delay, 2 Skip), p3 = flags (1 = the worker installs a ctrl-c handler; default 1)
 *  ELECTION n_nodes = 4 (server, two candidates, an observer); p0 = seconds the first leader holds before it revokes its
 *           lease (default 30), p1 = SimServer::timeout_rate as IEEE f32 bits (default 0: any timed-out call fails the
 *           test, as its unwrap() does in the reference); user[] = {leader changes observed, final revision, checksum, requests}
 *  CHAOS    p0 = rounds, p1 = period ms, p2 = worker panics on its first p2 starts (restart_on_panic),
 *           p3 = k > 0: restart_on_panic_matching("0") .. ("k-1") instead; the worker panics with its start count as message
 *                                                         */
#define MSIM_RAFT_FAULT_CLOG_NODE 1u
#define MSIM_RAFT_FAULT_CLOG_LINK 2u
#define MSIM_RAFT_FAULT_CLOCK_JUMP 4u

typedef struct msim_workload_desc {
  uint32_t workload;
  uint32_t n_nodes; /* 0 = workload default */
  uint64_t p[8];
} msim_workload_desc;

/* ---- results -------------------------------------------------------------- */
typedef struct msim_seed_result { /* 64 bytes */
  uint64_t trace_hash;
  uint64_t elapsed_ns; /* virtual time at the end (TimeHandle::elapsed)        */
  uint64_t draws;      /* xoshiro steps consumed                                */
  uint32_t ev_pop;     /* ready-queue pops                                      */
  uint32_t ev_fire;    /* timer fires                                           */
  uint32_t status;     /* MSIM_ST_*                                             */
  uint32_t fail_code;  /* workload assertion id or MSIM_CAP_*                   */
  uint32_t user[4];    /* workload outputs (RAFT: leader node, term, elections, msg_count) */
  uint32_t peak_timers;
  uint32_t peak_tasks;
} msim_seed_result;

typedef struct msim_stats {
  uint64_t n_seeds;
  uint64_t ev_pop;  /* Σ over seeds */
  uint64_t ev_fire; /* Σ over seeds */
  uint64_t draws;
  uint64_t n_pass, n_fail, n_unfinished;
  uint64_t first_fail_seed; /* UINT64_MAX if none: builder.rs:129-150 prints the first panicking seed */
  double kernel_ms;         /* device time of the step kernels of the last run (CUDA events)            */
  uint32_t launches;        /* kernels launched by the last run                                        */
  uint32_t seeds_per_cta, lanes_per_warp, ctas, smem_bytes;
  uint32_t escalated;       /* seeds of the last run re-run on the large-capacity variant (capacity escalation)   */
} msim_stats;

typedef struct msim_batch msim_batch; /* opaque */

/* library version (no reference counterpart) */
int msim_version(void);
const char* msim_last_error(void);
/* number of CUDA devices visible, or MSIM_ENODEV */
int msim_device_count(void);

/* Runtime::with_seed_and_config for a whole batch (runtime/mod.rs:53-69): one
 * batch = many Runtimes that share a config.  device = CUDA ordinal.          */
int msim_config_default(msim_config* cfg);
/* `impl FromStr for Config` (config.rs:28-35): parse the TOML of a MADSIM_TEST_CONFIG file (builder.rs:82-88) —
 * `[net] packet_loss_rate = .., send_latency = { start = { secs, nanos }, end = { secs, nanos } }`, `[tcp]` —
 * over the defaults; time_limit_ns and buggify are not part of the file and come back 0.                      */
int msim_config_from_toml(const char* text, msim_config* cfg);
int msim_batch_create(const msim_config* cfg, int device, msim_batch** out);
void msim_batch_destroy(msim_batch* b);

/* What `f` is in Builder::run(f) (builder.rs:121): the program every seed runs. */
int msim_workload_default(uint32_t workload, msim_workload_desc* out);
int msim_workload_load(msim_batch* b, const msim_workload_desc* w);

/* Launch tuning: seeds per CTA, active lanes per warp (0 = auto).  lanes_per_warp | MSIM_GEOMETRY_LOCKSTEP sends
 * run-to-completion launches with >= 2 lanes through the lock-step kernel (a warp's lanes walk the phases of the
 * event step together: profiles/r2a_*; the default lets the lanes of a warp run free).                          */
#define MSIM_GEOMETRY_LOCKSTEP 0x100u
int msim_batch_set_geometry(msim_batch* b, uint32_t seeds_per_cta, uint32_t lanes_per_warp);
/* Optional: run on the caller's stream (a cudaStream_t, e.g. torch's current). */
int msim_batch_set_stream(msim_batch* b, void* cuda_stream);
/* Capacity escalation (default on).  The reference's executor grows its Vec / BinaryHeap / HashMap without bound
 * (task/mod.rs:239-259, naive-timer, net/network.rs); a device seed lives in bounded tables and reports
 * MSIM_ST_FAIL_CAPACITY when one overflows.  With escalation on, the seeds of a run-to-completion batch that
 * overflowed run again on the program's large-capacity variant (tables 2-4x) before any result is read, and
 * msim_batch_trace / msim_batch_determinism_log replay such a seed on that variant; msim_stats.escalated counts
 * them.  A seed that overflows even the large tables keeps MSIM_ST_FAIL_CAPACITY.  on = 0: report, never re-run.
 * on = 2: run EVERY seed of a run-to-completion batch on the large variant straight away (slower — the general kernel
 * at the large variant's occupancy; for sweeps known to overflow, and how the tests cover the large variants).   */
int msim_batch_set_capacity_escalation(msim_batch* b, int on);

/* Builder::run seed sweep (builder.rs:129-150): seeds seed_base..seed_base+n.
 * max_events_per_seed = 0 → run every seed to completion in one launch (at most one
 * persistent CTA per SM; a slot whose seed has finished pulls the next seed index from a
 * device-side queue, so seeds of very different lengths do not idle the GPU);
 * otherwise each launch advances every unfinished seed by at most that many
 * events, checkpoints the per-seed state to HBM and msim_batch_resume continues. */
int msim_batch_run(msim_batch* b, uint64_t seed_base, uint64_t n_seeds, uint64_t max_events_per_seed);
int msim_batch_resume(msim_batch* b, uint64_t max_events_per_seed);
/* asynchronous variant: enqueue only; msim_batch_sync waits                    */
int msim_batch_run_async(msim_batch* b, uint64_t seed_base, uint64_t n_seeds, uint64_t max_events_per_seed);
int msim_batch_sync(msim_batch* b);

/* results → HOST buffers (any pointer may be NULL)                             */
int msim_batch_results(msim_batch* b, msim_seed_result* out /*n_seeds*/, uint32_t* fail_bitmap /*ceil(n/32) words*/,
                       msim_stats* stats);
int msim_batch_trace_hash(msim_batch* b, uint64_t* hash /*n_seeds*/);
/* Device-side summary into caller-provided DEVICE memory (e.g. a torch tensor that is then
 * all-reduced over NCCL — the only cross-GPU step): u64 counters[8] =
 * {ev_pop, ev_fire, draws, n_pass, n_fail, n_unfinished, first_fail_seed, n_seeds} followed by
 * u32 fail_bitmap[ceil(n/32)].  Enqueued on the batch's stream.                              */
int msim_batch_summarize_device(msim_batch* b, void* d_out, size_t bytes);
/* The same for one shard of an N-GPU sweep over seeds seed_base..seed_base+n_total, shaped so that ONE sum
 * all-reduce of the buffer over the ranks gives the global result: u64 counters[8] (slot 6 unused), u64
 * first_fail_of_rank[world] (own slot: first failing seed or UINT64_MAX; others 0), u32 bitmap[ceil(n_total/32)]
 * holding this shard's bits at their GLOBAL position.  A rank with an empty shard calls it without having run.   */
int msim_batch_summarize_sharded(msim_batch* b, void* d_out, size_t bytes, uint32_t rank, uint32_t world,
                                 uint64_t seed_base, uint64_t n_total);
/* bytes of per-seed device state of a workload (capacity table in DESIGN.md)                */
int msim_workload_state_bytes(uint32_t workload);
/* device pointer of the result array (msim_seed_result[n_seeds]) for zero-copy consumers (torch / NCCL); waits for the
 * run on the host first (a capacity escalation may still have records to replace)                                  */
int msim_batch_results_device(msim_batch* b, void** dptr, size_t* bytes);
/* Full canonical trace of ONE seed: replays that seed on the device with the
 * trace ring enabled (replay is the reference's own debugging story:
 * MADSIM_TEST_SEED, runtime/mod.rs:205-210).  *n = records produced (may exceed cap). */
int msim_batch_trace(msim_batch* b, uint64_t seed, msim_trace_rec* out, size_t cap, size_t* n,
                     msim_seed_result* result);

/* The log of Runtime::check_determinism (runtime/mod.rs:178-202, rand.rs:64-89) for ONE seed (device replay): one
 * byte per GlobalRng::with call — every gen_range / gen_bool / random() of the runtime and of the program —
 * = (next u8 of a clone of the rng) ^ (xor of the bytes of the elapsed nanos).  Two runs of a deterministic program
 * give equal logs; the first differing index is the first diverging draw.  *n = bytes produced (may exceed cap). */
int msim_batch_determinism_log(msim_batch* b, uint64_t seed, uint8_t* out, size_t cap, size_t* n,
                               msim_seed_result* result);

#ifdef __cplusplus
}
#endif
#endif /* MSIM_B200_H */
