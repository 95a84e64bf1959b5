// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points of liboracle.so: the CPU restatement of
// the reference executor (oracle/rt/*.hpp) running the registered workloads
// (oracle/workloads.hpp).  Used by tests/ (parity checker), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  NEVER linked into libmsim_b200.so.
//
// Parity status: UNPINNED by the reference's own tests — the reference holds no golden
// vectors for this path and cannot be compiled here (no Rust toolchain).
// The oracle is pinned instead by public KATs of the third-party arithmetic and by the
// reference's outcome assertions (tests/test_oracle_*.py).
#include <malloc.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

#include "workloads.hpp"

using namespace msim_oracle;

namespace {

Config to_config(const msim_config* c) {
  Config cfg;
  if (c) {
    cfg.net.packet_loss_rate = c->packet_loss_rate;
    cfg.net.send_latency_lo = c->send_latency_lo_ns;
    cfg.net.send_latency_hi = c->send_latency_hi_ns;
  }
  return cfg;
}

struct Extra {
  uint32_t peak_ready, peak_registered, peak_msgs, tasks_spawned;
  uint64_t sum_log2_heap;
};

thread_local size_t g_peak_registered = 0, g_peak_msgs = 0;

int run_one(const msim_config* c, const msim_workload_desc* wl, u64 seed, msim_seed_result* out,
            std::vector<msim_trace_rec>* full, std::string* msg, Extra* extra, std::vector<u8>* dlog = nullptr,
            const std::vector<u8>* dcheck = nullptr) {
  g_peak_registered = g_peak_msgs = 0;
  tls_peak_registered = &g_peak_registered;
  tls_peak_msgs = &g_peak_msgs;
  Runtime rt(seed, to_config(c));
  if (c && c->time_limit_ns) rt.set_time_limit(c->time_limit_ns);
  if (c && c->buggify) rt.rand.enable_buggify();
  // Runtime::check_determinism (runtime/mod.rs:185-187, 194-196): the log / the check is switched on after construction
  if (dcheck)
    rt.rand.enable_check(*dcheck);
  else if (dlog)
    rt.rand.enable_log();
  rt.tracer.full = full;
  msim_seed_result r{};
  std::array<u32, 4> user{0, 0, 0, 0};
  RaftOutcome ro;
  try {
    switch (wl->workload) {
      case MSIM_WL_SLEEP:
        rt.block_on(wl_sleep_main(wl->p[0]));
        break;
      case MSIM_WL_PINGPONG:
        user[0] = rt.block_on(wl_pingpong_main(wl->p[0], wl->p[1]));
        break;
      case MSIM_WL_RAFT: {
        auto cl = std::make_shared<RaftCluster>();
        cl->p.n = wl->n_nodes ? wl->n_nodes : 5;
        cl->p.dur_ms = wl->p[0];
        cl->p.fault_mask = u32(wl->p[1]);
        cl->p.bug = u32(wl->p[2]);
        cl->p.rpc_timeout_ms = wl->p[3];
        cl->p.hb_ms = wl->p[4];
        cl->p.el_lo_ms = wl->p[5];
        cl->p.el_hi_ms = wl->p[6];
        if (cl->p.n < 1 || cl->p.n > 8) return MSIM_EINVAL;
        rt.block_on(wl_raft_main(cl, &ro));
        user = {ro.leader, ro.term, ro.elections, ro.msgs};
        // break shared_ptr cycles before teardown
        rt.teardown();
        break;
      }
      case MSIM_WL_STREAM:
        rt.block_on(wl_stream_main(wl->p[0], wl->p[1], u32(wl->p[2]), &user));
        rt.teardown();
        break;
      case MSIM_WL_ETCD: {
        EtcdParams ep;
        ep.n_clients = (wl->n_nodes ? wl->n_nodes : 3) - 1;
        ep.ops = wl->p[0];
        ep.think_ns = wl->p[1];
        u32 bits = u32(wl->p[2]);
        std::memcpy(&ep.timeout_rate, &bits, 4);
        ep.bug = u32(wl->p[3]);
        ep.lease_ttl = u32(wl->p[4]);
        ep.lease_admin = u32(wl->p[7]);
        if (ep.n_clients < 1 || ep.n_clients > 3 || !(ep.timeout_rate >= 0.0f && ep.timeout_rate <= 1.0f)) return MSIM_EINVAL;
        rt.block_on(wl_etcd_main(ep, &user));
        rt.teardown();
        break;
      }
      case MSIM_WL_TONIC: {
        TonicParams tp;
        tp.n_clients = (wl->n_nodes ? wl->n_nodes : 7) - 1;
        tp.rounds = wl->p[0];
        tp.think_ns = wl->p[1];
        tp.flags = u32(wl->p[2]);
        if (tp.n_clients < 1 || tp.n_clients > 6 || tp.flags > 31 || ((tp.flags & 16) && (tp.flags & 10))) return MSIM_EINVAL;
        rt.block_on(wl_tonic_main(tp, &user));
        rt.teardown();
        break;
      }
      case MSIM_WL_SIGNAL: {
        SignalParams sp;
        sp.rounds = wl->p[0];
        sp.period_ms = wl->p[1];
        sp.behavior = u32(wl->p[2]);
        sp.flags = u32(wl->p[3]);
        if (sp.period_ms == 0 || sp.behavior > 2) return MSIM_EINVAL;
        rt.block_on(wl_signal_main(sp, &user));
        rt.teardown();
        break;
      }
      case MSIM_WL_CHAOS:
        if (wl->p[3] > 10 || (wl->p[3] && wl->p[2] > 10)) return MSIM_EINVAL;  // matching mode: single-digit panic messages
        rt.block_on(wl_chaos_main(wl->p[0], wl->p[1], u32(wl->p[2]), &user, u32(wl->p[3])));
        break;
      case MSIM_WL_ELECTION: {
        ElectionParams ep;
        ep.hold_s = wl->p[0];
        u32 bits = u32(wl->p[1]);
        std::memcpy(&ep.timeout_rate, &bits, 4);
        if (!(ep.timeout_rate >= 0.0f && ep.timeout_rate <= 1.0f) || ep.hold_s < 20 || ep.hold_s > 50) return MSIM_EINVAL;  // the observer counts two keys ~20 s in; leases last 60 s
        rt.block_on(wl_election_main(ep, &user));
        rt.teardown();
        break;
      }
      default:
        return MSIM_EINVAL;
    }
    r.status = MSIM_ST_PASS;
  } catch (Panic& e) {
    r.status = MSIM_ST_FAIL_ASSERT;
    r.fail_code = e.code;
    if (msg) *msg = e.what();
  } catch (Deadlock& e) {
    r.status = MSIM_ST_FAIL_DEADLOCK;
    if (msg) *msg = e.what();
  } catch (TimeLimitExceeded& e) {
    r.status = MSIM_ST_FAIL_TIME_LIMIT;
    if (msg) *msg = e.what();
  }
  r.trace_hash = rt.tracer.hash;
  r.elapsed_ns = rt.task.time.elapsed();
  r.draws = rt.rand.draws();
  r.ev_pop = u32(rt.tracer.n_pop);
  r.ev_fire = u32(rt.tracer.n_fire);
  for (int i = 0; i < 4; i++) r.user[i] = user[i];
  r.peak_timers = u32(rt.task.time.timer().peak());
  r.peak_tasks = rt.task.handle.sh.counters.peak;
  if (extra) {
    extra->peak_ready = u32(rt.task.peak_ready);
    extra->tasks_spawned = rt.task.handle.sh.next_task_id;
    extra->sum_log2_heap = rt.tracer.sum_log2_heap;
    extra->peak_registered = u32(g_peak_registered);
    extra->peak_msgs = u32(g_peak_msgs);
  }
  rt.tracer.full = nullptr;
  if (dlog) {
    auto l = rt.rand.take_log();
    if (l) *dlog = std::move(*l);
  }
  *out = r;
  return MSIM_OK;
}

}  // namespace

// The CPU baseline must not be handicapped by allocator artefacts: with the default glibc
// thresholds every Runtime teardown trims/unmaps memory, which serialises 100+ worker threads on
// the kernel's mmap lock (measured: 128 threads reached only ~9x one thread).  Keep freed memory
// in the per-thread arenas instead.
static void tune_malloc_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
  });
}

extern "C" {

int oracle_version(void) { return MSIM_VERSION; }

/// One seed; optionally the full canonical trace.
int oracle_run_seed(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed, msim_seed_result* out,
                    msim_trace_rec* trace, size_t cap, size_t* n_trace, char* msg, size_t msg_cap,
                    uint32_t* extra4 /* [6] or NULL */) {
  std::vector<msim_trace_rec> full;
  std::string m;
  Extra ex{};
  int rc = run_one(cfg, wl, seed, out, (trace || n_trace) ? &full : nullptr, &m, &ex);
  if (rc) return rc;
  if (n_trace) *n_trace = full.size();
  if (trace) std::memcpy(trace, full.data(), std::min(cap, full.size()) * sizeof(msim_trace_rec));
  if (msg && msg_cap) {
    std::strncpy(msg, m.c_str(), msg_cap - 1);
    msg[msg_cap - 1] = 0;
  }
  if (extra4) {
    extra4[0] = ex.peak_ready;
    extra4[1] = ex.peak_registered;
    extra4[2] = ex.peak_msgs;
    extra4[3] = ex.tasks_spawned;
    extra4[4] = uint32_t(ex.sum_log2_heap);
    extra4[5] = uint32_t(ex.sum_log2_heap >> 32);
  }
  return MSIM_OK;
}

/// Builder::run seed sweep (runtime/builder.rs:129-150): seeds seed_base..+n on `jobs` threads.
int oracle_run_many(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed_base, uint64_t n,
                    uint32_t jobs, msim_seed_result* out) {
  tune_malloc_once();
  if (jobs == 0) jobs = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> next{0};
  std::atomic<int> rc{0};
  auto worker = [&]() {
    for (;;) {
      uint64_t i = next.fetch_add(1);
      if (i >= n) break;
      int r = run_one(cfg, wl, seed_base + i, &out[i], nullptr, nullptr, nullptr);
      if (r) rc = r;
    }
  };
  std::vector<std::thread> th;
  for (uint32_t j = 1; j < jobs; j++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return rc;
}

/// Runtime::check_determinism (runtime/mod.rs:178-202): run twice, compare the per-draw log.
/// Returns 1 if deterministic, 0 if "non-determinism detected".
int oracle_check_determinism(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed) {
  auto run = [&](std::optional<std::vector<u8>> check) -> std::optional<std::vector<u8>> {
    Runtime rt(seed, to_config(cfg));
    if (check)
      rt.rand.enable_check(std::move(*check));
    else
      rt.rand.enable_log();
    try {
      if (wl->workload == MSIM_WL_SLEEP)
        rt.block_on(wl_sleep_main(wl->p[0]));
      else if (wl->workload == MSIM_WL_PINGPONG)
        rt.block_on(wl_pingpong_main(wl->p[0], wl->p[1]));
      else
        return std::nullopt;
    } catch (Panic&) {
      return std::nullopt;
    }
    return rt.rand.take_log();
  };
  auto log = run(std::nullopt);
  if (!log) return 0;
  auto log2 = run(log);
  return log2.has_value() ? 1 : 0;
}

/// The determinism log of one seed of ANY registered workload (rand.rs:64-89): what the first run of
/// Runtime::check_determinism records.  *n = bytes produced (may exceed cap).
int oracle_determinism_log(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed, uint8_t* out, size_t cap,
                           size_t* n, msim_seed_result* result) {
  std::vector<u8> log;
  msim_seed_result r{};
  int rc = run_one(cfg, wl, seed, &r, nullptr, nullptr, nullptr, &log);
  if (rc) return rc;
  if (n) *n = log.size();
  if (out) std::memcpy(out, log.data(), std::min(cap, log.size()));
  if (result) *result = r;
  return MSIM_OK;
}

/// The second run of Runtime::check_determinism against a given log: 1 = every draw matched, 0 = "non-determinism
/// detected" (the panic of rand.rs:79-84), for any registered workload.
int oracle_check_against_log(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed, const uint8_t* log, size_t n) {
  std::vector<u8> chk(log, log + n);
  msim_seed_result r{};
  std::string msg;
  int rc = run_one(cfg, wl, seed, &r, nullptr, &msg, nullptr, nullptr, &chk);
  if (rc) return rc;
  return msg.find("non-determinism") == std::string::npos ? 1 : 0;
}

// ---- KAT hooks: the third-party arithmetic in isolation ----
void oracle_kat_splitmix64(uint64_t seed, uint32_t n, uint64_t* out) {
  SplitMix64 sm{seed};
  for (uint32_t i = 0; i < n; i++) out[i] = sm.next();
}
void oracle_kat_xoshiro_state(const uint64_t s[4], uint32_t n, uint64_t* out) {
  Xoshiro256PlusPlus x{{s[0], s[1], s[2], s[3]}};
  for (uint32_t i = 0; i < n; i++) out[i] = x.next_u64();
}
void oracle_kat_seed_from_u64(uint64_t seed, uint64_t state[4], uint32_t n, uint64_t* out) {
  auto x = Xoshiro256PlusPlus::seed_from_u64(seed);
  for (int i = 0; i < 4; i++) state[i] = x.s[i];
  for (uint32_t i = 0; i < n; i++) out[i] = x.next_u64();
}
/// kind: 0 gen_range u64, 1 gen_range u32, 2 gen_range Duration(ns), 3 gen_bool(p = lo / 2^32 as a fraction … see tests)
void oracle_kat_draws(uint64_t seed, int kind, uint64_t lo, uint64_t hi, double p, uint32_t n, uint64_t* out,
                      uint64_t* draws_used) {
  GlobalRng g(seed);
  for (uint32_t i = 0; i < n; i++) {
    switch (kind) {
      case 0: out[i] = g.gen_range_u64(lo, hi); break;
      case 1: out[i] = g.gen_range_u32(u32(lo), u32(hi)); break;
      case 2: out[i] = g.gen_range_duration(lo, hi); break;
      default: out[i] = g.gen_bool(p) ? 1 : 0; break;
    }
  }
  if (draws_used) *draws_used = g.draws();
}
/// naive_timer / BinaryHeap order: push deadlines[0..n) (callback i), then expire everything;
/// out[k] = index of the k-th callback to fire.  `pops_after[i]` (optional, may be NULL):
/// after pushing entry i, pop this many entries first (interleaved push/pop sequences).
void oracle_kat_heap_order(const uint64_t* deadlines, uint32_t n, const uint32_t* pops_after, uint32_t* out) {
  Timer t;
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; i++) {
    t.add(deadlines[i], TimerMeta{}, [i, &k, out]() { out[k++] = i; });
    if (pops_after)
      for (uint32_t j = 0; j < pops_after[i] && t.len(); j++) {
        // pop exactly one: expire up to the current minimum, but equal deadlines would all go;
        // emulate `pop` with a one-shot guard
        uint64_t d = *t.next();
        bool fired = false;
        // expire(d) fires every entry with deadline <= d; to pop ONE we re-add the rest is
        // not faithful, so interleaving uses expire(d) semantics (what the reference does).
        t.expire(d, [&](const Timer::Event&) { fired = true; });
        (void)fired;
      }
  }
  t.expire(UINT64_MAX, [](const Timer::Event&) {});
}

}  // extern "C"
