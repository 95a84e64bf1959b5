// msim_b200.hpp — C++ host mirror of the reference's runtime interface, over the C-ABI
// (include/msim_b200.h → libmsim_b200.so).  Header-only; same names, argument meaning and
// error behaviour as the Rust API for the hot path:
//
//   madsim::Config                         madsim/src/sim/config.rs:11-35, net/network.rs:66-89
//   madsim::runtime::Runtime               madsim/src/sim/runtime/mod.rs:34-210
//       ::with_seed_and_config(seed, cfg)    :53-69
//       ::block_on(workload)                 :127-130   (one seed; panics ⇒ exception with the seed note)
//       ::set_time_limit(d)                  :132-137
//       ::check_determinism(seed, cfg, wl)   :178-202   (run twice, compare the per-draw rng log of rand.rs:64-89)
//   madsim::runtime::Builder               madsim/src/sim/runtime/builder.rs:7-162
//       ::from_env()                         :64-118    MADSIM_TEST_{SEED,NUM,JOBS,CONFIG,TIME_LIMIT,CHECK_DETERMINISM}
//                                                        with the reference's defaults (time-based seed, count >= 2
//                                                        under the check)
//       ::run(workload)                      :121-162   seeds seed..seed+count on the GPU(s); the first
//                                                        failing seed is re-raised with the reference's note
//
// A `Workload` is the registered program that stands where the reference takes an async closure
// (arbitrary futures cannot run on a GPU; see include/msim_b200.h MSIM_WL_*).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "msim_b200.h"

namespace madsim {

/// A panic that escaped `block_on` in the reference (runtime/mod.rs:205-210).
struct SimulationFailure : std::runtime_error {
  uint64_t seed;
  uint32_t status, fail_code;
  SimulationFailure(uint64_t s, uint32_t st, uint32_t code)
      : std::runtime_error("seed " + std::to_string(s) + " failed: status " + std::to_string(st) + " code " +
                           std::to_string(code) + "\nnote: run with `MADSIM_TEST_SEED=" + std::to_string(s) +
                           "` environment variable to reproduce this error"),
        seed(s), status(st), fail_code(code) {}
};
/// any non-zero return of the C-ABI
struct MsimError : std::runtime_error {
  int code;
  MsimError(int c, const char* m) : std::runtime_error(std::string("msim error ") + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
  if (rc != MSIM_OK) throw MsimError(rc, msim_last_error());
}

/// madsim::Config { net: { packet_loss_rate, send_latency } }
struct Config {
  struct Net {
    double packet_loss_rate = 0.0;
    uint64_t send_latency_lo_ns = 1000000, send_latency_hi_ns = 10000000;  // 1 ms..10 ms
  } net;
  bool buggify = false;
  /// `impl FromStr for Config` (config.rs:28-35): TOML text of a MADSIM_TEST_CONFIG file
  static Config from_toml(const std::string& text) {
    msim_config c{};
    check(msim_config_from_toml(text.c_str(), &c));
    Config out;
    out.net.packet_loss_rate = c.packet_loss_rate;
    out.net.send_latency_lo_ns = c.send_latency_lo_ns;
    out.net.send_latency_hi_ns = c.send_latency_hi_ns;
    return out;
  }
  msim_config to_c(uint64_t time_limit_ns = 0) const {
    msim_config c{};
    c.packet_loss_rate = net.packet_loss_rate;
    c.send_latency_lo_ns = net.send_latency_lo_ns;
    c.send_latency_hi_ns = net.send_latency_hi_ns;
    c.time_limit_ns = time_limit_ns;
    c.buggify = buggify ? 1 : 0;
    return c;
  }
};

struct Workload {
  msim_workload_desc desc{};
  static Workload sleep(uint64_t duration_ns) {
    Workload w;
    check(msim_workload_default(MSIM_WL_SLEEP, &w.desc));
    w.desc.p[0] = duration_ns;
    return w;
  }
  static Workload pingpong(uint64_t rounds = 1000, uint64_t think_ns = 10000000) {
    Workload w;
    check(msim_workload_default(MSIM_WL_PINGPONG, &w.desc));
    w.desc.p[0] = rounds, w.desc.p[1] = think_ns;
    return w;
  }
  static Workload raft(uint32_t n_nodes = 5, uint64_t duration_ms = 10000, uint32_t faults = 0, uint32_t bug = 0) {
    Workload w;
    check(msim_workload_default(MSIM_WL_RAFT, &w.desc));
    w.desc.n_nodes = n_nodes, w.desc.p[0] = duration_ms, w.desc.p[1] = faults, w.desc.p[2] = bug;
    return w;
  }
  static Workload chaos(uint64_t rounds = 8, uint64_t period_ms = 2000, uint64_t panics = 0) {
    Workload w;
    check(msim_workload_default(MSIM_WL_CHAOS, &w.desc));
    w.desc.p[0] = rounds, w.desc.p[1] = period_ms, w.desc.p[2] = panics;
    return w;
  }
  /// connect1 / accept1 echo stream under partitions (net/mod.rs:337-430)
  static Workload stream(uint64_t messages = 200, uint64_t think_ns = 5000000, uint32_t faults = 1) {
    Workload w;
    check(msim_workload_default(MSIM_WL_STREAM, &w.desc));
    w.desc.p[0] = messages, w.desc.p[1] = think_ns, w.desc.p[2] = faults;
    return w;
  }
  /// madsim-etcd-client SimServer (timeout_rate: server.rs:21-25) + KV clients + linearizability check
  static Workload etcd(uint32_t clients = 2, uint64_t ops = 32, float timeout_rate = 0.0f, uint32_t bug = 0,
                       bool lease_admin = false) {
    Workload w;
    check(msim_workload_default(MSIM_WL_ETCD, &w.desc));
    uint32_t bits;
    static_assert(sizeof(bits) == sizeof(timeout_rate), "f32");
    std::memcpy(&bits, &timeout_rate, 4);
    w.desc.n_nodes = clients + 1, w.desc.p[0] = ops, w.desc.p[2] = bits, w.desc.p[3] = bug, w.desc.p[7] = lease_admin ? 1 : 0;
    return w;
  }
  /// madsim-etcd-client's election API as its own election() test drives it (tests/test.rs:129-232)
  static Workload election(uint64_t hold_s = 30, float timeout_rate = 0.0f) {
    Workload w;
    check(msim_workload_default(MSIM_WL_ELECTION, &w.desc));
    uint32_t bits;
    std::memcpy(&bits, &timeout_rate, 4);
    w.desc.p[0] = hold_s, w.desc.p[1] = bits;
    return w;
  }
  /// time::interval + MissedTickBehavior (0 Burst, 1 Delay, 2 Skip), signal::ctrl_c, Handle::send_ctrl_c
  static Workload signal(uint64_t rounds = 8, uint64_t period_ms = 100, uint32_t behavior = 0, bool catch_ctrl_c = true) {
    Workload w;
    check(msim_workload_default(MSIM_WL_SIGNAL, &w.desc));
    w.desc.p[0] = rounds, w.desc.p[1] = period_ms, w.desc.p[2] = behavior, w.desc.p[3] = catch_ctrl_c ? 1 : 0;
    return w;
  }
  /// tonic-example Greeter under madsim-tonic + kill/restart chaos of the client nodes
  static Workload tonic(uint32_t clients = 6, uint64_t rounds = 10, uint64_t think_ns = 100000000, bool restart_only = false,
                        bool streaming = false, bool server_crash = false, bool timeouts = false, bool interceptors = false) {
    Workload w;
    check(msim_workload_default(MSIM_WL_TONIC, &w.desc));
    w.desc.n_nodes = clients + 1, w.desc.p[0] = rounds, w.desc.p[1] = think_ns, w.desc.p[2] = (restart_only ? 1 : 0) | (streaming ? 2 : 0) | (server_crash ? 4 : 0) | (timeouts ? 8 : 0) | (interceptors ? 16 : 0);
    return w;
  }
};

namespace runtime {

/// RAII owner of a msim_batch
class Batch {
 public:
  Batch(const Config& cfg, uint64_t time_limit_ns, int device) {
    msim_config c = cfg.to_c(time_limit_ns);
    check(msim_batch_create(&c, device, &b_));
  }
  ~Batch() { msim_batch_destroy(b_); }
  Batch(const Batch&) = delete;
  Batch& operator=(const Batch&) = delete;
  msim_batch* get() const { return b_; }
  /// seeds that overflow a bounded table re-run on the program's large-capacity variant (default on)
  Batch& set_capacity_escalation(bool on) {
    check(msim_batch_set_capacity_escalation(b_, on ? 1 : 0));
    return *this;
  }

 private:
  msim_batch* b_ = nullptr;
};

struct RunOutput {
  std::vector<msim_seed_result> results;
  msim_stats stats{};
};

inline RunOutput run_seeds(const Config& cfg, uint64_t time_limit_ns, const Workload& wl, uint64_t seed, uint64_t count,
                           int device = 0) {
  Batch b(cfg, time_limit_ns, device);
  check(msim_workload_load(b.get(), &wl.desc));
  check(msim_batch_run(b.get(), seed, count, 0));
  RunOutput out;
  out.results.resize(count);
  check(msim_batch_results(b.get(), out.results.data(), nullptr, &out.stats));
  return out;
}

/// madsim::runtime::Runtime — one seed
class Runtime {
 public:
  Runtime() : Runtime(0, Config{}) {}
  static Runtime with_seed_and_config(uint64_t seed, Config cfg) { return Runtime(seed, std::move(cfg)); }
  void set_time_limit(uint64_t ns) { time_limit_ns_ = ns; }
  uint64_t seed() const { return seed_; }
  /// Runs the program to completion; a failing seed is the reference's panic
  msim_seed_result block_on(const Workload& wl) const {
    RunOutput o = run_seeds(cfg_, time_limit_ns_, wl, seed_, 1);
    const msim_seed_result& r = o.results[0];
    if (r.status != MSIM_ST_PASS) throw SimulationFailure(seed_, r.status, r.fail_code);
    return r;
  }
  /// Runtime::check_determinism (runtime/mod.rs:178-202): the first run records one byte per GlobalRng::with call,
  /// (next u8 of a clone of the rng) ^ fold(elapsed); the second run must reproduce every byte
  static msim_seed_result check_determinism(uint64_t seed, Config cfg, const Workload& wl, int device = 0) {
    Batch b(cfg, 0, device);
    check(msim_workload_load(b.get(), &wl.desc));
    auto log_of = [&](msim_seed_result* r) {
      size_t n = 0;
      check(msim_batch_determinism_log(b.get(), seed, nullptr, 0, &n, r));
      std::vector<uint8_t> log(n);
      check(msim_batch_determinism_log(b.get(), seed, log.data(), n, &n, r));
      return log;
    };
    msim_seed_result r1{}, r2{};
    std::vector<uint8_t> first = log_of(&r1);
    if (r1.status != MSIM_ST_PASS) throw SimulationFailure(seed, r1.status, r1.fail_code);
    std::vector<uint8_t> second = log_of(&r2);
    for (size_t i = 0; i < first.size() || i < second.size(); i++)
      if (i >= first.size() || i >= second.size() || first[i] != second[i])
        throw std::runtime_error("non-determinism detected at draw " + std::to_string(i));
    return r2;
  }

 private:
  Runtime(uint64_t seed, Config cfg) : seed_(seed), cfg_(std::move(cfg)) {}
  uint64_t seed_;
  Config cfg_;
  uint64_t time_limit_ns_ = 0;
};

/// madsim::runtime::Builder — the many-seed test driver behind #[madsim::test]
struct Builder {
  uint64_t seed = 0;
  uint64_t count = 1;
  uint64_t jobs = 1;  // MADSIM_TEST_JOBS: GPUs a sweep may use (one process per GPU shards the range)
  Config config;
  uint64_t time_limit_ns = 0;
  bool check = false;

  static Builder from_env() {
    Builder b;
    if (const char* s = std::getenv("MADSIM_TEST_SEED"))
      b.seed = std::strtoull(s, nullptr, 10);
    else  // builder.rs:69-74: the current time in nanoseconds
      b.seed = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    if (const char* s = std::getenv("MADSIM_TEST_JOBS")) b.jobs = std::strtoull(s, nullptr, 10);
    if (const char* s = std::getenv("MADSIM_TEST_CONFIG")) {
      std::ifstream f(s);
      if (!f) throw std::runtime_error("failed to read config file");
      std::stringstream ss;
      ss << f.rdbuf();
      b.config = Config::from_toml(ss.str());
    }
    if (const char* s = std::getenv("MADSIM_TEST_NUM")) b.count = std::strtoull(s, nullptr, 10);
    if (const char* s = std::getenv("MADSIM_TEST_TIME_LIMIT")) b.time_limit_ns = uint64_t(std::strtod(s, nullptr) * 1e9);
    b.check = std::getenv("MADSIM_TEST_CHECK_DETERMINISM") != nullptr;
    if (b.check && b.count < 2) b.count = 2;
    return b;
  }
  /// seeds seed..seed+count; throws SimulationFailure for the FIRST failing seed (builder.rs:155-159).  Under `check`
  /// only `seed` runs, through Runtime::check_determinism (builder.rs:126-128).  `jobs` > 1 shards the range over that
  /// many GPUs: every launch is enqueued before the first result is read.
  RunOutput run(const Workload& wl, int device = 0) const {
    if (check) {
      RunOutput o;
      o.results.push_back(Runtime::check_determinism(seed, config, wl, device));
      return o;
    }
    int n_dev = msim_device_count();
    if (n_dev < 0) throw MsimError(n_dev, msim_last_error());
    uint64_t shards = jobs < 1 ? 1 : jobs;
    if (shards > (uint64_t)n_dev) shards = n_dev;
    if (shards > count) shards = count ? count : 1;
    std::vector<std::unique_ptr<Batch>> bs;
    std::vector<std::pair<uint64_t, uint64_t>> ranges;
    uint64_t first = seed;
    for (uint64_t j = 0; j < shards; j++) {
      uint64_t n = count / shards + (j < count % shards ? 1 : 0);
      bs.emplace_back(new Batch(config, time_limit_ns, shards > 1 ? device + (int)j : device));
      madsim::check(msim_workload_load(bs.back()->get(), &wl.desc));
      madsim::check(msim_batch_run_async(bs.back()->get(), first, n, 0));
      ranges.emplace_back(first, n);
      first += n;
    }
    RunOutput o;
    for (uint64_t j = 0; j < shards; j++) {
      std::vector<msim_seed_result> part(ranges[j].second);
      msim_stats st{};
      madsim::check(msim_batch_results(bs[j]->get(), part.data(), nullptr, &st));
      if (st.n_fail) {
        const msim_seed_result& r = part[st.first_fail_seed - ranges[j].first];
        throw SimulationFailure(st.first_fail_seed, r.status, r.fail_code);
      }
      o.results.insert(o.results.end(), part.begin(), part.end());
      o.stats.n_seeds += st.n_seeds, o.stats.ev_pop += st.ev_pop, o.stats.ev_fire += st.ev_fire, o.stats.draws += st.draws;
      o.stats.n_pass += st.n_pass, o.stats.n_unfinished += st.n_unfinished;
      o.stats.first_fail_seed = ~0ull;
      if (st.kernel_ms > o.stats.kernel_ms) o.stats.kernel_ms = st.kernel_ms;
      o.stats.launches += st.launches;
    }
    return o;
  }
};

}  // namespace runtime
}  // namespace madsim
