// ORACLE — TEST INFRASTRUCTURE ONLY (see future.hpp).
//
// Runtime / Handle / NodeBuilder / Builder: restatement of
// madsim/src/sim/runtime/mod.rs:34-442, runtime/context.rs:9-64, runtime/builder.rs:7-162.
#pragma once
#include <atomic>
#include <cstdlib>
#include <thread>

#include "net.hpp"

namespace msim_oracle {

struct Config {  // config.rs:11-21
  NetConfig net;
};

/// runtime/mod.rs:34-210
struct Runtime {
  GlobalRng rand;
  Tracer tracer;
  Executor task;
  Config config;
  std::unique_ptr<NetSim> net;  // FsSim has no events and no draws
  bool torn_down = false;

  /// with_seed_and_config (:53-69): rng → executor (TimeRuntime::new draws) → simulators
  explicit Runtime(u64 seed = 0, Config cfg = Config{}) : rand(seed), task(rand, tracer), config(cfg) {
    rand.elapsed_ns_hook = [this]() { return task.time.elapsed(); };
    net = std::make_unique<NetSim>(rand, task.time, cfg.net);
    net->create_node(0);  // add_simulator: create node for supervisor (:80)
    task.handle.sims.push_back(net.get());
  }
  ~Runtime() { teardown(); }
  void teardown() {
    if (torn_down) return;
    torn_down = true;
    Runtime* prev = std::exchange(tls_runtime, this);
    task.handle.drop_all_tasks();
    tls_runtime = prev;
  }
  Runtime(const Runtime&) = delete;

  void set_time_limit(u64 ns) { task.time_limit = ns; }
  u64 seed() const { return rand.seed(); }

  /// block_on (:127-130): enter the context, run the executor
  template <class F>
  auto block_on(F fut) {
    Runtime* prev = std::exchange(tls_runtime, this);
    struct Restore {
      Runtime* p;
      ~Restore() { tls_runtime = p; }
    } g{prev};
    return task.block_on(std::move(fut));
  }

  // ---- Handle (:214-322) ----
  void kill(u64 id) { task.handle.kill(id); }
  void restart(u64 id) { task.handle.restart(id); }
  void pause(u64 id) { task.handle.pause(id); }
  void resume(u64 id) { task.handle.resume(id); }
  bool is_exit(u64 id) { return task.handle.is_exit(id); }
  void send_ctrl_c(u64 id) { task.handle.send_ctrl_c(id); }

  struct NodeHandle {
    Spawner task;
    u64 id() const { return task.node_id(); }
    template <class F>
    auto spawn(F f) const {
      return task.spawn(std::move(f));
    }
  };
  /// NodeBuilder (:325-418)
  struct NodeBuilder {
    Runtime* rt;
    NodeBuilderSpec spec;
    NodeBuilder& name(std::string n) {
      spec.name = std::move(n);
      return *this;
    }
    NodeBuilder& ip(u32 ip) {
      spec.ip = ip;
      return *this;
    }
    NodeBuilder& cores(size_t c) {
      spec.cores = c;
      return *this;
    }
    NodeBuilder& restart_on_panic() {
      spec.restart_on_panic = true;
      return *this;
    }
    NodeBuilder& restart_on_panic_matching(std::string m) {
      spec.restart_on_panic_matching.push_back(std::move(m));
      return *this;
    }
    /// init (:359-370): the task is respawned on restart; the node exits when it returns
    template <class NewTask>
    NodeBuilder& init(NewTask new_task) {
      spec.init = [new_task](const Spawner& handle) {
        handle.spawn(init_wrapper(new_task(), handle));
      };
      return *this;
    }
    /// build (:405-418): create_node runs init BEFORE the simulators learn about the node
    NodeHandle build() {
      Spawner task = rt->task.handle.create_node(spec);
      for (auto* sim : rt->task.handle.sims) sim->create_node(task.node_id());
      if (spec.ip) rt->net->set_ip(task.node_id(), *spec.ip);
      return NodeHandle{task};
    }

   private:
    template <class F>
    static Async<Unit> init_wrapper(F fut, Spawner h) {
      co_await fut;
      h.exit();
      co_return {};
    }
  };
  NodeBuilder create_node() { return NodeBuilder{this, {}}; }
};

inline Runtime& current_runtime() {
  if (!tls_runtime) panic("there is no reactor running, must be called from the context of a Madsim runtime");
  return *tls_runtime;
}
inline std::shared_ptr<TaskInfo> current_task() {
  if (!tls_task) panic("there is no task running");
  return tls_task;
}
inline NetSim& current_net() { return *current_runtime().net; }
inline GlobalRng& thread_rng() { return current_runtime().rand; }
inline GlobalRng& Endpoint::g_rand() { return thread_rng(); }

// ---- madsim::task / madsim::time free functions ----
template <class F>
auto spawn(F fut) {  // task/mod.rs:666-672: Spawner::current() = the current task's node
  auto& rt = current_runtime();
  Spawner s = rt.task.handle.spawner_for(current_task()->node);
  return s.spawn(std::move(fut));
}
inline Sleep sleep(u64 dur_ns) { return current_runtime().task.time.sleep(dur_ns); }
inline Sleep sleep_until(u64 deadline) { return current_runtime().task.time.sleep_until(deadline); }
inline u64 now() { return current_runtime().task.time.now_instant(); }  // Instant::now()
inline void advance(u64 dur_ns) { current_runtime().task.time.advance(dur_ns); }
template <class F>
Timeout<F> timeout(u64 dur_ns, F fut) {  // time/mod.rs:128-140: the Sleep is created NOW
  return Timeout<F>{std::move(fut), sleep(dur_ns)};
}
constexpr u64 MS = 1000000ull, SEC = 1000000000ull, US = 1000ull;

// ---- time/interval.rs:39-169 ----
enum class MissedTickBehavior { Burst, Delay, Skip };
struct Interval {
  Sleep delay;
  u64 period;
  MissedTickBehavior missed_tick_behavior = MissedTickBehavior::Burst;
  /// MissedTickBehavior::next_timeout (:73-100)
  u64 next_timeout(u64 timeout, u64 now_, u64 period_) const {
    switch (missed_tick_behavior) {
      case MissedTickBehavior::Burst: return timeout + period_;
      case MissedTickBehavior::Delay: return now_ + period_;
      case MissedTickBehavior::Skip: return now_ + period_ - ((now_ - timeout) % period_);
    }
    return 0;
  }
  /// poll_tick (:134-160): the Sleep, then re-arm it; yields the instant the tick was scheduled for
  Poll<u64> poll(Context& cx) {
    if (!delay.poll(cx).ready()) return Poll<u64>::pending();
    u64 timeout = delay.deadline;
    u64 now_ = now();
    u64 next = now_ > timeout + 5 * MS ? next_timeout(timeout, now_, period) : timeout + period;
    delay.reset(next);
    return Poll<u64>::done(timeout);
  }
  struct Tick {
    Interval* iv;
    Poll<u64> poll(Context& cx) { return iv->poll(cx); }
  };
  Tick tick() { return Tick{this}; }
  void set_missed_tick_behavior(MissedTickBehavior b) { missed_tick_behavior = b; }
};
inline Interval interval(u64 period_ns) {  // :39-42
  if (period_ns == 0) panic("`period` must be non-zero.");
  return Interval{sleep_until(now()), period_ns};
}

// ---- signal.rs:4-8 ctrl_c(): subscribe to the node's watch channel, then rx.changed() ----
struct CtrlC {
  std::shared_ptr<CtrlCChannel> ch;
  u64 seen;
  u64 key = 0;
  CtrlC(std::shared_ptr<CtrlCChannel> c) : ch(std::move(c)), seen(ch->version) {}
  CtrlC(CtrlC&& o) noexcept : ch(std::move(o.ch)), seen(o.seen), key(std::exchange(o.key, 0)) {}
  CtrlC(const CtrlC&) = delete;
  ~CtrlC() { unregister(); }
  void unregister() {
    if (!key || !ch) return;
    for (size_t i = 0; i < ch->waiters.size(); i++)
      if (ch->waiters[i].first == key) {
        ch->waiters.erase(ch->waiters.begin() + long(i));
        break;
      }
    key = 0;
  }
  Poll<Unit> poll(Context& cx) {
    if (ch->version != seen) {
      unregister();
      return Poll<Unit>::done({});
    }
    if (key) {
      for (auto& kv : ch->waiters)
        if (kv.first == key) kv.second = cx.waker;
    } else {
      key = ch->next_key++;
      ch->waiters.emplace_back(key, cx.waker);
    }
    return Poll<Unit>::pending();
  }
};
inline CtrlC ctrl_c() {
  auto node = current_task()->node;
  if (!node->ctrl_c) node->ctrl_c = std::make_shared<CtrlCChannel>();  // "ctrl-c signal handler installed"
  return CtrlC(node->ctrl_c);
}

// ---- rpc.rs:96-105, 134-166 ----
inline Async<Result<Payload>> Endpoint::call_timeout(SocketAddr dst, u64 req_tag, Payload request,
                                                     u64 timeout_ns) const {
  auto t = timeout(timeout_ns, call(dst, req_tag, request));
  auto r = co_await t;
  if (r.is_err()) co_return Result<Payload>::Error(Err::TimedOut);
  co_return std::move(*r.ok);
}
namespace detail {
inline Async<Unit> rpc_reply(Endpoint net, Async<Payload> rsp_future, SocketAddr from, u64 rsp_tag) {
  Payload rsp = co_await rsp_future;
  (co_await net.send_to_raw(from, rsp_tag, rsp)).unwrap();
  co_return {};
}
inline Async<Unit> rpc_loop(Endpoint net, u64 req_tag, std::function<Async<Payload>(Payload)> f) {
  for (;;) {
    auto r = co_await net.recv_from_raw(req_tag);
    auto [data, from] = r.unwrap();
    u64 rsp_tag = data.w;
    auto rsp_future = f(data);
    spawn(rpc_reply(net, std::move(rsp_future), from, rsp_tag));
  }
}
}  // namespace detail
inline void Endpoint::add_rpc_handler(u64 req_tag, std::function<Async<Payload>(Payload)> f) const {
  spawn(detail::rpc_loop(*this, req_tag, std::move(f)));
}

/// Outcome of one seed, in the C-ABI's vocabulary.
struct SeedOutcome {
  msim_seed_result r{};
  std::string message;
};

/// runtime/builder.rs:7-162 — the many-seed driver.
struct Builder {
  u64 seed = 0;
  u64 count = 1;
  u64 jobs = 1;
  Config config;
  std::optional<u64> time_limit;
  bool check = false;

  /// from_env (:64-118): MADSIM_TEST_{SEED,NUM,JOBS,TIME_LIMIT,CHECK_DETERMINISM}
  static Builder from_env() {
    Builder b;
    if (const char* s = std::getenv("MADSIM_TEST_SEED")) b.seed = std::strtoull(s, nullptr, 10);
    if (const char* s = std::getenv("MADSIM_TEST_NUM")) b.count = std::strtoull(s, nullptr, 10);
    if (const char* s = std::getenv("MADSIM_TEST_JOBS")) b.jobs = std::strtoull(s, nullptr, 10);
    if (const char* s = std::getenv("MADSIM_TEST_TIME_LIMIT")) b.time_limit = u64(std::strtod(s, nullptr) * 1e9);
    b.check = std::getenv("MADSIM_TEST_CHECK_DETERMINISM") != nullptr;
    return b;
  }
};

}  // namespace msim_oracle
