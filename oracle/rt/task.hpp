// ORACLE — TEST INFRASTRUCTURE ONLY (see future.hpp).
//
// Executor: restatement of madsim/src/sim/task/mod.rs:43-672, task/join.rs:1-183,
// utils/mpsc.rs:55-83, and the un-vendored `async-task 4.4` state machine
// (SCHEDULED / RUNNING / COMPLETED / CLOSED) that decides when a woken
// task is pushed to the ready Vec.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "time.hpp"

namespace msim_oracle {

struct NodeInfo;
struct Executor;

/// task/mod.rs:70-85.  Lifetime == lifetime of the task's future.
struct TaskCounters {
  u32 live = 0, peak = 0;
};
struct TaskInfo {
  u32 id;  // per-runtime spawn index (the reference's Id is process-global; SURVEY §7 hard part 8)
  std::string name;
  std::shared_ptr<NodeInfo> node;
  Waker waker;  // used for cancellation / kill
  bool cancelled = false;
  TaskCounters* counters = nullptr;
  ~TaskInfo() {
    if (counters) counters->live--;
  }
};

/// tokio::sync::watch<()> as NodeInfo::ctrl_c uses it (task/mod.rs:105-112,167-176): a version
/// counter + the wakers of the receivers parked in `changed()` (Notify::notify_waiters order =
/// registration order).
struct CtrlCChannel {
  u64 version = 0;
  u64 next_key = 1;
  std::vector<std::pair<u64, Waker>> waiters;
  /// watch::Sender::send(()): Err when no receiver is alive — the caller ignores it
  void send() {
    version++;
    auto ws = std::move(waiters);
    waiters.clear();
    for (auto& kv : ws) kv.second.wake();
  }
};

/// task/mod.rs:87-112
struct NodeInfo {
  u64 id;
  std::string name;  // empty = unnamed
  size_t cores = 1;
  bool restart_on_panic = false;
  std::vector<std::string> restart_on_panic_matching;
  bool paused = false;
  bool killed = false;
  std::vector<std::weak_ptr<TaskInfo>> tasks;  // spawn order
  std::shared_ptr<CtrlCChannel> ctrl_c;        // None until signal::ctrl_c() is first called on the node
  /// task/mod.rs:133-140
  void kill() {
    killed = true;
    auto drained = std::move(tasks);
    tasks.clear();
    for (auto& w : drained)
      if (auto t = w.lock()) t->waker.wake();
  }
  bool is_killed() const { return killed; }
  size_t num_tasks() {
    tasks.erase(std::remove_if(tasks.begin(), tasks.end(), [](auto& w) { return w.expired(); }), tasks.end());
    return tasks.size();
  }
};

/// Type-erased root future of a task.
struct DynFuture {
  virtual ~DynFuture() = default;
  virtual bool poll_root(Context& cx) = 0;  // true when Ready (output stored by the typed task)
};

/// async-task RawTask header + the madsim schedule fn `|runnable| sender.send(runnable)`.
struct RawTask : std::enable_shared_from_this<RawTask> {
  bool scheduled = false, running = false, completed = false, closed = false;
  std::unique_ptr<DynFuture> fut;        // dropped on completion / close
  std::shared_ptr<TaskInfo> info_strong;  // `let _info = info;` inside the future
  std::weak_ptr<TaskInfo> metadata;       // Runnable::metadata()
  std::optional<Waker> awaiter;
  std::vector<std::shared_ptr<RawTask>>* queue = nullptr;  // mpsc::Sender<Runnable>
  u32 id = 0;

  virtual ~RawTask() = default;
  void schedule() { queue->push_back(shared_from_this()); }  // utils/mpsc.rs:55-61
  /// Waker::wake
  void wake() {
    if (completed || closed) return;
    if (scheduled) return;
    scheduled = true;
    if (!running) schedule();
  }
  void drop_future() {
    fut.reset();          // the inner future first …
    info_strong.reset();  // … then `_info`
  }
  void notify_awaiter() {
    if (awaiter) {
      Waker w = std::move(*awaiter);
      awaiter.reset();
      w.wake();
    }
  }
  /// Runnable::run
  void run() {
    if (closed) {
      drop_future();
      scheduled = false;
      notify_awaiter();
      return;
    }
    scheduled = false;
    running = true;
    Context cx{Waker{shared_from_this()}};
    bool ready;
    try {
      ready = fut->poll_root(cx);
    } catch (...) {  // async-task's panic Guard: close, drop the future, notify the awaiter
      closed = true;
      running = false;
      scheduled = false;
      drop_future();
      notify_awaiter();
      throw;
    }
    if (ready) {
      drop_future();
      running = false;
      scheduled = false;
      completed = true;
      notify_awaiter();
    } else {
      running = false;
      if (closed) {
        scheduled = false;
        drop_future();
        notify_awaiter();
      } else if (scheduled) {
        schedule();  // woken while running
      }
    }
  }
  /// impl Drop for Runnable (the kill / cancel path, task/mod.rs:273)
  void drop_runnable() {
    if (!(completed || closed)) closed = true;
    drop_future();
    scheduled = false;
    notify_awaiter();
  }
  /// Task::cancel / drop of a FallibleTask obtained by cancel_on_drop
  void cancel() {
    if (completed || closed) return;
    closed = true;
    if (!scheduled && !running) {
      scheduled = true;
      schedule();
    }
  }
  bool is_finished() const { return completed || closed; }
};

inline void Waker::wake() const {
  if (task) task->wake();
}
inline u32 waker_task_id(const Waker& w) { return w.task ? w.task->id : 0xffffffu; }

template <class T>
struct TypedTask : RawTask {
  std::optional<T> output;
};

template <class F, class T>
struct RootFuture : DynFuture {
  F f;
  TypedTask<T>* owner;
  RootFuture(F&& f_, TypedTask<T>* o) : f(std::move(f_)), owner(o) {}
  bool poll_root(Context& cx) override {
    auto r = f.poll(cx);
    if (!r.ready()) return false;
    owner->output = std::move(r.v);
    return true;
  }
};

struct JoinError {
  u32 id;
  bool is_panic = false;
  bool is_cancelled() const { return !is_panic; }
};

/// task/join.rs:6-81
template <class T>
struct JoinHandle {
  std::shared_ptr<TypedTask<T>> task;
  std::weak_ptr<TaskInfo> info;
  u32 id_ = 0;
  u32 id() const { return id_; }
  bool valid() const { return bool(task); }
  bool is_finished() const { return task->is_finished(); }
  /// AbortHandle::abort (join.rs:158-163)
  void abort() const {
    if (auto i = info.lock()) {
      i->cancelled = true;
      i->waker.wake();
    }
  }
  /// async_task::Task::poll_task, single-threaded
  Poll<Result<T>> poll(Context& cx) {
    RawTask& t = *task;
    if (t.closed) {
      if (t.scheduled || t.running) {
        t.awaiter = cx.waker;
        return Poll<Result<T>>::pending();
      }
      return Poll<Result<T>>::done(Result<T>::Error(Err::Cancelled));
    }
    if (!t.completed) {
      t.awaiter = cx.waker;
      return Poll<Result<T>>::pending();
    }
    t.closed = true;
    return Poll<Result<T>>::done(Result<T>::Ok(std::move(*task->output)));
  }
  // Drop = detach (join.rs:75-81): nothing to do.
};

using InitFn = std::function<void(const struct Spawner&)>;

/// task/mod.rs:583-661
struct SpawnShared {
  std::vector<std::shared_ptr<RawTask>> queue;  // the ready "queue": a Vec popped at a random index
  u32 next_task_id = 0;
  TaskCounters counters;
  std::vector<std::weak_ptr<RawTask>> all_tasks;  // teardown only (breaks waker<->future cycles)
};
struct Spawner {
  SpawnShared* sh;
  std::shared_ptr<NodeInfo> info;

  u64 node_id() const { return info->id; }

  std::shared_ptr<TaskInfo> new_task(const std::string& name) const {  // NodeInfo::new_task :116-131
    auto t = std::make_shared<TaskInfo>();
    t->id = sh->next_task_id++;
    t->name = name;
    t->node = info;
    t->counters = &sh->counters;
    sh->counters.live++;
    sh->counters.peak = std::max(sh->counters.peak, sh->counters.live);
    info->tasks.push_back(t);
    return t;
  }
  /// spawn_inner (:627-654)
  template <class F>
  auto spawn(F fut, const std::string& name = "") const -> JoinHandle<poll_value_t<F>> {
    using T = poll_value_t<F>;
    auto ti = new_task(name);
    auto task = std::make_shared<TypedTask<T>>();
    task->id = ti->id;
    task->queue = &sh->queue;
    task->metadata = ti;
    task->info_strong = ti;
    task->fut = std::make_unique<RootFuture<F, T>>(std::move(fut), task.get());
    sh->all_tasks.push_back(task);
    ti->waker = Waker{task};
    task->scheduled = true;  // runnable.schedule()
    task->schedule();
    return JoinHandle<T>{task, ti, ti->id};
  }
  /// :656-661 — FIXME in the reference: paused list not cleared, no reset_node
  void exit() const { info->kill(); }
};

struct SimulatorHooks {  // plugin::Simulator (plugin.rs:18-40) — only the node lifecycle hooks
  virtual ~SimulatorHooks() = default;
  virtual void create_node(u64 id) = 0;
  virtual void reset_node(u64 id) = 0;
};

struct NodeBuilderSpec {
  std::string name;
  std::optional<u32> ip;
  std::optional<size_t> cores;
  InitFn init;
  bool restart_on_panic = false;
  std::vector<std::string> restart_on_panic_matching;
};

/// task/mod.rs:333-482 TaskHandle
struct TaskHandle {
  struct Node {
    std::shared_ptr<NodeInfo> info;
    std::vector<std::shared_ptr<RawTask>> paused;
    InitFn init;
  };
  SpawnShared sh;
  std::vector<std::shared_ptr<RawTask>>& queue = sh.queue;
  std::map<u64, Node> nodes;
  u64 next_node_id = 1;
  std::shared_ptr<NodeInfo> main_info;
  std::vector<SimulatorHooks*> sims;

  TaskHandle() {
    main_info = std::make_shared<NodeInfo>();
    main_info->id = 0;
    main_info->name = "madsim-main";
  }
  Spawner spawner_for(const std::shared_ptr<NodeInfo>& info) { return Spawner{&sh, info}; }
  /// Runtime teardown: drop every remaining future (Rust drops them with the queue / timer).
  void drop_all_tasks() {
    for (size_t i = 0; i < sh.all_tasks.size(); i++)  // futures may spawn while being dropped
      if (auto t = sh.all_tasks[i].lock()) {
        t->closed = true;
        t->drop_future();
        t->awaiter.reset();
      }
    sh.queue.clear();
    for (auto& kv : nodes) kv.second.paused.clear();
  }

  Node& node(u64 id) {
    auto it = nodes.find(id);
    if (it == nodes.end()) panic("node not found");
    return it->second;
  }
  /// kill_id :362-371
  void kill(u64 id) {
    Node& n = node(id);
    auto parked = std::move(n.paused);
    n.paused.clear();
    for (auto& r : parked) r->drop_runnable();
    n.info->kill();
    for (auto* s : sims) s->reset_node(id);
  }
  /// :374-401
  void restart(u64 id) {
    Node& n = node(id);
    auto ni = std::make_shared<NodeInfo>();
    ni->id = id;
    ni->name = n.info->name;
    ni->cores = n.info->cores;
    ni->restart_on_panic = n.info->restart_on_panic;
    ni->restart_on_panic_matching = n.info->restart_on_panic_matching;
    auto old = std::exchange(n.info, ni);
    auto parked = std::move(n.paused);
    n.paused.clear();
    for (auto& r : parked) r->drop_runnable();
    old->kill();
    if (n.init) n.init(spawner_for(n.info));
  }
  /// :427-441: delivered if a handler was ever installed on this NodeInfo, else the node is killed
  void send_ctrl_c(u64 id) {
    Node& n = node(id);
    if (n.info->ctrl_c) {
      n.info->ctrl_c->send();
      return;
    }
    kill(id);
  }
  void pause(u64 id) { node(id).info->paused = true; }
  void resume(u64 id) {
    Node& n = node(id);
    n.info->paused = false;
    auto parked = std::move(n.paused);
    n.paused.clear();
    for (auto& r : parked) queue.push_back(r);
  }
  bool is_exit(u64 id) { return node(id).info->killed; }
  /// :444-481
  Spawner create_node(const NodeBuilderSpec& b) {
    u64 id = next_node_id++;
    auto info = std::make_shared<NodeInfo>();
    info->id = id;
    info->name = b.name;
    info->cores = b.cores.value_or(1);
    info->restart_on_panic = b.restart_on_panic;
    info->restart_on_panic_matching = b.restart_on_panic_matching;
    Spawner h = spawner_for(info);
    if (b.init) b.init(h);
    nodes[id] = Node{info, {}, b.init};
    return h;
  }
  std::optional<Spawner> get_node(u64 id) {
    if (id == 0) return spawner_for(main_info);
    auto it = nodes.find(id);
    if (it == nodes.end()) return std::nullopt;
    return spawner_for(it->second.info);
  }
  size_t num_tasks() {
    size_t n = 0;
    for (auto& kv : nodes) n += kv.second.info->num_tasks();
    return n;
  }
};

/// Thread-local context (runtime/context.rs:9-64)
struct Runtime;
inline thread_local Runtime* tls_runtime = nullptr;
inline thread_local std::shared_ptr<TaskInfo> tls_task;

struct Deadlock : std::runtime_error {
  Deadlock() : std::runtime_error("no events, all tasks will block forever") {}
};
struct TimeLimitExceeded : std::runtime_error {
  TimeLimitExceeded() : std::runtime_error("time limit exceeded") {}
};

/// task/mod.rs:43-49,178-323
struct Executor {
  GlobalRng& rand;
  Tracer& tracer;
  TaskHandle handle;
  TimeHandle time;
  std::optional<u64> time_limit;
  size_t peak_ready = 0;

  Executor(GlobalRng& r, Tracer& tr) : rand(r), tracer(tr), time(r, tr) {}

  /// :220-260
  template <class F>
  auto block_on(F fut) -> poll_value_t<F> {
    using T = poll_value_t<F>;
    Spawner main = handle.spawner_for(handle.main_info);
    auto ti = main.new_task("");
    auto task = std::make_shared<TypedTask<T>>();
    task->id = ti->id;
    task->queue = &handle.queue;
    task->metadata = ti;
    task->info_strong = ti;
    task->fut = std::make_unique<RootFuture<F, T>>(std::move(fut), task.get());
    handle.sh.all_tasks.push_back(task);
    // NOTE: block_on does not set info.waker (stays a noop waker, :127)
    task->scheduled = true;
    task->schedule();
    ti.reset();  // `info` was moved into the future (task/mod.rs:229-231)
    for (;;) {
      run_all_ready();
      if (task->is_finished()) {
        if (!task->output) panic("main task was cancelled");
        return std::move(*task->output);
      }
      bool going = time.advance_to_next_event();
      if (!going) throw Deadlock();
      if (time_limit && time.elapsed() >= *time_limit) throw TimeLimitExceeded();
    }
  }

  /// utils/mpsc.rs:73-83
  std::shared_ptr<RawTask> try_recv_random(size_t& idx_out, size_t& len_out) {
    auto& q = handle.queue;
    if (q.empty()) return nullptr;
    peak_ready = std::max(peak_ready, q.size());
    size_t idx = size_t(rand.gen_range_u64(0, q.size()));
    idx_out = idx;
    len_out = q.size();
    auto r = std::move(q[idx]);  // swap_remove
    q[idx] = std::move(q.back());
    q.pop_back();
    return r;
  }

  /// :263-323
  void run_all_ready() {
    size_t idx, len;
    while (auto runnable = try_recv_random(idx, len)) {
      u64 aux = (u64(len) << 32) | u64(idx);
      auto info = runnable->metadata.lock();
      if (!info) {  // future has been dropped
        tracer.record(time.elapsed(), aux, MSIM_EV_SKIP, runnable->id, rand.draws());
        continue;
      }
      bool do_drop = info->cancelled || info->node->killed;
      if (!do_drop && info->node->paused) {
        tracer.record(time.elapsed(), aux, MSIM_EV_PARK, info->id, rand.draws());
        handle.node(info->node->id).paused.push_back(runnable);
        continue;
      }
      tracer.record(time.elapsed(), aux, do_drop ? MSIM_EV_DROP : MSIM_EV_POLL, info->id, rand.draws());
      {
        auto prev = std::exchange(tls_task, info);  // enter_task guard
        struct Restore {
          std::shared_ptr<TaskInfo> p;
          ~Restore() { tls_task = std::move(p); }
        } guard{std::move(prev)};
        try {
          if (do_drop)
            runnable->drop_runnable();
          else
            runnable->run();
        } catch (Panic& e) {
          std::string msg = e.what();
          bool matching = false;
          for (auto& s : info->node->restart_on_panic_matching)
            if (msg.find(s) != std::string::npos) matching = true;
          if (info->node->restart_on_panic || matching) {
            u64 node_id = info->node->id;
            u64 delay = rand.gen_range_duration(1000000000ull, 10000000000ull);
            handle.kill(node_id);
            TaskHandle* h = &handle;
            time.add_timer(delay, TimerMeta{MSIM_EV_RESTART, u32(node_id)}, [h, node_id]() { h->restart(node_id); });
          } else {
            throw;
          }
        }
      }
      // advance time: 50-100ns
      u64 dur = rand.gen_range_u64(50, 100);
      time.advance(dur);
    }
  }
};

}  // namespace msim_oracle
