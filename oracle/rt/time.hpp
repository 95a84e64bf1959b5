// ORACLE — TEST INFRASTRUCTURE ONLY (see future.hpp).
//
// Virtual time: restatement of madsim/src/sim/time/mod.rs:20-255 and time/sleep.rs:20-55,
// plus the un-vendored `naive-timer 0.2` (madsim/Cargo.toml:42) which is a
// std::collections::BinaryHeap<Event> ordered by reversed deadline.  The heap below copies
// Rust's BinaryHeap::push (sift_up) and BinaryHeap::pop (swap last into root,
// sift_down_to_bottom, then sift_up) literally, because they decide the order in which
// equal-deadline timers fire.
#pragma once
#include <algorithm>
#include <cassert>
#include <vector>

#include "../../include/msim_b200.h"
#include "rand.hpp"

namespace msim_oracle {

/// Canonical trace sink (include/msim_b200.h: msim_trace_rec, MSIM_HASH_*).
struct Tracer {
  u64 hash = MSIM_HASH_INIT;
  u64 n_pop = 0, n_fire = 0;
  u64 sum_log2_heap = 0;  // Σ floor(log2(heap length)) over timer fires: roofline byte model
  std::vector<msim_trace_rec>* full = nullptr;  // optional full trace
  static u64 mix(u64 h, u64 w) {
    h ^= w;
    h *= MSIM_HASH_MUL;
    h ^= h >> 32;
    return h;
  }
  void record(u64 now_ns, u64 aux, u32 kind, u32 id24, u64 draws) {
    u32 kind_id = (kind << 24) | (id24 & 0xffffffu);
    u32 d = u32(draws);
    u64 kd = (u64(kind_id) << 32) | d;
    hash = mix(hash, now_ns ^ ((aux << 17) | (aux >> 47)) ^ ((kd << 41) | (kd >> 23)));
    if (kind <= MSIM_EV_SKIP)
      n_pop++;
    else
      n_fire++;
    if (full) full->push_back(msim_trace_rec{now_ns, aux, kind_id, d});
  }
};

/// What a timer callback is, for the trace (the reference stores an opaque boxed closure).
struct TimerMeta {
  u32 kind = MSIM_EV_WAKE;
  u32 id = 0;
};

/// naive_timer::Timer
class Timer {
 public:
  struct Event {
    u64 deadline;  // Duration since the clock's base instant, in ns
    TimerMeta meta;
    std::function<void()> callback;
  };
  void add(u64 deadline, TimerMeta meta, std::function<void()> cb) {
    // BinaryHeap::push
    size_t old_len = ev_.size();
    ev_.push_back(Event{deadline, meta, std::move(cb)});
    sift_up(0, old_len);
    peak_ = std::max(peak_, ev_.size());
  }
  std::optional<u64> next() const {
    if (ev_.empty()) return std::nullopt;
    return ev_[0].deadline;
  }
  /// `expire(now)`: while peek.deadline <= now { pop; call }
  template <class OnFire>
  void expire(u64 now, OnFire&& on_fire) {
    assert(!expiring_ && "timer callbacks never re-enter the timer (spin lock held, time/mod.rs:46,105)");
    expiring_ = true;
    while (!ev_.empty() && ev_[0].deadline <= now) {
      len_at_fire_ = ev_.size();
      Event e = pop();
      on_fire(e);
      e.callback();
    }
    expiring_ = false;
  }
  size_t len() const { return ev_.size(); }
  size_t peak() const { return peak_; }
  size_t len_at_fire() const { return len_at_fire_; }
  bool expiring() const { return expiring_; }

 private:
  // Ord for Event: other.deadline.cmp(&self.deadline)  ⇒  a <= b  ⇔  a.deadline >= b.deadline
  static bool le(const Event& a, const Event& b) { return a.deadline >= b.deadline; }
  size_t sift_up(size_t start, size_t pos) {
    Event elem = std::move(ev_[pos]);
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elem, ev_[parent])) break;
      ev_[pos] = std::move(ev_[parent]);
      pos = parent;
    }
    ev_[pos] = std::move(elem);
    return pos;
  }
  void sift_down_to_bottom(size_t pos) {
    size_t end = ev_.size();
    size_t start = pos;
    Event elem = std::move(ev_[pos]);
    size_t child = 2 * pos + 1;
    while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {  // child <= end.saturating_sub(2)
      if (le(ev_[child], ev_[child + 1])) child += 1;
      ev_[pos] = std::move(ev_[child]);
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      ev_[pos] = std::move(ev_[child]);
      pos = child;
    }
    ev_[pos] = std::move(elem);
    sift_up(start, pos);
  }
  Event pop() {
    Event item = std::move(ev_.back());
    ev_.pop_back();
    if (!ev_.empty()) {
      std::swap(item, ev_[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }
  std::vector<Event> ev_;
  size_t peak_ = 0, len_at_fire_ = 1;
  bool expiring_ = false;
};

/// time/mod.rs:200-255 Clock (base_instant is zeroed ⇒ Instant == elapsed ns).
struct Clock {
  u64 base_time_secs = 0;  // SystemTime base (seconds since epoch)
  u64 advance_ns = 0;
  u64 elapsed() const { return advance_ns; }
  void set_elapsed(u64 t) { advance_ns = t; }
  u64 advance(u64 d) {
    advance_ns += d;
    return advance_ns;
  }
  u64 now_instant() const { return advance_ns; }
};

class TimeHandle;
/// time/sleep.rs:20-55
struct Sleep {
  TimeHandle* handle;
  u64 deadline;  // Instant
  bool is_elapsed() const;
  void reset(u64 d) { deadline = d; }
  Poll<Unit> poll(Context& cx);
};

/// time/mod.rs:64-170 TimeHandle (+ TimeRuntime::advance_to_next_event, :45-60)
class TimeHandle {
 public:
  TimeHandle(GlobalRng& rand, Tracer& tr) : rand_(rand), tracer_(tr) {
    // TimeRuntime::new (time/mod.rs:26-38): the FIRST draw of every runtime.
    clock_.base_time_secs = 60ull * 60 * 24 * 365 * (2022 - 1970) + rand.gen_range_u64(0, 60ull * 60 * 24 * 365);
  }
  u64 now_instant() const { return clock_.now_instant(); }
  u64 elapsed() const { return clock_.elapsed(); }
  u64 now_time_secs() const { return clock_.base_time_secs + clock_.advance_ns / 1000000000ull; }
  u64 base_time_secs() const { return clock_.base_time_secs; }

  /// time/mod.rs:103-106
  void advance(u64 dur) {
    u64 t = clock_.advance(dur);
    expire(t);
  }
  /// time/mod.rs:45-60
  bool advance_to_next_event() {
    auto t = timer_.next();
    if (!t) return false;
    u64 time = *t + 50;
    expire(time);  // callbacks observe the OLD clock
    clock_.set_elapsed(time);
    return true;
  }
  /// time/mod.rs:111-124
  Sleep sleep(u64 dur) { return sleep_until(clock_.now_instant() + dur); }
  Sleep sleep_until(u64 deadline) {
    u64 min_deadline = clock_.now_instant() + 1000000ull;
    return Sleep{this, std::max(deadline, min_deadline)};
  }
  /// time/mod.rs:158-169
  void add_timer_at(u64 deadline, TimerMeta meta, std::function<void()> cb) {
    timer_.add(deadline, meta, std::move(cb));  // deadline - base_instant (== 0)
  }
  void add_timer(u64 dur, TimerMeta meta, std::function<void()> cb) {
    add_timer_at(clock_.now_instant() + dur, meta, std::move(cb));
  }
  const Timer& timer() const { return timer_; }

 private:
  void expire(u64 now) {
    timer_.expire(now, [&](const Timer::Event& e) {
      tracer_.record(clock_.elapsed(), e.deadline, e.meta.kind, e.meta.id, rand_.draws());
      tracer_.sum_log2_heap += 63 - __builtin_clzll(u64(timer_.len_at_fire()));
    });
  }
  GlobalRng& rand_;
  Tracer& tracer_;
  Timer timer_;
  Clock clock_;
};

inline bool Sleep::is_elapsed() const { return handle->now_instant() >= deadline; }

u32 waker_task_id(const Waker& w);  // task.hpp

inline Poll<Unit> Sleep::poll(Context& cx) {
  if (is_elapsed()) return Poll<Unit>::done({});
  Waker w = cx.waker;  // cx.waker().clone()
  handle->add_timer_at(deadline, TimerMeta{MSIM_EV_WAKE, waker_task_id(w)}, [w]() { w.wake(); });
  return Poll<Unit>::pending();
}

/// time/mod.rs:128-140: select_biased!{ fut, sleep } — poll `fut` first, then the Sleep.
template <class F>
struct Timeout {
  using R = poll_value_t<F>;
  F fut;
  Sleep sleep;
  Poll<Result<R>> poll(Context& cx) {
    auto r = fut.poll(cx);
    if (r.ready()) return Poll<Result<R>>::done(Result<R>::Ok(std::move(*r.v)));
    if (sleep.poll(cx).ready()) return Poll<Result<R>>::done(Result<R>::Error(Err::Elapsed));
    return Poll<Result<R>>::pending();
  }
};

}  // namespace msim_oracle
