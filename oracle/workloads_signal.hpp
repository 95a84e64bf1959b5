// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// MSIM_WL_SIGNAL — SURVEY §8(f) row 4, the rest of the fault / time surface:
//   time::interval + MissedTickBehavior   madsim/src/sim/time/interval.rs:39-169
//   signal::ctrl_c                        madsim/src/sim/signal.rs:4-8
//   Handle::send_ctrl_c                   madsim/src/sim/task/mod.rs:427-441
// A worker node (NodeBuilder::init) ticks an Interval and "works" for gen_range(0..2 periods)
// after every tick, so roughly half of the ticks are missed and the chosen MissedTickBehavior
// decides the schedule; optionally a second task on the node catches ctrl-c.  The supervisor
// sends ctrl-c (delivered if a handler was ever installed on this incarnation, else it kills the
// node — signal.rs:23-66 ctrl_c_kill / ctrl_c_catch) or restarts the node.
#pragma once
#include <array>

#include "rt/runtime.hpp"

namespace msim_oracle {

struct SignalParams {
  u64 rounds = 8;
  u64 period_ms = 100;
  u32 behavior = 0;  // 0 Burst, 1 Delay, 2 Skip
  u32 flags = 1;     // 1: the worker installs a ctrl-c handler
};
struct SignalState {
  u32 ticks = 0, caught = 0, starts = 0;
};
inline Async<Unit> signal_catcher(std::shared_ptr<SignalState> st) {
  for (;;) {
    co_await ctrl_c();
    st->caught++;
  }
}
inline Async<Unit> signal_worker(std::shared_ptr<SignalState> st, SignalParams p) {
  st->starts++;
  if (p.flags & 1) spawn(signal_catcher(st));
  u64 period = p.period_ms * MS;
  Interval iv = interval(period);
  iv.set_missed_tick_behavior(MissedTickBehavior(p.behavior));
  GlobalRng& rng = thread_rng();
  u64 first = 0, prev = 0;
  bool have_prev = false;
  for (;;) {
    u64 at = co_await iv.tick();
    if (now() < at) panic("tick completed before its scheduled instant", 80);
    if (!have_prev) {
      first = at;
    } else {
      // Burst and Skip stay on the grid of the first tick; Burst never skips a slot; Delay keeps the period
      if (p.behavior != 1 && (at - first) % period != 0) panic("tick off the period grid", 81);
      if (p.behavior == 0 && at != prev + period) panic("Burst skipped a slot", 82);
      if (at < prev + period) panic("ticks closer than one period", 83);
    }
    prev = at, have_prev = true;
    st->ticks++;
    co_await sleep(rng.gen_range_u64(0, 2 * period));  // the work between ticks
  }
}
inline Async<u32> wl_signal_main(SignalParams p, std::array<u32, 4>* user) {
  auto& rt = current_runtime();
  auto st = std::make_shared<SignalState>();
  u64 period = p.period_ms * MS;
  auto nb = rt.create_node();
  nb.init([st, p]() { return signal_worker(st, p); });
  u64 id = nb.build().id();
  GlobalRng& rng = thread_rng();
  for (u64 r = 0; r < p.rounds; r++) {
    co_await sleep(rng.gen_range_u64(1, 4) * 3 * period + period / 2);
    u64 k = rng.gen_range_u64(0, 3);
    if (k < 2)
      rt.send_ctrl_c(id);
    else
      rt.restart(id);
  }
  co_await sleep(period);
  (*user)[0] = st->ticks;
  (*user)[1] = st->caught;
  (*user)[2] = st->starts;
  (*user)[3] = rt.is_exit(id);
  co_return 0;
}

}  // namespace msim_oracle
