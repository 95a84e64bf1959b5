// MSIM_WL_SIGNAL — SURVEY §8(f) row 4, the rest of the fault / time surface:
//   time::interval + MissedTickBehavior   madsim/src/sim/time/interval.rs:39-169
//   signal::ctrl_c                        madsim/src/sim/signal.rs:4-8           (RQ_CTRL_C / W_CTRLC)
//   Handle::send_ctrl_c                   madsim/src/sim/task/mod.rs:427-441     (Engine::send_ctrl_c)
// A worker node's init task ticks an Interval and works gen_range(0..2 periods) after every tick
// (so ticks are missed and the MissedTickBehavior decides the schedule), optionally with a second
// task catching ctrl-c; the supervisor sends ctrl-c or restarts the node.  Interval::poll_tick is
// a Sleep that is re-armed with Sleep::reset — the raw, un-clamped deadline (RQ_SLEEP_RAW) — after
// the first, clamped sleep_until(now).  Oracle twin: oracle/workloads_signal.hpp.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Signal {
  enum : int { S_TICKS = 0, S_CAUGHT, S_STARTS, S_ROUND, S_HAVE_PREV, S_END };
  enum : int { FQ_DEADLINE = 0, FQ_FIRST, FQ_PREV, FQ_N };  // wl_q of the worker node: the Interval + the checker
  struct Caps {
    static constexpr int NT = 8, TQ = 32, NMSG = 1, NSOCK = 1, NREG = 1, NMBOX = 1, NNODE = 2, NPARK = 1;
    static constexpr int WQ = FQ_N, WW = 0, WS = S_END;
    static constexpr bool USES_JOIN = false;
    static constexpr bool USES_KILL = true;
    static constexpr bool USES_CTRLC = true;  // does any task await signal::ctrl_c()?
    static constexpr bool USES_CLOSED = false;  // does any task await Sender::closed()?
    static constexpr bool USES_SPLIT = false;   // does any task give the Sender half of its connection to another task?
    static constexpr bool USES_TIMEOUT = false;  // does any task wrap awaits in time::timeout (beyond call_timeout)?
    static constexpr bool USES_CONN2 = false;    // does any task hold two connection ends at once?
    static constexpr bool USES_PMATCH = false;   // does any node use restart_on_panic_matching?
    static constexpr unsigned PMATCH_BASE = 0;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    static constexpr bool COLD_SPAWN = MSIM_COLD_DEFAULT & 1, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;  // once-per-request runtime pieces out of line
    static constexpr int NCONN = 0, NCQ = 1, NACC = 1;
  };
  enum : u32 { K_MAIN = 0, K_WORKER, K_CATCH };
  enum : u32 { BURST = 0, DELAY = 1, SKIP = 2 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static MSIM_HD int init_kind(int) { return K_WORKER; }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const Params& P = *e.P;
    const u64 period = P.p[1] * NS_PER_MS;
    const u32 behavior = (u32)P.p[2];
    for (;;) switch (e.t_kind(t) * 16 + e.t_pc(t)) {
        // ---- signal_worker: the node's init task ----
        case K_WORKER * 16 + 0:
          e.wl_s(S_STARTS)++;
          e.t_pc(t) = 1;
          if (P.p[3] & 1) {  // spawn(signal_catcher)
            e.ya = K_CATCH, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          // fallthrough
        case K_WORKER * 16 + 1:  // interval(period): delay = sleep_until(now) — clamped to now + 1 ms
          e.wl_q(FQ_DEADLINE, e.t_node(t)) = e.sleep_deadline(e.now);
          e.wl_s(S_HAVE_PREV) = 0;
          e.t_pc(t) = 2;
          // fallthrough
        case K_WORKER * 16 + 2:  // iv.tick(): poll the delay
          e.t_pc(t) = 3;
          e.xa = e.wl_q(FQ_DEADLINE, e.t_node(t));
          return RQ_SLEEP_RAW;
        case K_WORKER * 16 + 3: {  // poll_tick after the delay: re-arm, yield the scheduled instant
          int n = e.t_node(t);
          u64 at = e.wl_q(FQ_DEADLINE, n), now = e.now, next;
          if (now > at + 5 * NS_PER_MS) {
            next = behavior == BURST ? at + period : behavior == DELAY ? now + period : now + period - ((now - at) % period);
          } else {
            next = at + period;
          }
          e.wl_q(FQ_DEADLINE, n) = next;  // delay.reset(next)
          if (now < at) return panic(e, 80);
          if (!e.wl_s(S_HAVE_PREV)) {
            e.wl_q(FQ_FIRST, n) = at;
          } else {
            u64 first = e.wl_q(FQ_FIRST, n), prev = e.wl_q(FQ_PREV, n);
            if (behavior != DELAY && (at - first) % period != 0) return panic(e, 81);
            if (behavior == BURST && at != prev + period) return panic(e, 82);
            if (at < prev + period) return panic(e, 83);
          }
          e.wl_q(FQ_PREV, n) = at;
          e.wl_s(S_HAVE_PREV) = 1;
          e.wl_s(S_TICKS)++;
          e.t_pc(t) = 4;
          e.xa = 0, e.xb = 2 * period;
          return RQ_GEN_RANGE;
        }
        case K_WORKER * 16 + 4:  // the work between ticks
          e.t_pc(t) = 2;
          e.xa = e.ret64;
          return RQ_SLEEP_FOR;

        // ---- signal_catcher: loop { ctrl_c().await; caught += 1 } ----
        case K_CATCH * 16 + 0:
          e.t_pc(t) = 1;
          return RQ_CTRL_C;
        case K_CATCH * 16 + 1:
          e.wl_s(S_CAUGHT)++;
          e.t_pc(t) = 0;
          continue;

        // ---- wl_signal_main ----
        case K_MAIN * 16 + 0:
          e.t_pc(t) = 1;
          e.ya = NODE_INIT;
          return RQ_CREATE_NODE;
        case K_MAIN * 16 + 1:
          e.t_c(t) = (u32)e.ret;  // node id
          e.t_pc(t) = 2;
          // fallthrough
        case K_MAIN * 16 + 2:  // for r in 0..rounds { sleep(gen_range(1..4) * 3 * period + period / 2)
          if ((u64)e.wl_s(S_ROUND) >= P.p[0]) {
            e.t_pc(t) = 6;
            e.xa = period;
            return RQ_SLEEP_FOR;
          }
          e.t_pc(t) = 3;
          e.xa = 1, e.xb = 4;
          return RQ_GEN_RANGE;
        case K_MAIN * 16 + 3:
          e.t_pc(t) = 4;
          e.xa = e.ret64 * 3 * period + period / 2;
          return RQ_SLEEP_FOR;
        case K_MAIN * 16 + 4:  // k = gen_range(0..3)
          e.t_pc(t) = 5;
          e.xa = 0, e.xb = 3;
          return RQ_GEN_RANGE;
        case K_MAIN * 16 + 5:
          if (e.ret64 < 2)
            e.send_ctrl_c((int)e.t_c(t));
          else
            e.restart((int)e.t_c(t));
          e.wl_s(S_ROUND)++;
          e.t_pc(t) = 2;
          continue;
        case K_MAIN * 16 + 6:
          e.w(L::W_USER0) = e.wl_s(S_TICKS);
          e.w(L::W_USER1) = e.wl_s(S_CAUGHT);
          e.w(L::W_USER2) = e.wl_s(S_STARTS);
          e.w(L::W_USER3) = e.is_exit((int)e.t_c(t));
          return RQ_READY;
        default:
          return panic(e, 0xbad);
      }
  }
};

}  // namespace msim
