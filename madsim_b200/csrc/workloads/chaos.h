// MSIM_WL_CHAOS — the fault surface of Handle: kill / restart / pause / resume
// (reference tests task/mod.rs:858-937 kill, restart; :985-1016 pause_resume, folded into one
// supervisor schedule).  A worker node (NodeBuilder::init) counts p[1]-ms ticks while the
// supervisor kills+restarts, pauses+resumes or restarts it p[0] times, each choice drawn from the
// seed's RNG; with p[2] > 0 the worker also panics on its first p[2] starts under
// NodeBuilder::restart_on_panic (restart after gen_range(1s..10s)).  Oracle twin: oracle/workloads.hpp chaos_worker / wl_chaos_main.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Chaos {
  struct Caps {
    static constexpr int NT = 8, TQ = 32, NMSG = 1, NSOCK = 1, NREG = 1, NMBOX = 1, NNODE = 2, NPARK = 4;
    static constexpr int WQ = 0, WW = 0, WS = 2;
    // USES_KILL: can a task be dropped while queued (kill / restart / pause / abort / a returning init task)?
    static constexpr bool USES_JOIN = false;  // does any task await a JoinHandle?
    static constexpr bool USES_KILL = true;
    static constexpr bool USES_CTRLC = false;  // does any task await signal::ctrl_c()?
    static constexpr bool USES_CLOSED = false;  // does any task await Sender::closed()?
    static constexpr bool USES_SPLIT = false;   // does any task give the Sender half of its connection to another task?
    static constexpr bool USES_TIMEOUT = false;  // does any task wrap awaits in time::timeout (beyond call_timeout)?
    static constexpr bool USES_CONN2 = false;    // does any task hold two connection ends at once?
    static constexpr bool USES_PMATCH = true;    // p3: restart_on_panic_matching("0") .. ("p3 - 1") instead of restart_on_panic
    static constexpr unsigned PMATCH_BASE = 100;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    static constexpr bool COLD_SPAWN = MSIM_COLD_DEFAULT & 1, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;  // once-per-request runtime pieces out of line
    static constexpr int NCONN = 0, NCQ = 1, NACC = 1;
  };
  enum : u32 { K_MAIN = 0, K_WORKER = 1 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 PANIC_MATCH_BASE = Caps::PMATCH_BASE;  // panic!("{}", starts - 1) (task/mod.rs:965-983) carries code 100 + starts - 1
  static MSIM_HD int init_kind(int) { return K_WORKER; }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const u64 period = e.P->p[1] * NS_PER_MS;
    for (;;) switch (e.t_kind(t) * 16 + e.t_pc(t)) {
        // ---- chaos_worker: the node's init task ----
        case K_WORKER * 16 + 0:
          e.wl_s(1)++;  // starts
          if ((u64)e.wl_s(1) <= e.P->p[2]) {  // restart_on_panic (task/mod.rs:939-963)
            e.ya = e.P->p[3] ? PANIC_MATCH_BASE + e.wl_s(1) - 1 : 23u;  // matching mode: panic!("{}", starts - 1)
            return RQ_PANIC;
          }
          // fallthrough
        case K_WORKER * 16 + 1:  // loop { sleep(period); flag += 2 }
          e.t_pc(t) = 2;
          e.xa = period;
          return RQ_SLEEP_FOR;
        case K_WORKER * 16 + 2:
          e.wl_s(0) += 2;
          e.t_pc(t) = 1;
          continue;

        // ---- wl_chaos_main ----
        case K_MAIN * 16 + 0:  // node = create_node().init(chaos_worker).build()
          e.t_pc(t) = 1;
          if (e.P->p[3]) {  // .restart_on_panic_matching("0") .. ("p3 - 1") (task/mod.rs:965-983)
            e.ya = NODE_INIT | NODE_PANIC_MATCHING;
            e.yb = (1u << e.P->p[3]) - 1;
          } else {
            e.ya = NODE_INIT | (e.P->p[2] ? NODE_RESTART_ON_PANIC : 0);
            e.yb = 0;
          }
          return RQ_CREATE_NODE;
        case K_MAIN * 16 + 1:
          e.t_c(t) = (u32)e.ret;  // node id
          e.t_a(t) = 0;           // round
          e.t_pc(t) = 2;
          // fallthrough
        case K_MAIN * 16 + 2:  // for r in 0..rounds { sleep(gen_range(1..4) * period + period / 2)
          if ((u64)e.t_a(t) >= e.P->p[0]) {
            e.t_pc(t) = 9;
            e.xa = period;
            return RQ_SLEEP_FOR;
          }
          e.t_pc(t) = 3;
          e.xa = 1, e.xb = 4;
          return RQ_GEN_RANGE;
        case K_MAIN * 16 + 3:
          e.t_pc(t) = 4;
          e.xa = e.ret64 * period + period / 2;
          return RQ_SLEEP_FOR;
        case K_MAIN * 16 + 4:  // k = gen_range(0..3)
          e.t_pc(t) = 5;
          e.xa = 0, e.xb = 3;
          return RQ_GEN_RANGE;
        case K_MAIN * 16 + 5: {
          int id = (int)e.t_c(t);
          u64 k = e.ret64;
          if (k == 0) {  // kill; sleep(period); restart
            e.kill(id);
            if (!e.is_exit(id)) {
              e.ya = 20;
              return RQ_PANIC;
            }
            e.t_pc(t) = 6;
            e.xa = period;
            return RQ_SLEEP_FOR;
          }
          if (k == 1) {  // pause; sleep(3 * period); resume
            e.pause(id);
            e.t_b(t) = e.wl_s(0);
            e.t_pc(t) = 7;
            e.xa = 3 * period;
            return RQ_SLEEP_FOR;
          }
          e.restart(id);  // restart without kill: no reset_node (task/mod.rs:374-401)
          e.t_a(t)++;
          e.t_pc(t) = 2;
          continue;
        }
        case K_MAIN * 16 + 6: {
          int id = (int)e.t_c(t);
          e.restart(id);
          if (e.is_exit(id)) {
            e.ya = 21;
            return RQ_PANIC;
          }
          e.t_a(t)++;
          e.t_pc(t) = 2;
          continue;
        }
        case K_MAIN * 16 + 7:
          if (e.wl_s(0) != e.t_b(t)) {
            e.ya = 22;  // paused node made progress
            return RQ_PANIC;
          }
          e.resume((int)e.t_c(t));
          e.t_a(t)++;
          e.t_pc(t) = 2;
          continue;
        case K_MAIN * 16 + 9:
          e.w(L::W_USER0) = e.wl_s(0);
          e.w(L::W_USER1) = e.wl_s(1);
          return RQ_READY;
        default:
          e.ya = 0xbad;
          return RQ_PANIC;
      }
  }
};

}  // namespace msim
