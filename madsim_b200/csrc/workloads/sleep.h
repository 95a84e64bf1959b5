// MSIM_WL_SLEEP — block_on(async { sleep(p[0]).await })   (SURVEY.md Appendix B).
// Oracle twin: oracle/workloads.hpp wl_sleep_main.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Sleep {
  struct Caps {
    static constexpr int NT = 2, TQ = 8, NMSG = 1, NSOCK = 1, NREG = 1, NMBOX = 1, NNODE = 1, NPARK = 1;
    static constexpr int WQ = 0, WW = 0, WS = 0;
    // USES_KILL: can a task be dropped while queued (kill / restart / pause / abort / a returning init task)?
    static constexpr bool USES_JOIN = false;  // does any task await a JoinHandle?
    static constexpr bool USES_KILL = false;
    static constexpr bool USES_CTRLC = false;  // does any task await signal::ctrl_c()?
    static constexpr bool USES_CLOSED = false;  // does any task await Sender::closed()?
    static constexpr bool USES_SPLIT = false;   // does any task give the Sender half of its connection to another task?
    static constexpr bool USES_TIMEOUT = false;  // does any task wrap awaits in time::timeout (beyond call_timeout)?
    static constexpr bool USES_CONN2 = false;    // does any task hold two connection ends at once?
    static constexpr bool USES_PMATCH = false;   // does any node use restart_on_panic_matching?
    static constexpr unsigned PMATCH_BASE = 0;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    static constexpr bool COLD_SPAWN = MSIM_COLD_DEFAULT & 1, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;  // once-per-request runtime pieces out of line
    static constexpr int NCONN = 0, NCQ = 1, NACC = 1;
  };
  static constexpr u32 MAIN_KIND = 0;
  static MSIM_HD int init_kind(int) { return -1; }
  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    if (e.t_pc(t) == 0) {
      e.t_pc(t) = 1;
      e.xa = e.P->p[0];
      return RQ_SLEEP_FOR;
    }
    return RQ_READY;
  }
};

}  // namespace msim
