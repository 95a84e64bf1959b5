// MSIM_WL_SLEEP — block_on(async { sleep(p[0]).await })   (SURVEY.md Appendix B).
// Oracle twin: oracle/workloads.hpp wl_sleep_main.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Sleep {
  struct Caps {
    static constexpr int NT = 2, TQ = 8, NMSG = 1, NSOCK = 1, NREG = 1, NMBOX = 1, NNODE = 1, NPARK = 1;
    static constexpr int WQ = 0, WW = 0, WS = 0;
  };
  static constexpr u32 MAIN_KIND = 0;
  template <class E>
  static MSIM_HD int poll(E& e, int t) {
    switch (e.t_pc(t)) {
      case 0:
        e.t_dl0(t) = e.sleep_deadline(e.now + e.P->p[0]);
        e.t_pc(t) = 1;
        // fallthrough
      case 1:
        if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
        return POLL_READY;
    }
    return e.panic(0xbad);
  }
  template <class E>
  static MSIM_HD void on_drop(E&, int) {}
  template <class E>
  static MSIM_HD void node_init(E&, int) {}
  template <class E>
  static MSIM_HD void finish(E&) {}
};

}  // namespace msim
