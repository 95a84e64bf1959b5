// MSIM_WL_PINGPONG — BASELINE config 1: 2-node ping-pong over net::Endpoint + time::sleep
// (reference test net/endpoint.rs:550-584 `connect_send_recv`, repeated p[0] times with the
// un-simulated tokio Barrier replaced by a 1 s sleep).  Oracle twin: oracle/workloads.hpp
// pp_server / pp_client / wl_pingpong_main: every `case` is the code between two await points of
// that async code; the awaited primitive itself is a request executed by the engine.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_PingPong {
  struct Caps {
    static constexpr int NT = 4, TQ = 16, NMSG = 4, NSOCK = 2, NREG = 2, NMBOX = 2, NNODE = 3, NPARK = 1;
    static constexpr int WQ = 0, WW = 0, WS = 0;
    // USES_KILL: can a task be dropped while queued (kill / restart / pause / abort / a returning init task)?
    static constexpr bool USES_JOIN = true;  // does any task await a JoinHandle?
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
  enum : u32 { K_MAIN = 0, K_SERVER = 1, K_CLIENT = 2 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 PING = 0x676e6970u, PONG = 0x676e6f70u;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 1u;  // 10.0.0.1:1
  static MSIM_HD int init_kind(int) { return -1; }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    u32 pc = e.t_pc(t);
    switch (e.t_kind(t) * 8 + pc) {
      // ---- wl_pingpong_main ----
      case K_MAIN * 8 + 0:  // node1 = create_node().ip(10.0.0.1).build()
      case K_MAIN * 8 + 1:  // node2 = ...
        e.t_pc(t) = pc + 1;
        e.ya = NODE_IP;
        return RQ_CREATE_NODE;
      case K_MAIN * 8 + 2:  // node1.spawn(pp_server)
        e.t_pc(t) = 3;
        e.ya = K_SERVER, e.yb = 1, e.yc = e.yd = 0, e.xa = e.xb = 0;
        return RQ_SPAWN;
      case K_MAIN * 8 + 3:  // let f = node2.spawn(pp_client)
        e.t_pc(t) = 4;
        e.ya = K_CLIENT | SPAWN_JOINABLE, e.yb = 2, e.yc = e.yd = 0, e.xa = e.xb = 0;
        return RQ_SPAWN;
      case K_MAIN * 8 + 4:  // f.await
        e.t_pc(t) = 5;
        e.ya = (u32)e.ret;
        return RQ_JOIN;
      case K_MAIN * 8 + 5:
        if (e.ret == 2) {
          e.ya = 14;  // .unwrap() on a JoinError
          return RQ_PANIC;
        }
        e.w(L::W_USER0) = e.rv0;
        return RQ_READY;

      // ---- pp_server ----
      case K_SERVER * 8 + 0:  // Endpoint::bind(addr)
        e.t_pc(t) = 1;
        e.ya = SERVER_ADDR;
        return RQ_BIND;
      case K_SERVER * 8 + 1:
        if (e.ret < 0) {
          e.ya = 10;
          return RQ_PANIC;
        }
        // fallthrough
      case K_SERVER * 8 + 2:  // loop { recv_from(1)
        e.t_pc(t) = 3;
        e.xa = 1;
        return RQ_RECV;
      case K_SERVER * 8 + 3:  // send_to(from, 1, "pong") }
        if (e.rv0 != PING) {
          e.ya = 11;
          return RQ_PANIC;
        }
        e.t_pc(t) = 2;
        e.ya = e.rfrom, e.xa = 1, e.xb = 0, e.yb = PONG, e.yc = 0;
        return RQ_SEND;

      // ---- pp_client ----
      case K_CLIENT * 8 + 0:  // sleep(1 s)
        e.t_pc(t) = 1;
        e.xa = NS_PER_SEC;
        return RQ_SLEEP_FOR;
      case K_CLIENT * 8 + 1:  // Endpoint::connect = bind("0.0.0.0:0")
        e.t_pc(t) = 2;
        e.ya = mk_addr(IPK_UNSPEC, 0, 0);
        return RQ_BIND;
      case K_CLIENT * 8 + 2:
        if (e.ret < 0) {
          e.ya = 12;
          return RQ_PANIC;
        }
        e.t_a(t) = 0;  // done
        e.t_b(t) = 0;  // r
        // fallthrough
      case K_CLIENT * 8 + 3:  // for r in 0..rounds { ep.send(1, "ping")
        if ((u64)e.t_b(t) >= e.P->p[0]) return RQ_READY;  // output = t_a; the Endpoint drops with the future
        e.t_pc(t) = 4;
        e.ya = SERVER_ADDR, e.xa = 1, e.xb = 0, e.yb = PING, e.yc = 0;
        return RQ_SEND;
      case K_CLIENT * 8 + 4:  // ep.recv(1)
        e.t_pc(t) = 5;
        e.xa = 1;
        return RQ_RECV;
      case K_CLIENT * 8 + 5:
        if (e.rfrom != SERVER_ADDR) {
          e.ya = 15;  // "receive a message but not from the connected address"
          return RQ_PANIC;
        }
        if (e.rv0 != PONG) {
          e.ya = 13;
          return RQ_PANIC;
        }
        e.t_a(t) = e.t_a(t) + 1;
        e.t_pc(t) = 6;
        e.xa = e.P->p[1];  // sleep(think)
        return RQ_SLEEP_FOR;
      case K_CLIENT * 8 + 6:
        e.t_b(t) = e.t_b(t) + 1;
        e.t_pc(t) = 3;
        if ((u64)e.t_b(t) >= e.P->p[0]) return RQ_READY;
        e.t_pc(t) = 4;
        e.ya = SERVER_ADDR, e.xa = 1, e.xb = 0, e.yb = PING, e.yc = 0;
        return RQ_SEND;
    }
    e.ya = 0xbad;
    return RQ_PANIC;
  }
};

}  // namespace msim
