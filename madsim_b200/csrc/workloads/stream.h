// MSIM_WL_STREAM — SURVEY §8(f) row 1: the reliable channel layer (connect1 / accept1 / Sender /
// Receiver, net/mod.rs:337-430, net/endpoint.rs:178-274): an echo server and a client exchanging
// p[0] messages over ONE ordered connection while a supervisor task partitions the two nodes (in
// the spirit of net/tcp/mod.rs:104-179 disconnect_and_recovery).  Lost packets become delay
// (exponential back-off in the receiving stream), never loss.  Oracle twin: oracle/workloads.hpp
// stream_server / stream_echo / stream_client / stream_fault_task / wl_stream_main.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Stream {
  struct Caps {
    static constexpr int NT = 8, TQ = 32, NMSG = 2, NSOCK = 2, NREG = 2, NMBOX = 2, NNODE = 3, NPARK = 1;
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
    static constexpr int NCONN = 2, NCQ = 4, NACC = 2;
  };
  enum : u32 { K_MAIN = 0, K_SERVER, K_ECHO, K_CLIENT, K_FAULT };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 1u;  // 10.0.0.1:1
  static constexpr u32 XOR = 0x5a5a5a5au;
  static MSIM_HD int init_kind(int) { return -1; }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const Params& P = *e.P;
    for (;;) switch (e.t_kind(t) * 16 + e.t_pc(t)) {
        // ---- wl_stream_main ----
        case K_MAIN * 16 + 0:
        case K_MAIN * 16 + 1:
          e.t_pc(t)++;
          e.ya = NODE_IP;
          return RQ_CREATE_NODE;
        case K_MAIN * 16 + 2:  // node1.spawn(stream_server)
          e.t_pc(t) = 3;
          e.ya = K_SERVER, e.yb = 1, e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_MAIN * 16 + 3:  // f = node2.spawn(stream_client)
          e.t_pc(t) = 4;
          e.ya = K_CLIENT | SPAWN_JOINABLE, e.yb = 2, e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_MAIN * 16 + 4:
          e.t_a(t) = (u32)e.ret;
          e.t_pc(t) = 5;
          if (P.p[2] & 1) {  // spawn(stream_fault_task)
            e.ya = K_FAULT, e.yb = 0, e.yc = e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          // fallthrough
        case K_MAIN * 16 + 5:  // f.await
          e.t_pc(t) = 6;
          e.ya = e.t_a(t);
          return RQ_JOIN;
        case K_MAIN * 16 + 6:
          if (e.ret == 2) return panic(e, 0);
          e.w(L::W_USER0) = (u32)P.p[0];
          e.w(L::W_USER1) = e.rv0;
          e.w(L::W_USER3) = e.w(L::W_MSG_COUNT);
          return RQ_READY;

        // ---- stream_server: bind; loop { conn = accept1(); spawn(stream_echo(conn)) } ----
        case K_SERVER * 16 + 0:
          e.t_pc(t) = 1;
          e.ya = SERVER_ADDR;
          return RQ_BIND;
        case K_SERVER * 16 + 1:
          if (e.ret < 0) return panic(e, 0);
          // fallthrough
        case K_SERVER * 16 + 2:
          e.t_pc(t) = 3;
          return RQ_ACCEPT1;
        case K_SERVER * 16 + 3:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          e.ya = K_ECHO | SPAWN_EP | SPAWN_CONN, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;

        // ---- stream_echo ----
        case K_ECHO * 16 + 0:
          e.t_pc(t) = 1;
          return RQ_CH_RECV;
        case K_ECHO * 16 + 1:
          if (e.ret < 0) return RQ_READY;  // ConnectionReset: break; the connection end drops with the future
          e.t_pc(t) = 2;
          e.yb = e.rv0, e.yc = e.rv0 ^ XOR;
          return RQ_CH_SEND;
        case K_ECHO * 16 + 2:
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 0;
          continue;

        // ---- stream_client ----
        case K_CLIENT * 16 + 0:
          e.t_pc(t) = 1;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        case K_CLIENT * 16 + 1:
          e.t_pc(t) = 2;
          e.ya = mk_addr(IPK_UNSPEC, 0, 0);
          return RQ_BIND;
        case K_CLIENT * 16 + 2:
          if (e.ret < 0) return panic(e, 0);
          e.t_a(t) = 0;  // attempts
          e.t_b(t) = 0;  // i
          // fallthrough
        case K_CLIENT * 16 + 3:  // loop { attempts += 1; match ep.connect1(server) ..
          e.t_a(t)++;
          e.t_pc(t) = 4;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        case K_CLIENT * 16 + 4:
          if (e.ret == 0) {
            e.t_pc(t) = 6;
            continue;
          }
          e.t_pc(t) = 3;  // ConnectionRefused: sleep(100 ms), retry
          e.xa = 100 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_CLIENT * 16 + 6:  // for i in 0..msgs { tx.send(i)
          if ((u64)e.t_b(t) >= P.p[0]) return RQ_READY;  // output = attempts (t_a)
          e.t_pc(t) = 7;
          e.yb = e.t_b(t), e.yc = 0;
          return RQ_CH_SEND;
        case K_CLIENT * 16 + 7:  // rx.recv()
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 8;
          return RQ_CH_RECV;
        case K_CLIENT * 16 + 8:
          if (e.ret < 0) return panic(e, 0);
          if (e.rv0 != e.t_b(t) || e.rv1 != (e.t_b(t) ^ XOR)) return panic(e, 41);  // echo mismatch / reordering
          e.t_pc(t) = 9;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        case K_CLIENT * 16 + 9:
          e.t_b(t)++;
          e.t_pc(t) = 6;
          continue;

        // ---- stream_fault_task ----
        case K_FAULT * 16 + 0:
          e.t_pc(t) = 1;
          e.xa = 300, e.xb = 1500;
          return RQ_GEN_RANGE;
        case K_FAULT * 16 + 1:
          e.t_pc(t) = 2;
          e.xa = e.ret64 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_FAULT * 16 + 2:
          e.t_pc(t) = 3;
          e.xa = 0, e.xb = 3;
          return RQ_GEN_RANGE;
        case K_FAULT * 16 + 3: {
          u32 k = (u32)e.ret64;
          e.t_c(t) = k;
          if (k == 0) {
            e.clog_link(1, 2);
            e.clog_link(2, 1);
          } else {
            e.clog_node(k);
          }
          e.t_pc(t) = 4;
          e.xa = 100, e.xb = 3000;
          return RQ_GEN_RANGE;
        }
        case K_FAULT * 16 + 4:
          e.t_pc(t) = 5;
          e.xa = e.ret64 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_FAULT * 16 + 5: {
          u32 k = e.t_c(t);
          if (k == 0) {
            e.unclog_link(1, 2);
            e.unclog_link(2, 1);
          } else {
            e.unclog_node(k);
          }
          e.t_pc(t) = 0;
          continue;
        }
        default:
          return panic(e, 0xbad);
      }
  }
};

}  // namespace msim
