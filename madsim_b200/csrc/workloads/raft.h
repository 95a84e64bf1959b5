// MSIM_WL_RAFT — BASELINE configs 2 and 3: 5-node Raft leader election (terms, votes,
// heartbeats over rpc call_timeout / add_rpc_handler) with an optional supervisor fault schedule
// (NetSim::clog_node / clog_link partitions, time::advance clock jumps).
//
// MadRaft is not part of the reference tree (README.md:78 only names it), so the
// program is authored here on the reference's API.  Oracle twin: oracle/workloads.hpp raft_*.
// Every `case` below is the code between two await points of that async code; the awaited
// primitive (Endpoint::bind, call_timeout, recv_from_raw, send_to_raw, sleep, spawn, gen_range …)
// is returned as a REQUEST and executed by the engine (engine.h poll_task), which implements its
// poll-level behaviour once for all tasks:
//   add_rpc_handler           net/rpc.rs:134-166 (loop task K_LOOP_* + one task K_REQ_* per request)
//   call / call_timeout       net/rpc.rs:96-131, time/mod.rs:128-140
#pragma once
#include "../engine/engine.h"

namespace msim {

// rpc.rs:71-81 hash33 — Request::ID = hash_str(module_path + name)
constexpr u64 rpc_hash_str(const char* s) {
  u64 h = 0;
  for (; *s; ++s) h = h * 33 + (u64)(unsigned char)*s;
  return h;
}

struct WL_Raft {
  struct Caps {
    // split-array timer heap: 16-byte records (Caps::HEAP16) measured 2 % fewer instructions but no faster on this
    // program (profiles/r2k_footprint_experiments.txt), and 176 x 12 B buys more timers than 132 x 16 B.  Observed peaks: 60 timers /
    // 37 tasks over 8192 seeds of config 2, 80 / 47 over the 65 536 seeds of config 3.
    static constexpr bool HEAP16 = false;
    static constexpr bool PORT_FIRST = true;  // socket lookup tests the port first: +2.7 % here, a placement effect (profiles/r2ad_port_first.txt)
    static constexpr int NT = 64, TQ = 176, NMSG = 48, NSOCK = 6, NREG = 12, NMBOX = 16, NNODE = 6, NPARK = 2;
    static constexpr int WQ = 1, WW = 6, WS = 1;
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
  enum : u32 {
    K_MAIN = 0, K_FAULT, K_NODE, K_LOOP_VOTE, K_LOOP_APPEND, K_REQ_VOTE, K_REQ_APPEND, K_VOTE, K_APPEND, K_LEADER
  };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr int PCS = 16;                                           // pc values per kind
  enum : int { F_TERM = 0, F_VOTED, F_ROLE, F_VOTES, F_LEADER, F_VALID };  // wl_w fields
  enum : int { FQ_LAST_HB = 0 };                                           // wl_q fields
  enum : u32 { FOLLOWER = 0, CANDIDATE = 1, LEADER = 2 };
  static constexpr u64 TAG_VOTE = rpc_hash_str("raft::RequestVote");
  static constexpr u64 TAG_APPEND = rpc_hash_str("raft::AppendEntries");
  static MSIM_HD u32 raft_addr(u32 i) { return mk_addr(IPK_NODE, i, 1); }
  /// NodeBuilder::init(|| raft_node_main(c, i)) — also what Handle::restart re-runs
  static MSIM_HD int init_kind(int) { return K_NODE; }

  template <class E>
  static MSIM_HD void step_down(E& e, u32 me, u32 term) {
    e.wl_w(F_TERM, me) = term;
    e.wl_w(F_ROLE, me) = FOLLOWER;
    e.wl_w(F_VOTED, me) = 0;
  }
  /// spawn(kind) on this node with a clone of the node's Endpoint; locals a, b
  template <class E>
  static MSIM_HD int spawn_ep(E& e, u32 kind, u32 node, u32 a, u32 b2, u32 c = 0, u64 dl1 = 0) {
    e.ya = kind | SPAWN_EP, e.yb = node, e.yc = a, e.yd = b2, e.xa = c, e.xb = dl1;
    return RQ_SPAWN;
  }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }
  /// outputs of wl_raft_main (computed by main's last poll)
  template <class E>
  static MSIM_HD void outputs(E& e) {
    using L = typename E::L;
    u32 n = e.P->n_nodes, term = 0, leader = 0, best = 0;
    MSIM_NOUNROLL
    for (u32 a = 1; a <= n; a++) {
      if (!e.wl_w(F_VALID, a)) continue;
      if (e.wl_w(F_TERM, a) > term) term = e.wl_w(F_TERM, a);
      if (e.wl_w(F_ROLE, a) == LEADER && e.wl_w(F_TERM, a) >= best) {
        best = e.wl_w(F_TERM, a);
        leader = a;
      }
    }
    e.w(L::W_USER0) = leader;
    e.w(L::W_USER1) = term;
    e.w(L::W_USER2) = e.wl_s(0);
    e.w(L::W_USER3) = e.w(L::W_MSG_COUNT);
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    const Params& P = *e.P;
    const u32 n = P.n_nodes;
    const u32 me = e.t_node(t);
    for (;;) switch (e.t_kind(t) * PCS + e.t_pc(t)) {
        // =============== supervisor: wl_raft_main (node 0) ===============
        case K_MAIN * PCS + 0:  // for i in 1..=n { create_node().ip(10.0.0.i).init(raft_node_main).build() }
          if (e.t_a(t) < n) {
            e.t_a(t)++;
            e.ya = NODE_IP | NODE_INIT;
            return RQ_CREATE_NODE;
          }
          e.t_a(t) = 0;  // tick
          e.t_pc(t) = 1;
          if ((u32)P.p[1]) {  // if fault_mask { spawn(raft_fault_task) }
            e.ya = K_FAULT, e.yb = 0, e.yc = e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          // fallthrough
        case K_MAIN * PCS + 1:  // for tick in 1..=ticks { sleep(100 ms)
          if ((u64)e.t_a(t) >= P.p[0] / 100) {
            outputs(e);
            return RQ_READY;
          }
          e.t_pc(t) = 2;
          e.xa = 100 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_MAIN * PCS + 2: {
          u32 tick = ++e.t_a(t);
          bool any_leader = false;
          MSIM_NOUNROLL
          for (u32 a = 1; a <= n; a++) {
            if (!e.wl_w(F_VALID, a) || e.wl_w(F_ROLE, a) != LEADER) continue;
            any_leader = true;
            MSIM_NOUNROLL
            for (u32 b2 = a + 1; b2 <= n; b2++)
              if (e.wl_w(F_VALID, b2) && e.wl_w(F_ROLE, b2) == LEADER && e.wl_w(F_TERM, b2) == e.wl_w(F_TERM, a))
                return panic(e, 1);  // two leaders in one term
          }
          if ((u32)P.p[1] == 0 && tick == 20 && !any_leader) return panic(e, 2);  // no leader elected by 2 s
          e.t_pc(t) = 1;
          continue;
        }

        // =============== raft_fault_task (config 3) ===============
        case K_FAULT * PCS + 0:  // loop { sleep(gen_range(500..2000) ms)
          e.t_pc(t) = 1;
          e.xa = 500, e.xb = 2000;
          return RQ_GEN_RANGE;
        case K_FAULT * PCS + 1:
          e.t_pc(t) = 2;
          e.xa = e.ret64 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_FAULT * PCS + 2: {  // k = kinds[gen_range(0..nk)]
          u32 nk = 0, mask = (u32)P.p[1];
          for (u32 k = 0; k < 3; k++) nk += (mask >> k) & 1;
          e.t_pc(t) = 3;
          e.xa = 0, e.xb = nk;
          return RQ_GEN_RANGE;
        }
        case K_FAULT * PCS + 3: {
          u32 mask = (u32)P.p[1], pick = (u32)e.ret64, k = 0;
          for (;; k++)
            if ((mask >> k) & 1) {
              if (pick == 0) break;
              pick--;
            }
          e.t_c(t) = k;
          e.t_pc(t) = 4;
          if (k == 2) {  // advance(gen_range(0..500) ms)
            e.xa = 0, e.xb = 500;
          } else {  // i = gen_range(1..n+1)
            e.xa = 1, e.xb = n + 1;
          }
          return RQ_GEN_RANGE;
        }
        case K_FAULT * PCS + 4: {
          u32 k = e.t_c(t);
          if (k == 2) {
            e.t_pc(t) = 0;
            e.xa = e.ret64 * NS_PER_MS;
            return RQ_ADVANCE;
          }
          e.t_a(t) = (u32)e.ret64;  // i
          if (k == 0) {             // clog_node(i); sleep(gen_range(200..1000) ms)
            e.clog_node(e.t_a(t));
            e.t_pc(t) = 6;
            e.xa = 200, e.xb = 1000;
          } else {  // j = gen_range(1..n); if j >= i { j += 1 }
            e.t_pc(t) = 5;
            e.xa = 1, e.xb = n;
          }
          return RQ_GEN_RANGE;
        }
        case K_FAULT * PCS + 5: {
          u32 i = e.t_a(t), j = (u32)e.ret64;
          if (j >= i) j++;
          e.t_b(t) = j;
          e.clog_link(i, j);
          e.clog_link(j, i);
          e.t_pc(t) = 6;
          e.xa = 200, e.xb = 1000;
          return RQ_GEN_RANGE;
        }
        case K_FAULT * PCS + 6:
          e.t_pc(t) = 7;
          e.xa = e.ret64 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_FAULT * PCS + 7:
          if (e.t_c(t) == 0) {
            e.unclog_node(e.t_a(t));
          } else {
            e.unclog_link(e.t_a(t), e.t_b(t));
            e.unclog_link(e.t_b(t), e.t_a(t));
          }
          e.t_pc(t) = 0;
          continue;

        // =============== raft_node_main: bind, serve RPCs, election ticker ===============
        case K_NODE * PCS + 0:
          e.t_pc(t) = 1;
          e.ya = raft_addr(me);
          return RQ_BIND;
        case K_NODE * PCS + 1:
          if (e.ret < 0) return panic(e, 30);
          e.wl_w(F_TERM, me) = 0;
          e.wl_w(F_VOTED, me) = 0;
          e.wl_w(F_ROLE, me) = FOLLOWER;
          e.wl_w(F_VOTES, me) = 0;
          e.wl_w(F_LEADER, me) = 0;
          e.wl_q(FQ_LAST_HB, me) = e.now;
          e.wl_w(F_VALID, me) = 1;
          e.t_pc(t) = 2;
          return spawn_ep(e, K_LOOP_VOTE, me, 0, 0);  // ep.add_rpc_handler::<RequestVote>
        case K_NODE * PCS + 2:
          e.t_pc(t) = 3;
          return spawn_ep(e, K_LOOP_APPEND, me, 0, 0);  // ep.add_rpc_handler::<AppendEntries>
        case K_NODE * PCS + 3:  // loop { timeout = thread_rng().gen_range(lo..hi) ms
          e.t_pc(t) = 4;
          e.xa = P.p[5], e.xb = P.p[6];
          return RQ_GEN_RANGE;
        case K_NODE * PCS + 4:  // sleep(timeout)
          e.t_a(t) = (u32)e.ret64;
          e.t_pc(t) = 5;
          e.xa = e.ret64 * NS_PER_MS;
          return RQ_SLEEP_FOR;
        case K_NODE * PCS + 5: {
          u64 to = (u64)e.t_a(t) * NS_PER_MS;
          if (e.wl_w(F_ROLE, me) != LEADER && e.now - e.wl_q(FQ_LAST_HB, me) >= to) {
            // start an election
            e.t_b(t) = ++e.wl_w(F_TERM, me);  // term0
            e.wl_w(F_ROLE, me) = CANDIDATE;
            e.wl_w(F_VOTED, me) = me;
            e.wl_w(F_VOTES, me) = 1;
            e.wl_q(FQ_LAST_HB, me) = e.now;
            e.wl_s(0)++;  // elections
            e.t_c(t) = 1;  // next peer
            e.t_pc(t) = 6;
            continue;
          }
          e.t_pc(t) = 3;
          continue;
        }
        case K_NODE * PCS + 6: {  // for peer in others { spawn(raft_vote_task(peer, term0)) }
          u32 peer = e.t_c(t);
          if (peer == me) peer++;
          if (peer > n) {
            e.t_pc(t) = 3;
            continue;
          }
          e.t_c(t) = peer + 1;
          return spawn_ep(e, K_VOTE, me, peer, e.t_b(t));
        }

        // =============== add_rpc_handler: receive loop (rpc.rs:143-165) ===============
        case K_LOOP_VOTE * PCS + 0:
        case K_LOOP_APPEND * PCS + 0:  // loop { (data, from) = recv_from_raw(R::ID)
          e.t_pc(t) = 1;
          e.xa = e.t_kind(t) == K_LOOP_VOTE ? TAG_VOTE : TAG_APPEND;
          return RQ_RECV;
        case K_LOOP_VOTE * PCS + 1:
        case K_LOOP_APPEND * PCS + 1:  // spawn(async { rsp = f(req).await; send_to_raw(from, rsp_tag, rsp) }) }
          e.t_pc(t) = 0;
          return spawn_ep(e, e.t_kind(t) == K_LOOP_VOTE ? K_REQ_VOTE : K_REQ_APPEND, me, e.rv0, e.rv1, e.rfrom, e.rw);

        // =============== one task per request: handler body, then the reply ===============
        case K_REQ_VOTE * PCS + 0:
        case K_REQ_APPEND * PCS + 0: {  // rsp_future.await — the handler runs in THIS task's first poll
          u32 term = e.t_a(t), who = e.t_b(t), ok;
          if (e.t_kind(t) == K_REQ_VOTE) {  // raft_on_request_vote
            if (term > e.wl_w(F_TERM, me)) step_down(e, me, term);
            u32 voted = e.wl_w(F_VOTED, me);
            ok = term == e.wl_w(F_TERM, me) && ((u32)P.p[2] == 1 || voted == 0 || voted == who);
            if (ok) {
              e.wl_w(F_VOTED, me) = who;
              e.wl_q(FQ_LAST_HB, me) = e.now;
            }
          } else {  // raft_on_append
            ok = term >= e.wl_w(F_TERM, me);
            if (ok) {
              if (term > e.wl_w(F_TERM, me)) {
                e.wl_w(F_TERM, me) = term;
                e.wl_w(F_VOTED, me) = 0;
              }
              e.wl_w(F_ROLE, me) = FOLLOWER;
              e.wl_w(F_LEADER, me) = who;
              e.wl_q(FQ_LAST_HB, me) = e.now;
            }
          }
          e.t_pc(t) = 1;  // net.send_to_raw(from, rsp_tag, rsp)
          e.ya = e.t_c(t), e.xa = e.t_dl1(t), e.xb = 0, e.yb = e.wl_w(F_TERM, me), e.yc = ok;
          return RQ_SEND;
        }
        case K_REQ_VOTE * PCS + 1:
        case K_REQ_APPEND * PCS + 1:
          return RQ_READY;

        // =============== raft_vote_task / raft_append_task ===============
        case K_VOTE * PCS + 0:
        case K_APPEND * PCS + 0:  // rsp = ep.call_timeout(addr(peer), Req{term0, me}, rpc_timeout)
          e.t_pc(t) = 1;
          e.ya = raft_addr(e.t_a(t));
          e.xa = e.t_kind(t) == K_VOTE ? TAG_VOTE : TAG_APPEND;
          e.yb = e.t_b(t), e.yc = me;
          e.xb = P.p[3] * NS_PER_MS;
          return RQ_CALL;
        case K_VOTE * PCS + 1:
        case K_APPEND * PCS + 1: {
          if (e.ret != CALL_OK) return RQ_READY;
          u32 term0 = e.t_b(t), v0 = e.rv0, v1 = e.rv1;
          if (v0 > e.wl_w(F_TERM, me)) {
            step_down(e, me, v0);
          } else if (e.t_kind(t) == K_VOTE && e.wl_w(F_ROLE, me) == CANDIDATE && e.wl_w(F_TERM, me) == term0 && v1) {
            u32 votes = ++e.wl_w(F_VOTES, me);
            if (votes == n / 2 + 1) {
              e.wl_w(F_ROLE, me) = LEADER;
              e.wl_w(F_LEADER, me) = me;
              e.t_pc(t) = 2;
              return spawn_ep(e, K_LEADER, me, term0, 0);  // spawn(raft_leader_loop)
            }
          }
          return RQ_READY;
        }
        case K_VOTE * PCS + 2:
          return RQ_READY;

        // =============== raft_leader_loop ===============
        case K_LEADER * PCS + 0:  // while role == Leader && term == term0 {
          if (!(e.wl_w(F_ROLE, me) == LEADER && e.wl_w(F_TERM, me) == e.t_a(t))) return RQ_READY;
          e.t_c(t) = 1;
          e.t_pc(t) = 1;
          // fallthrough
        case K_LEADER * PCS + 1: {  // for peer in others { spawn(raft_append_task(peer, term0)) }
          u32 peer = e.t_c(t);
          if (peer == me) peer++;
          if (peer > n) {  // sleep(heartbeat) }
            e.t_pc(t) = 0;
            e.xa = P.p[4] * NS_PER_MS;
            return RQ_SLEEP_FOR;
          }
          e.t_c(t) = peer + 1;
          return spawn_ep(e, K_APPEND, me, peer, e.t_a(t));
        }
        default:
          return panic(e, 0xbad);
      }
  }
};

}  // namespace msim
