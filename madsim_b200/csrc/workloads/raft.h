// MSIM_WL_RAFT — BASELINE configs 2 and 3: 5-node Raft leader election (terms, votes,
// heartbeats over rpc call_timeout / add_rpc_handler) with an optional supervisor fault schedule
// (NetSim::clog_node / clog_link partitions, time::advance clock jumps).
//
// MadRaft is not part of the reference tree (README.md:78 only names it), so the
// program is authored here on the reference's API.  Oracle twin: oracle/workloads.hpp raft_* —
// every `case` below is one await point of that async code, expanded into the poll-level
// behaviour of SURVEY.md §3.4:
//   Endpoint::bind            net/endpoint.rs:23-36, net/mod.rs:446-480 (rand_delay, then bind)
//   add_rpc_handler           net/rpc.rs:134-166 (loop task + one task per request)
//   call / call_timeout       net/rpc.rs:96-131, time/mod.rs:128-140 (select_biased: call first, then Sleep)
//   send_to_raw / recv_from_raw  net/endpoint.rs:120-149 (rand_delay before send, AFTER receipt)
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
    static constexpr int NT = 64, TQ = 176, NMSG = 48, NSOCK = 6, NREG = 12, NMBOX = 16, NNODE = 6, NPARK = 2;
    static constexpr int WQ = 1, WW = 6, WS = 1;
  };
  enum : u32 {
    K_MAIN = 0, K_FAULT, K_NODE, K_LOOP_VOTE, K_LOOP_APPEND, K_REQ_VOTE, K_REQ_APPEND, K_VOTE, K_LEADER, K_APPEND
  };
  static constexpr u32 MAIN_KIND = K_MAIN;
  enum : int { F_TERM = 0, F_VOTED, F_ROLE, F_VOTES, F_LEADER, F_VALID };  // wl_w fields
  enum : int { FQ_LAST_HB = 0 };                                           // wl_q fields
  enum : u32 { FOLLOWER = 0, CANDIDATE = 1, LEADER = 2 };
  static constexpr u64 TAG_VOTE = rpc_hash_str("raft::RequestVote");
  static constexpr u64 TAG_APPEND = rpc_hash_str("raft::AppendEntries");
  static MSIM_HD u32 raft_addr(u32 i) { return mk_addr(IPK_NODE, i, 1); }

  // ---- parameters (msim_workload_desc.p) ----
  template <class E> static MSIM_HD u32 n_nodes(E& e) { return e.P->n_nodes; }
  template <class E> static MSIM_HD u64 dur_ms(E& e) { return e.P->p[0]; }
  template <class E> static MSIM_HD u32 fault_mask(E& e) { return (u32)e.P->p[1]; }
  template <class E> static MSIM_HD u32 bug(E& e) { return (u32)e.P->p[2]; }
  template <class E> static MSIM_HD u64 rpc_timeout(E& e) { return e.P->p[3] * NS_PER_MS; }
  template <class E> static MSIM_HD u64 hb(E& e) { return e.P->p[4] * NS_PER_MS; }

  // spawn a task that takes a clone of the node's Endpoint
  template <class E>
  static MSIM_HD int spawn_with_ep(E& e, u32 kind, int node, int sock) {
    int c = e.spawn(kind, node);
    if (c >= 0) e.sock_ref(c, sock);
    return c;
  }

  // ---------------- supervisor (node 0) ----------------
  template <class E>
  static MSIM_HD int poll_main(E& e, int t) {
    for (;;) switch (e.t_pc(t)) {
        case 0: {
          for (u32 i = 1; i <= n_nodes(e); i++) e.create_node(true, true);  // .ip(10.0.0.i).init(raft_node_main)
          if (fault_mask(e)) e.spawn(K_FAULT, 0);
          e.t_a(t) = 0;  // tick
          e.t_pc(t) = 1;
        }
        // fallthrough
        case 1:  // for tick in 1..=ticks { sleep(100ms)
          if ((u64)e.t_a(t) >= dur_ms(e) / 100) {
            outputs(e);  // computed by main's last poll, not when block_on returns
            return POLL_READY;
          }
          e.t_dl0(t) = e.sleep_deadline(e.now + 100 * NS_PER_MS);
          e.t_pc(t) = 2;
          // fallthrough
        case 2: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          u32 tick = ++e.t_a(t);
          bool any_leader = false;
          u32 n = n_nodes(e);
          for (u32 a = 1; a <= n; a++) {
            if (!e.wl_w(F_VALID, a) || e.wl_w(F_ROLE, a) != LEADER) continue;
            any_leader = true;
            for (u32 b2 = a + 1; b2 <= n; b2++)
              if (e.wl_w(F_VALID, b2) && e.wl_w(F_ROLE, b2) == LEADER && e.wl_w(F_TERM, b2) == e.wl_w(F_TERM, a))
                return e.panic(1);  // two leaders in one term
          }
          if (fault_mask(e) == 0 && tick == 20 && !any_leader) return e.panic(2);  // no leader by 2 s
          e.t_pc(t) = 1;
          continue;
        }
        default:
          return e.panic(0xbad);
      }
  }

  // ---------------- fault schedule (config 3) ----------------
  template <class E>
  static MSIM_HD int poll_fault(E& e, int t) {
    u32 n = n_nodes(e);
    for (;;) switch (e.t_pc(t)) {
        case 0:
          e.t_dl0(t) = e.sleep_deadline(e.now + e.gen_range_u64(500, 2000) * NS_PER_MS);
          e.t_pc(t) = 1;
          // fallthrough
        case 1: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          u32 kinds[3], nk = 0, mask = fault_mask(e);
          for (u32 k = 0; k < 3; k++)
            if (mask & (1u << k)) kinds[nk++] = k;
          u32 k = kinds[e.gen_range_u64(0, nk)];
          if (k == 0) {
            u32 i = (u32)e.gen_range_u64(1, n + 1);
            e.clog_node(i);
            e.t_a(t) = i;
            e.t_dl0(t) = e.sleep_deadline(e.now + e.gen_range_u64(200, 1000) * NS_PER_MS);
            e.t_pc(t) = 2;
            continue;
          } else if (k == 1) {
            u32 i = (u32)e.gen_range_u64(1, n + 1);
            u32 j = (u32)e.gen_range_u64(1, n);
            if (j >= i) j++;
            e.clog_link(i, j);
            e.clog_link(j, i);
            e.t_a(t) = i;
            e.t_b(t) = j;
            e.t_dl0(t) = e.sleep_deadline(e.now + e.gen_range_u64(200, 1000) * NS_PER_MS);
            e.t_pc(t) = 3;
            continue;
          } else {
            e.time_advance(e.gen_range_u64(0, 500) * NS_PER_MS);  // time::advance: fires due timers in this poll
            e.t_pc(t) = 0;
            continue;
          }
        }
        case 2:
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          e.unclog_node(e.t_a(t));
          e.t_pc(t) = 0;
          continue;
        case 3:
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          e.unclog_link(e.t_a(t), e.t_b(t));
          e.unclog_link(e.t_b(t), e.t_a(t));
          e.t_pc(t) = 0;
          continue;
        default:
          return e.panic(0xbad);
      }
  }

  // ---------------- node init task: bind, serve, election ticker ----------------
  template <class E>
  static MSIM_HD int poll_node(E& e, int t) {
    using L = typename E::L;
    u32 me = e.t_node(t);
    for (;;) switch (e.t_pc(t)) {
        case 0:
          e.t_dl0(t) = e.rand_delay_deadline();  // Endpoint::bind
          e.t_pc(t) = 1;
          // fallthrough
        case 1: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int k = e.sock_bind(me, raft_addr(me));
          if (k < 0) return e.panic(30);
          e.sock_adopt(t, k);
          e.wl_w(F_TERM, me) = 0;
          e.wl_w(F_VOTED, me) = 0;
          e.wl_w(F_ROLE, me) = FOLLOWER;
          e.wl_w(F_VOTES, me) = 0;
          e.wl_w(F_LEADER, me) = 0;
          e.wl_q(FQ_LAST_HB, me) = e.now;
          e.wl_w(F_VALID, me) = 1;
          spawn_with_ep(e, K_LOOP_VOTE, me, k);    // ep.add_rpc_handler::<RequestVote>
          spawn_with_ep(e, K_LOOP_APPEND, me, k);  // ep.add_rpc_handler::<AppendEntries>
          e.t_pc(t) = 2;
        }
        // fallthrough
        case 2: {  // loop { timeout = thread_rng().gen_range(lo..hi) ms; sleep(timeout)
          u64 to_ms = e.gen_range_u64(e.P->p[5], e.P->p[6]);
          e.t_a(t) = (u32)to_ms;
          e.t_dl0(t) = e.sleep_deadline(e.now + to_ms * NS_PER_MS);
          e.t_pc(t) = 3;
        }
        // fallthrough
        case 3: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          u64 to = (u64)e.t_a(t) * NS_PER_MS;
          if (e.wl_w(F_ROLE, me) != LEADER && e.now - e.wl_q(FQ_LAST_HB, me) >= to) {
            // start an election
            u32 term0 = ++e.wl_w(F_TERM, me);
            e.wl_w(F_ROLE, me) = CANDIDATE;
            e.wl_w(F_VOTED, me) = me;
            e.wl_w(F_VOTES, me) = 1;
            e.wl_q(FQ_LAST_HB, me) = e.now;
            e.wl_s(0)++;  // elections
            int k = e.t_sock(t);
            for (u32 peer = 1; peer <= n_nodes(e); peer++)
              if (peer != me) {
                int c = spawn_with_ep(e, K_VOTE, me, k);
                if (c < 0) return POLL_PENDING;
                e.t_a(c) = peer;
                e.t_b(c) = term0;
              }
          }
          e.t_pc(t) = 2;
          continue;
        }
        default:
          return e.panic(0xbad);
      }
    (void)sizeof(L);
  }

  // ---------------- add_rpc_handler: the receive loop (rpc.rs:143-165) ----------------
  template <class E>
  static MSIM_HD int poll_rpc_loop(E& e, int t, u64 req_tag, u32 req_kind) {
    using L = typename E::L;
    for (;;) switch (e.t_pc(t)) {
        case 0:  // recv_from_raw(req_tag): mailbox.recv
          e.mbox_recv(t, e.t_sock(t), req_tag);
          e.t_pc(t) = 1;
          // fallthrough
        case 1:
          if (e.t_rx(t) == NONE8) return POLL_PENDING;
          e.t_dl0(t) = e.rand_delay_deadline();
          e.t_pc(t) = 2;
          // fallthrough
        case 2: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int m = e.t_rx(t);
          u64 rsp_tag = e.q(L::Q_M_W, m);
          u32 v0 = e.w(L::W_M_V0, m), v1 = e.w(L::W_M_V1, m), from = e.w(L::W_M_FROM, m);
          e.msg_free(m);
          e.t_rx(t) = NONE8;
          // let rsp_future = f(req); spawn(async move { rsp = rsp_future.await; send_to_raw(from, rsp_tag, rsp) })
          int c = spawn_with_ep(e, req_kind, e.t_node(t), e.t_sock(t));
          if (c < 0) return POLL_PENDING;
          e.t_a(c) = v0;
          e.t_b(c) = v1;
          e.t_c(c) = from;
          e.t_dl1(c) = rsp_tag;
          e.t_pc(t) = 0;
          continue;
        }
        default:
          return e.panic(0xbad);
      }
  }

  // ---------------- one task per request: handler body + reply ----------------
  template <class E>
  static MSIM_HD int poll_rpc_req(E& e, int t, bool is_vote) {
    using L = typename E::L;
    u32 me = e.t_node(t);
    switch (e.t_pc(t)) {
      case 0: {  // rsp_future.await — the handler body runs in THIS task's first poll
        u32 term = e.t_a(t), who = e.t_b(t);
        u32 ok;
        if (is_vote) {  // raft_on_request_vote
          if (term > e.wl_w(F_TERM, me)) {
            e.wl_w(F_TERM, me) = term;
            e.wl_w(F_ROLE, me) = FOLLOWER;
            e.wl_w(F_VOTED, me) = 0;
          }
          u32 voted = e.wl_w(F_VOTED, me);
          ok = term == e.wl_w(F_TERM, me) && (bug(e) == 1 || voted == 0 || voted == who);
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
        e.t_a(t) = e.wl_w(F_TERM, me);
        e.t_b(t) = ok;
        e.t_dl0(t) = e.rand_delay_deadline();  // send_to_raw(from, rsp_tag, rsp): rand_delay
        e.t_pc(t) = 1;
      }
      // fallthrough
      case 1: {
        if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
        int m = e.msg_alloc();
        if (m < 0) return POLL_PENDING;
        e.q(L::Q_M_TAG, m) = e.t_dl1(t);
        e.q(L::Q_M_W, m) = 0;
        e.w(L::W_M_V0, m) = e.t_a(t);
        e.w(L::W_M_V1, m) = e.t_b(t);
        int k = e.t_sock(t);
        e.net_send(e.b(L::B_SK_NODE, k), e.sk_port(k), e.t_c(t), m);
        return POLL_READY;
      }
    }
    return e.panic(0xbad);
  }

  // ---------------- Endpoint::call_timeout as a nested state machine ----------------
  enum : int { CALL_PENDING = 0, CALL_OK = 1, CALL_TIMEOUT = 2, CALL_PANIC = 3 };
  /// Polls `timeout(T, ep.call(dst, req))`.  The Sleep of the timeout lives in t_dl1 (created by
  /// the caller when call_timeout is first polled); the call's own progress is t_sub.
  template <class E>
  static MSIM_HD int call_timeout_poll(E& e, int t, u32 dst, u64 req_tag, u32 req_v0, u32 req_v1, u32& rsp_v0,
                                       u32& rsp_v1) {
    using L = typename E::L;
    int k = e.t_sock(t);
    // ---- poll the call future first (select_biased) ----
    switch (e.t_sub(t)) {
      case 0: {  // rsp_tag = random::<u64>(); send_to_raw: rand_delay
        u64 rsp_tag = e.next_u64();
        int m = e.msg_alloc();
        if (m < 0) return CALL_PENDING;
        e.q(L::Q_M_TAG, m) = req_tag;
        e.q(L::Q_M_W, m) = rsp_tag;
        e.w(L::W_M_V0, m) = req_v0;
        e.w(L::W_M_V1, m) = req_v1;
        e.t_msg(t) = (u8)m;  // the request waits inside the future during rand_delay
        e.t_dl0(t) = e.rand_delay_deadline();
        e.t_sub(t) = 1;
      }
      // fallthrough
      case 1: {
        if (!e.sleep_poll(t, e.t_dl0(t))) break;
        int m = e.t_msg(t);
        e.t_msg(t) = NONE8;
        u64 rsp_tag = e.q(L::Q_M_W, m);
        e.net_send(e.b(L::B_SK_NODE, k), e.sk_port(k), dst, m);
        e.mbox_recv(t, k, rsp_tag);  // recv_from_raw(rsp_tag)
        e.t_sub(t) = 2;
      }
      // fallthrough
      case 2:
        if (e.t_rx(t) == NONE8) break;
        e.t_dl0(t) = e.rand_delay_deadline();
        e.t_sub(t) = 3;
        // fallthrough
      case 3: {
        if (!e.sleep_poll(t, e.t_dl0(t))) break;
        int m = e.t_rx(t);
        u32 from = e.w(L::W_M_FROM, m);
        rsp_v0 = e.w(L::W_M_V0, m);
        rsp_v1 = e.w(L::W_M_V1, m);
        e.msg_free(m);
        e.t_rx(t) = NONE8;
        e.t_sub(t) = 4;
        if (from != dst) return CALL_PANIC;  // assert_eq!(from, dst)
        return CALL_OK;
      }
    }
    if (!e.running()) return CALL_PENDING;
    // ---- then the timeout's Sleep ----
    if (e.now >= e.t_dl1(t)) {
      // Elapsed: the call future is dropped where it stands
      if (e.t_sub(t) == 2) e.mbox_unregister(t);
      if (e.t_rx(t) != NONE8) {
        e.msg_free(e.t_rx(t));
        e.t_rx(t) = NONE8;
      }
      if (e.t_msg(t) != NONE8) {
        e.msg_free(e.t_msg(t));
        e.t_msg(t) = NONE8;
      }
      e.t_sub(t) = 4;
      return CALL_TIMEOUT;
    }
    e.timer_push(e.t_dl1(t), timer_payload(TK_WAKE, t, e.t_gen(t)));
    return CALL_PENDING;
  }

  template <class E>
  static MSIM_HD void step_down(E& e, u32 me, u32 term) {
    e.wl_w(F_TERM, me) = term;
    e.wl_w(F_ROLE, me) = FOLLOWER;
    e.wl_w(F_VOTED, me) = 0;
  }

  // ---------------- raft_vote_task / raft_append_task ----------------
  template <class E>
  static MSIM_HD int poll_call_task(E& e, int t, bool is_vote) {
    u32 me = e.t_node(t);
    switch (e.t_pc(t)) {
      case 0:  // call_timeout(..): the timeout Sleep is created when it is first polled
        e.t_dl1(t) = e.sleep_deadline(e.now + rpc_timeout(e));
        e.t_sub(t) = 0;
        e.t_pc(t) = 1;
        // fallthrough
      case 1: {
        u32 peer = e.t_a(t), term0 = e.t_b(t), v0 = 0, v1 = 0;
        int r = call_timeout_poll(e, t, raft_addr(peer), is_vote ? TAG_VOTE : TAG_APPEND, term0, me, v0, v1);
        if (r == CALL_PENDING) return POLL_PENDING;
        if (r == CALL_PANIC) return e.panic(31);
        if (r == CALL_OK) {
          if (is_vote) {
            if (v0 > e.wl_w(F_TERM, me)) {
              step_down(e, me, v0);
            } else if (e.wl_w(F_ROLE, me) == CANDIDATE && e.wl_w(F_TERM, me) == term0 && v1 != 0) {
              u32 votes = ++e.wl_w(F_VOTES, me);
              if (votes == n_nodes(e) / 2 + 1) {
                e.wl_w(F_ROLE, me) = LEADER;
                e.wl_w(F_LEADER, me) = me;
                int c = spawn_with_ep(e, K_LEADER, me, e.t_sock(t));
                if (c < 0) return POLL_PENDING;
                e.t_a(c) = term0;
              }
            }
          } else {
            if (v0 > e.wl_w(F_TERM, me)) step_down(e, me, v0);
          }
        }
        return POLL_READY;
      }
    }
    return e.panic(0xbad);
  }

  // ---------------- raft_leader_loop ----------------
  template <class E>
  static MSIM_HD int poll_leader(E& e, int t) {
    u32 me = e.t_node(t);
    for (;;) switch (e.t_pc(t)) {
        case 0: {
          u32 term0 = e.t_a(t);
          if (!(e.wl_w(F_ROLE, me) == LEADER && e.wl_w(F_TERM, me) == term0)) return POLL_READY;
          int k = e.t_sock(t);
          for (u32 peer = 1; peer <= n_nodes(e); peer++)
            if (peer != me) {
              int c = spawn_with_ep(e, K_APPEND, me, k);
              if (c < 0) return POLL_PENDING;
              e.t_a(c) = peer;
              e.t_b(c) = term0;
            }
          e.t_dl0(t) = e.sleep_deadline(e.now + hb(e));
          e.t_pc(t) = 1;
        }
        // fallthrough
        case 1:
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          e.t_pc(t) = 0;
          continue;
        default:
          return e.panic(0xbad);
      }
  }

  template <class E>
  static MSIM_HD int poll(E& e, int t) {
    switch (e.t_kind(t)) {
      case K_MAIN: return poll_main(e, t);
      case K_FAULT: return poll_fault(e, t);
      case K_NODE: return poll_node(e, t);
      case K_LOOP_VOTE: return poll_rpc_loop(e, t, TAG_VOTE, K_REQ_VOTE);
      case K_LOOP_APPEND: return poll_rpc_loop(e, t, TAG_APPEND, K_REQ_APPEND);
      case K_REQ_VOTE: return poll_rpc_req(e, t, true);
      case K_REQ_APPEND: return poll_rpc_req(e, t, false);
      case K_VOTE: return poll_call_task(e, t, true);
      case K_APPEND: return poll_call_task(e, t, false);
      default: return poll_leader(e, t);
    }
  }
  template <class E>
  static MSIM_HD void on_drop(E& e, int t) {
    u32 kind = e.t_kind(t);
    bool registered = ((kind == K_LOOP_VOTE || kind == K_LOOP_APPEND) && e.t_pc(t) == 1) ||
                      ((kind == K_VOTE || kind == K_APPEND) && e.t_sub(t) == 2);
    e.drop_task_resources(t, registered);
  }
  /// NodeBuilder::init(|| raft_node_main(c, i)) — also what Handle::restart re-runs
  template <class E>
  static MSIM_HD void node_init(E& e, int n) {
    e.spawn(K_NODE, n, TF_INIT);
  }
  template <class E>
  static MSIM_HD void finish(E&) {}
  /// outputs of wl_raft_main
  template <class E>
  static MSIM_HD void outputs(E& e) {
    using L = typename E::L;
    u32 n = n_nodes(e), term = 0, leader = 0, best = 0;
    for (u32 a = 1; a <= n; a++) {
      if (!e.wl_w(F_VALID, a)) continue;
      if (e.wl_w(F_TERM, a) > term) term = e.wl_w(F_TERM, a);
    }
    for (u32 a = 1; a <= n; a++)
      if (e.wl_w(F_VALID, a) && e.wl_w(F_ROLE, a) == LEADER && e.wl_w(F_TERM, a) >= best) {
        best = e.wl_w(F_TERM, a);
        leader = a;
      }
    e.w(L::W_USER0) = leader;
    e.w(L::W_USER1) = term;
    e.w(L::W_USER2) = e.wl_s(0);
    e.w(L::W_USER3) = e.w(L::W_MSG_COUNT);
  }
};

}  // namespace msim
