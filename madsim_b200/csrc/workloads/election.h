// MSIM_WL_ELECTION — SURVEY §8(f) row 3: the election API of madsim-etcd-client and the watcher behind it, as the
// reference's own `election()` test drives it (madsim-etcd-client/tests/test.rs:129-232).
//
//   ElectionClient::{campaign, proclaim, leader, observe, resign}   src/election.rs:27-100
//   SimServer::serve — Campaign / Observe arms                      src/server.rs:71-95  (select_biased!{ tx.closed(), .. })
//   EtcdService::{campaign, ..}, timeout()                           src/service.rs:94-176
//   ServiceInner::{campaign, proclaim, leader, observe, resign}      src/service.rs:494-585
//   EventBus: subscribe / publish over mpsc::channel(4)              src/service.rs:196-245
//   lease_grant with a server-chosen id, lease_revoke, tick          src/service.rs:398-492
//
// Oracle twin: oracle/workloads_election.hpp (the same program as coroutines; its header explains the wire encoding:
// names {0 "leader", 1 "invalid_lease"}, values 10 / 11 / 12 / 13 / 20, a lease travels as the low 32 bits of its id).
// The store is a fixed table in the seed's state block: two leases, three keys, four watcher subscriptions.  The
// receivers of watcher events only count arrivals (every event makes them re-read the leader), so a subscription is
// a counter, two liveness bits and the parked receiver's waker.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_Election {
  static constexpr int NLEASE = 2, NKV = 3, NSUB = 4;
  enum : int {  // wl_s scalars
    S_REV = 0, S_HANDLED, S_CNT, S_SEEN, S_SUM, S_NLIST,
    S_JH,                        // [3] the clients' JoinHandles (main)
    S_LEASE = S_JH + 3,          // [NLEASE*4] id (low 32 bits), ttl, keys (mask of kv slots), used
    S_KV = S_LEASE + 4 * NLEASE, // [NKV*6] used, name, lease, value, create_revision, modify_revision
    S_SUB = S_KV + 6 * NKV,      // [NSUB*2] used | rx_alive<<1 | tx_alive<<2 | name<<8 | queued<<16 ; parked receiver
    S_LIST = S_SUB + 2 * NSUB,   // [NSUB] EventBus::list: subscription slots in subscribe order
    S_END = S_LIST + NSUB
  };
  enum : int { F_STEP = 0, F_LEASE, F_RSP0, F_RSP1, F_NFIELD };  // wl_w per client node
  struct Caps {
    static constexpr int NT = 16, TQ = 48, NMSG = 2, NSOCK = 5, NREG = 2, NMBOX = 2, NNODE = 5, NPARK = 1;
    static constexpr int WQ = 0, WW = F_NFIELD, WS = S_END;
    static constexpr bool USES_JOIN = true;
    static constexpr bool USES_KILL = false;
    static constexpr bool USES_CTRLC = false;
    static constexpr bool USES_CLOSED = true;   // the Campaign / Observe arms race tx.closed()
    static constexpr bool USES_SPLIT = false;
    static constexpr bool USES_TIMEOUT = false;
    static constexpr bool USES_CONN2 = true;    // the observer keeps its ObserveStream while it makes a kv call
    static constexpr bool USES_PMATCH = false;   // does any node use restart_on_panic_matching?
    static constexpr unsigned PMATCH_BASE = 0;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    static constexpr bool BIND_INLINE = true;  // every call binds an ephemeral endpoint: Network::bind inline, +3 % (profiles/r2w_bind_modes.txt)
    static constexpr bool COLD_SPAWN = MSIM_COLD_DEFAULT & 1, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;  // once-per-request runtime pieces out of line
    static constexpr int NCONN = 6, NCQ = 4, NACC = 4;
  };
  enum : u32 { K_MAIN = 0, K_SERVE, K_TICK, K_REQ, K_C1, K_C2, K_C3 };
  enum : u32 { OP_GRANT = 1, OP_REVOKE = 2, OP_CAMPAIGN = 3, OP_PROCLAIM = 4, OP_LEADER = 5, OP_OBSERVE = 6, OP_RESIGN = 7, OP_GET_PREFIX = 8 };
  enum : u32 { ST_OK = 0, ST_TIMED_OUT = 1, ST_LEASE_NOT_FOUND = 2, ST_SESSION_EXPIRED = 3 };
  enum : u32 { SUB_USED = 1, SUB_RX = 2, SUB_TX = 4 };
  static constexpr int PCS = 32;
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 2379u;  // 10.0.0.1:2379
  static MSIM_HD int init_kind(int) { return -1; }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }
  struct Rsp {
    u32 status = ST_OK, flag = 0, revision = 0, data = 0;
    int panic_code = 0;
    MSIM_HD u32 v0() const { return status | flag << 4 | revision << 8; }
  };

  // ---------------- EventBus (service.rs:196-245) ----------------
  template <class E>
  static MSIM_HD u32& sub_flags(E& e, int s) { return e.wl_s(S_SUB + 2 * s); }
  template <class E>
  static MSIM_HD u32& sub_waiter(E& e, int s) { return e.wl_s(S_SUB + 2 * s + 1); }
  /// watcher.subscribe(Prefix(name), tx): → the subscription slot (the Receiver's handle) or -1
  template <class E>
  static MSIM_HD int subscribe(E& e, u32 name) {
    int s = 0;
    while (s < NSUB && (sub_flags(e, s) & SUB_USED)) s++;
    u32 n = e.wl_s(S_NLIST);
    if (s >= NSUB || n >= (u32)NSUB) {
      e.fail_cap(MSIM_CAP_REGISTERED);
      return -1;
    }
    sub_flags(e, s) = SUB_USED | SUB_RX | SUB_TX | name << 8;
    sub_waiter(e, s) = NONE32;
    e.wl_s(S_LIST + n) = (u32)s;
    e.wl_s(S_NLIST) = n + 1;
    return s;
  }
  /// watcher.publish(event) for an event on `name`: try_send to every matching subscriber in list order; a
  /// subscriber whose channel is full or whose Receiver is gone leaves the list (its Sender drops)
  template <class E>
  static MSIM_HD void publish(E& e, u32 name) {
    u32 n = e.wl_s(S_NLIST), keep = 0;
MSIM_NOUNROLL
    for (u32 i = 0; i < n; i++) {
      int s = (int)e.wl_s(S_LIST + i);
      u32 f = sub_flags(e, s);
      bool stay = true;
      if (((f >> 8) & 0xff) == name) {
        u32 queued = f >> 16;
        if ((f & SUB_RX) && queued < 4) {  // try_send ok: the event is queued, the parked receiver wakes
          sub_flags(e, s) = f + (1u << 16);
          e.wake_waiter(sub_waiter(e, s));
        } else {
          stay = false;
          if (f & SUB_RX) {  // full: the Sender drops, the receiver will see None after draining
            sub_flags(e, s) = f & ~(u32)SUB_TX;
            e.wake_waiter(sub_waiter(e, s));
          } else {
            sub_flags(e, s) = 0;  // both ends gone
          }
        }
      }
      if (stay) e.wl_s(S_LIST + keep++) = (u32)s;
    }
    e.wl_s(S_NLIST) = keep;
  }
  /// the Receiver of subscription `s` drops
  template <class E>
  static MSIM_HD void sub_drop_rx(E& e, int s) {
    u32 f = sub_flags(e, s);
    sub_waiter(e, s) = NONE32;
    sub_flags(e, s) = (f & SUB_TX) ? (f & ~(u32)SUB_RX) : 0u;
  }
  /// rx.recv(): 1 an event, 0 None (the bus dropped the Sender and nothing is queued), -1 pending (parked)
  template <class E>
  static MSIM_HD int sub_recv(E& e, int s, int t) {
    u32 f = sub_flags(e, s);
    if (f >> 16) {
      sub_flags(e, s) = f - (1u << 16);
      return 1;
    }
    if (!(f & SUB_TX)) return 0;
    sub_waiter(e, s) = e.waker_of(t);
    return -1;
  }

  // ---------------- ServiceInner ----------------
  template <class E>
  static MSIM_HD int find_lease(E& e, u32 lo32) {
    for (int i = 0; i < NLEASE; i++)
      if (e.wl_s(S_LEASE + 4 * i + 3) && e.wl_s(S_LEASE + 4 * i) == lo32) return i;
    return -1;
  }
  template <class E>
  static MSIM_HD int find_kv(E& e, u32 name, u32 lease) {
    for (int i = 0; i < NKV; i++)
      if (e.wl_s(S_KV + 6 * i) && e.wl_s(S_KV + 6 * i + 1) == name && e.wl_s(S_KV + 6 * i + 2) == lease) return i;
    return -1;
  }
  /// leader(name): the key of the prefix with the smallest create_revision, or -1
  template <class E>
  static MSIM_HD int leader(E& e, u32 name) {
    int best = -1;
    for (int i = 0; i < NKV; i++)
      if (e.wl_s(S_KV + 6 * i) && e.wl_s(S_KV + 6 * i + 1) == name &&
          (best < 0 || e.wl_s(S_KV + 6 * i + 4) < e.wl_s(S_KV + 6 * best + 4)))
        best = i;
    return best;
  }
  /// what `new_leader.kv == leader.kv` compares, folded into one word: a key version is identified by its
  /// modify_revision (every change of a key takes a fresh revision)
  template <class E>
  static MSIM_HD u32 leader_snapshot(E& e, u32 name) {
    int k = leader(e, name);
    return k < 0 ? 0u : (0x80000000u | e.wl_s(S_KV + 6 * k + 5));
  }
  template <class E>
  static MSIM_HD Rsp leader_rsp(E& e, u32 name) {
    Rsp r;
    r.revision = e.wl_s(S_REV);
    int k = leader(e, name);
    if (k >= 0) r.flag = 1, r.data = e.wl_s(S_KV + 6 * k + 3) | e.wl_s(S_KV + 6 * k + 5) << 16;
    return r;
  }
  /// the keys of a lease go (revoke / expiry): kv.remove(key).expect("no key"); publish(Delete)
  template <class E>
  static MSIM_HD int drop_lease_keys(E& e, u32 keys) {
    for (int k = 0; k < NKV; k++) {
      if (!(keys >> k & 1)) continue;
      if (!e.wl_s(S_KV + 6 * k)) return 63;
      u32 name = e.wl_s(S_KV + 6 * k + 1);
      e.wl_s(S_KV + 6 * k) = 0;
      publish(e, name);
    }
    return 0;
  }
  template <class E>
  static MSIM_HD Rsp sv_lease_grant(E& e, u32 ttl) {
    Rsp r;
    u64 id = 0;
    while (id == 0 || find_lease(e, (u32)id) >= 0) {  // random::<i64>().abs()
      u64 x = e.random_u64();
      id = (x >> 63) ? (0 - x) : x;
    }
    int i = 0;
    while (i < NLEASE && e.wl_s(S_LEASE + 4 * i + 3)) i++;
    if (i >= NLEASE) {
      e.fail_cap(MSIM_CAP_SOCKETS);
      return r;
    }
    e.wl_s(S_LEASE + 4 * i) = (u32)id, e.wl_s(S_LEASE + 4 * i + 1) = ttl, e.wl_s(S_LEASE + 4 * i + 2) = 0, e.wl_s(S_LEASE + 4 * i + 3) = 1;
    r.revision = ++e.wl_s(S_REV);
    r.data = (u32)id;
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_lease_revoke(E& e, u32 lo32) {
    Rsp r;
    int i = find_lease(e, lo32);
    if (i < 0) {
      r.status = ST_LEASE_NOT_FOUND;
      return r;
    }
    u32 keys = e.wl_s(S_LEASE + 4 * i + 2);
    e.wl_s(S_LEASE + 4 * i + 3) = 0;
    r.panic_code = drop_lease_keys(e, keys);
    r.revision = ++e.wl_s(S_REV);
    return r;
  }
  template <class E>
  static MSIM_HD int sv_tick(E& e) {
    bool any = false;
    for (int i = 0; i < NLEASE; i++) {
      if (!e.wl_s(S_LEASE + 4 * i + 3)) continue;
      int ttl = (int)e.wl_s(S_LEASE + 4 * i + 1) - 1;
      e.wl_s(S_LEASE + 4 * i + 1) = (u32)ttl;
      if (ttl > 0) continue;
      e.wl_s(S_LEASE + 4 * i + 3) = 0;
      any = true;
      int bad = drop_lease_keys(e, e.wl_s(S_LEASE + 4 * i + 2));
      if (bad) return bad;
    }
    if (any) e.wl_s(S_REV)++;
    return 0;
  }
  /// ServiceInner::campaign (service.rs:494-538).  *sub = -2: the caller leads (r is the response); >= 0: subscribed
  template <class E>
  static MSIM_HD Rsp sv_campaign(E& e, u32 name, u32 value, u32 lease, int* sub) {
    Rsp r;
    *sub = -2;
    int k = find_kv(e, name, lease);
    if (k < 0 || e.wl_s(S_KV + 6 * k + 3) != value) {
      u32 rev = ++e.wl_s(S_REV);  // before the lease lookup: an unknown lease still moves the revision
      int l = find_lease(e, lease);
      if (l < 0) {
        r.status = ST_LEASE_NOT_FOUND;
        return r;
      }
      if (k < 0) {
        k = 0;
        while (k < NKV && e.wl_s(S_KV + 6 * k)) k++;
        if (k >= NKV) {
          e.fail_cap(MSIM_CAP_MSGS);
          return r;
        }
      }
      e.wl_s(S_LEASE + 4 * l + 2) |= 1u << k;
      e.wl_s(S_KV + 6 * k) = 1, e.wl_s(S_KV + 6 * k + 1) = name, e.wl_s(S_KV + 6 * k + 2) = lease;
      e.wl_s(S_KV + 6 * k + 3) = value, e.wl_s(S_KV + 6 * k + 4) = rev, e.wl_s(S_KV + 6 * k + 5) = rev;
      publish(e, name);
    }
    int ld = leader(e, name);
    if (e.wl_s(S_KV + 6 * ld + 2) == lease) {
      r.flag = 1, r.revision = e.wl_s(S_REV), r.data = e.wl_s(S_REV);
    } else {
      *sub = subscribe(e, name);
    }
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_proclaim(E& e, u32 name, u32 lease, u32 value) {
    Rsp r;
    int k = find_kv(e, name, lease);
    if (k < 0) {
      r.status = ST_SESSION_EXPIRED;
      return r;
    }
    u32 rev = ++e.wl_s(S_REV);
    e.wl_s(S_KV + 6 * k + 3) = value, e.wl_s(S_KV + 6 * k + 5) = rev;
    publish(e, name);
    r.revision = rev;
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_resign(E& e, u32 name, u32 lease) {
    Rsp r;
    int k = find_kv(e, name, lease);
    if (k < 0) {
      r.status = ST_SESSION_EXPIRED;
      return r;
    }
    int l = find_lease(e, lease);
    if (l < 0) {
      r.panic_code = 68;  // .expect("no lease")
      return r;
    }
    e.wl_s(S_LEASE + 4 * l + 2) &= ~(1u << k);
    e.wl_s(S_KV + 6 * k) = 0;
    publish(e, name);
    r.revision = ++e.wl_s(S_REV);
    return r;
  }
  /// EtcdService::timeout (service.rs:165-176): gen_bool(timeout_rate) — p == 1.0 draws nothing
  template <class E>
  static MSIM_HD bool timed_out(E& e) {
    return e.P->p[5] ? true : e.gen_bool_int(e.P->p[6]);
  }

  // ---------------- the clients' scripts (tests/test.rs:154-232) ----------------
  /// what the client does at step F_STEP: 0 a call was prepared (t_a, t_b), 1 a request to return (sleep), 2 done,
  /// 3 (observer) the next stream message, 4 (observer) open the stream, < 0: -(panic code)
  template <class E>
  static MSIM_HD int script(E& e, int t, u32 kind) {
    const u32 node = e.t_node(t), step = e.wl_w(F_STEP, node), lease = e.wl_w(F_LEASE, node);
    auto call = [&](u32 op, u32 name, u32 value, u32 ls) {
      e.t_a(t) = op | name << 4 | value << 8;
      e.t_b(t) = ls;
      return 0;
    };
    auto nap = [&](u64 secs) {
      e.wl_w(F_STEP, node) = step + 1;
      e.xa = secs * NS_PER_SEC;
      return 1;
    };
    if (kind == K_C1) switch (step) {
        case 0: return nap(5);
        case 1: return call(OP_GRANT, 0, 60, 0);
        case 2: case 4: return call(OP_CAMPAIGN, 0, 10, lease);
        case 3: case 7: return call(OP_LEADER, 0, 0, 0);
        case 5: return call(OP_CAMPAIGN, 0, 11, lease);
        case 6: return call(OP_PROCLAIM, 0, 12, lease);
        case 8: return nap(e.P->p[0]);
        case 9: return call(OP_REVOKE, 0, 0, lease);
        case 10: return call(OP_PROCLAIM, 0, 13, lease);
        case 11: return call(OP_CAMPAIGN, 1, 10, 1);
        default: return 2;
      }
    if (kind == K_C2) switch (step) {
        case 0: return nap(10);
        case 1: return call(OP_GRANT, 0, 60, 0);
        case 2: return call(OP_CAMPAIGN, 0, 20, lease);
        case 3: return call(OP_RESIGN, 0, 0, lease);
        default: return 2;
      }
    switch (step) {  // K_C3, the observer
      case 0: return 4;
      case 1: case 2: case 3: case 6: return 3;
      case 4: return nap(15);
      case 5: return call(OP_GET_PREFIX, 0, 0, 0);
      default: return 2;
    }
  }
  /// the assertions on a call's response; 0 or the panic code
  template <class E>
  static MSIM_HD int check(E& e, u32 kind, u32 node, u32 v0, u32 v1) {
    const u32 step = e.wl_w(F_STEP, node), status = v0 & 15, flag = (v0 >> 4) & 1;
    const bool ok = status == ST_OK;
    if (kind == K_C1) switch (step) {
        case 1:
          if (!ok) return 82;
          e.wl_w(F_LEASE, node) = v1;
          return 0;
        case 2: return (!ok || !flag) ? 83 : 0;
        case 3: return (!ok || !flag || (v1 & 0xffff) != 10) ? 84 : 0;
        case 4: case 5: return ok ? 0 : 85;
        case 6: return ok ? 0 : 86;
        case 7: return (!ok || !flag || (v1 & 0xffff) != 12) ? 84 : 0;
        case 9: return ok ? 0 : 87;
        default: return ok ? 88 : 0;  // proclaim after the revoke, campaign under an unknown lease: unwrap_err
      }
    if (kind == K_C2) switch (step) {
        case 1:
          if (!ok) return 82;
          e.wl_w(F_LEASE, node) = v1;
          return 0;
        case 2: return (!ok || !flag) ? 83 : 0;
        default: return ok ? 0 : 89;
      }
    return (!ok || v1 != 2) ? 92 : 0;  // the observer's get("leader", prefix): two keys
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const Params& P = *e.P;
    for (;;) switch (e.t_kind(t) * PCS + e.t_pc(t)) {
        // ---- wl_election_main ----
        case K_MAIN * PCS + 0:  // the server node and the three client nodes
          if (e.wl_s(S_CNT) < 4) {
            e.wl_s(S_CNT)++;
            e.ya = NODE_IP;
            return RQ_CREATE_NODE;
          }
          e.t_pc(t) = 1;
          e.ya = K_SERVE, e.yb = 1, e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_MAIN * PCS + 1:
          e.t_pc(t) = 2;
          e.wl_s(S_CNT) = 0;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        case K_MAIN * PCS + 3:
          e.wl_s(S_JH + e.wl_s(S_CNT)) = (u32)e.ret;
          e.wl_s(S_CNT)++;
          e.t_pc(t) = 2;
          // fallthrough
        case K_MAIN * PCS + 2:  // client1.spawn(first leader), client2.spawn(second), client3.spawn(observer)
          if (e.wl_s(S_CNT) < 3) {
            e.t_pc(t) = 3;
            e.ya = (K_C1 + e.wl_s(S_CNT)) | SPAWN_JOINABLE, e.yb = 2 + e.wl_s(S_CNT), e.yc = e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          e.wl_s(S_CNT) = 0;
          e.t_pc(t) = 4;
          // fallthrough
        case K_MAIN * PCS + 4:  // task1.await.unwrap(); task2..; task3..
          if (e.wl_s(S_CNT) < 3) {
            e.t_pc(t) = 5;
            e.ya = e.wl_s(S_JH + e.wl_s(S_CNT));
            return RQ_JOIN;
          }
          e.w(L::W_USER0) = e.wl_s(S_SEEN);
          e.w(L::W_USER1) = e.wl_s(S_REV);
          e.w(L::W_USER2) = e.wl_s(S_SUM);
          e.w(L::W_USER3) = e.wl_s(S_HANDLED);
          return RQ_READY;
        case K_MAIN * PCS + 5:
          if (e.ret == 2) return panic(e, 0);
          e.wl_s(S_CNT)++;
          e.t_pc(t) = 4;
          continue;

        // ---- SimServer::serve ----
        case K_SERVE * PCS + 0:
          e.t_pc(t) = 1;
          e.ya = SERVER_ADDR;
          return RQ_BIND;
        case K_SERVE * PCS + 1:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          e.ya = K_TICK, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_SERVE * PCS + 2:
          e.t_pc(t) = 3;
          return RQ_ACCEPT1;
        case K_SERVE * PCS + 3:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          e.ya = K_REQ | SPAWN_CONN, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_TICK * PCS + 0: {
          int bad = sv_tick(e);
          if (bad) return panic(e, (u32)bad);
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        }

        // ---- the task per accepted connection (server.rs:46-100) ----
        case K_REQ * PCS + 0:
          e.t_pc(t) = 1;
          return RQ_CH_RECV;
        case K_REQ * PCS + 1: {
          if (e.ret < 0) return RQ_READY;
          e.t_a(t) = e.rv0, e.t_b(t) = e.rv1;
          e.wl_s(S_HANDLED)++;
          u32 op = e.rv0 & 15;
          if (op == OP_CAMPAIGN) {
            e.t_pc(t) = 10;
            continue;
          }
          if (op == OP_OBSERVE) {
            e.t_pc(t) = 20;
            continue;
          }
          e.t_pc(t) = 2;
        }
          // fallthrough
        case K_REQ * PCS + 2: {  // service.<op>(..).await: timeout().await?, then the store
          if (timed_out(e)) {
            e.t_pc(t) = 3;
            e.xa = e.sample_inclusive_u64(5 * NS_PER_SEC, 15 * NS_PER_SEC - 1);
            return RQ_SLEEP_FOR;
          }
          u32 r0 = e.t_a(t), op = r0 & 15, name = (r0 >> 4) & 15, value = r0 >> 8, lease = e.t_b(t);
          Rsp r;
          switch (op) {
            case OP_GRANT: r = sv_lease_grant(e, value); break;
            case OP_REVOKE: r = sv_lease_revoke(e, lease); break;
            case OP_PROCLAIM: r = sv_proclaim(e, name, lease, value); break;
            case OP_LEADER: r = leader_rsp(e, name); break;
            case OP_RESIGN: r = sv_resign(e, name, lease); break;
            case OP_GET_PREFIX: {
              r.revision = e.wl_s(S_REV);
              for (int k = 0; k < NKV; k++) r.data += e.wl_s(S_KV + 6 * k) && e.wl_s(S_KV + 6 * k + 1) == name;
              r.flag = r.data != 0;
              break;
            }
            default: r.panic_code = 64;
          }
          if (r.panic_code) return panic(e, (u32)r.panic_code);
          e.t_pc(t) = 9;
          e.yb = r.v0(), e.yc = r.data;
          return RQ_CH_SEND;
        }
        case K_REQ * PCS + 3:  // Err("etcdserver: request timed out")
          e.t_pc(t) = 9;
          e.yb = ST_TIMED_OUT, e.yc = 0;
          return RQ_CH_SEND;
        case K_REQ * PCS + 9:
          return RQ_READY;  // tx.send(response).await?; tx, rx drop with the future

        // Campaign: select_biased!{ _ = tx.closed() => return, ret = service.campaign(..) => ret } (server.rs:71-76)
        case K_REQ * PCS + 10: {
          if (e.conn_tx_closed(t)) return RQ_READY;
          if (timed_out(e)) {  // campaign's own timeout().await? sleeps inside the select
            e.t_pc(t) = 11;
            e.xa = e.sample_inclusive_u64(5 * NS_PER_SEC, 15 * NS_PER_SEC - 1);
            return RQ_SLEEP_OR_CLOSED;
          }
          u32 r0 = e.t_a(t);
          int sub;
          Rsp r = sv_campaign(e, (r0 >> 4) & 15, r0 >> 8, e.t_b(t), &sub);
          if (!e.running()) return RQ_PENDING;
          if (sub >= 0) {  // not the leader yet: wait for events of the prefix
            e.t_c(t) = (u32)sub;
            e.t_pc(t) = 12;
            continue;
          }
          e.t_pc(t) = 9;
          e.yb = r.v0(), e.yc = r.data;
          return RQ_CH_SEND;
        }
        case K_REQ * PCS + 11:
          if (e.ret == -2) return RQ_READY;  // the client went away during the sleep
          e.t_pc(t) = 9;
          e.yb = ST_TIMED_OUT, e.yc = 0;
          return RQ_CH_SEND;
        case K_REQ * PCS + 12: {  // loop { rx.recv().await.expect(..); leader = inner.leader(name); .. } inside the select
          int s = (int)e.t_c(t);
          if (e.closed_or_register(t)) {
            sub_drop_rx(e, s);
            return RQ_READY;
          }
          int got = sub_recv(e, s, t);
          if (got < 0) return RQ_PENDING;  // parked on the watcher event and on tx.closed()
          if (got == 0) return panic(e, 81);  // .expect("sender should not drop")
          u32 r0 = e.t_a(t), name = (r0 >> 4) & 15;
          int ld = leader(e, name);
          Rsp r;
          if (ld < 0) {
            r.status = ST_SESSION_EXPIRED;
          } else if (e.wl_s(S_KV + 6 * ld + 2) == e.t_b(t)) {
            r.flag = 1, r.revision = e.wl_s(S_REV), r.data = e.wl_s(S_KV + 6 * ld + 5);  // rev: kv.modify_revision
          } else {
            continue;  // someone else leads: the next event
          }
          e.closed_unregister(t);
          sub_drop_rx(e, s);
          e.t_pc(t) = 9;
          e.yb = r.v0(), e.yc = r.data;
          return RQ_CH_SEND;
        }

        // Observe (server.rs:81-99): service.observe(name).await, then the event loop racing tx.closed()
        case K_REQ * PCS + 20: {
          if (timed_out(e)) {
            e.t_pc(t) = 3;
            e.xa = e.sample_inclusive_u64(5 * NS_PER_SEC, 15 * NS_PER_SEC - 1);
            return RQ_SLEEP_FOR;
          }
          u32 name = (e.t_a(t) >> 4) & 15;
          int s = subscribe(e, name);
          if (s < 0) return RQ_PENDING;
          e.t_c(t) = (u32)s;
          e.t_dl1(t) = leader_snapshot(e, name);
          e.t_pc(t) = 21;
        }
          // fallthrough
        case K_REQ * PCS + 21: {
          int s = (int)e.t_c(t);
          if (e.closed_or_register(t)) {
            sub_drop_rx(e, s);
            return RQ_READY;
          }
          int got = sub_recv(e, s, t);
          if (got < 0) return RQ_PENDING;
          e.closed_unregister(t);  // this select is over (the next one registers again)
          if (got == 0) {  // elem.is_none(): the bus dropped us
            sub_drop_rx(e, s);
            return RQ_READY;
          }
          u32 name = (e.t_a(t) >> 4) & 15;
          u32 now_leader = leader_snapshot(e, name);
          if (now_leader == (u32)e.t_dl1(t)) continue;  // the leader did not change
          e.t_dl1(t) = now_leader;
          Rsp r = leader_rsp(e, name);
          e.t_pc(t) = 22;
          e.yb = r.v0(), e.yc = r.data;
          return RQ_CH_SEND;
        }
        case K_REQ * PCS + 22:
          if (e.ret < 0) {  // tx.send(..).await?
            sub_drop_rx(e, (int)e.t_c(t));
            return RQ_READY;
          }
          e.t_pc(t) = 21;
          continue;

        // ---- the three clients: Client::connect, then the script ----
        case K_C1 * PCS + 0:
        case K_C2 * PCS + 0:
        case K_C3 * PCS + 0:
          e.wl_w(F_STEP, e.t_node(t)) = 0;
          e.t_pc(t) = 1;
          e.ya = mk_addr(IPK_UNSPEC, 0, 0);
          return RQ_BIND;
        case K_C1 * PCS + 1:
        case K_C2 * PCS + 1:
        case K_C3 * PCS + 1:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          // fallthrough
        case K_C1 * PCS + 2:
        case K_C2 * PCS + 2:
        case K_C3 * PCS + 2: {
          int what = script(e, t, e.t_kind(t));
          if (what == 1) return RQ_SLEEP_FOR;  // (the step already moved on)
          if (what == 2) return RQ_READY;
          if (what == 3) {  // leader_stream.message()
            e.t_pc(t) = 8;
            return RQ_CH_RECV;
          }
          if (what == 4) {  // client.observe("leader"): the stream's connection stays with the task
            e.t_pc(t) = 6;
            e.ya = SERVER_ADDR;
            return RQ_CONNECT1;
          }
          e.conn_swap(t);  // a call uses a connection of its own: park the stream (if any)
          e.t_pc(t) = 3;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        }
        case K_C1 * PCS + 3:
        case K_C2 * PCS + 3:
        case K_C3 * PCS + 3:
          if (e.ret < 0) return panic(e, 0);  // connect1(..).await? then .unwrap()
          e.t_pc(t) = 4;
          e.yb = e.t_a(t), e.yc = e.t_b(t);
          return RQ_CH_SEND;
        case K_C1 * PCS + 4:
        case K_C2 * PCS + 4:
        case K_C3 * PCS + 4:
          if (e.ret < 0) return panic(e, 65);
          e.t_pc(t) = 5;
          return RQ_CH_RECV;
        case K_C1 * PCS + 5:
        case K_C2 * PCS + 5:
        case K_C3 * PCS + 5: {
          if (e.ret < 0) return panic(e, 66);
          u32 node = e.t_node(t);
          e.conn_drop_end(t);  // the call's tx, rx
          e.conn_swap(t);      // back to the stream
          int bad = check(e, e.t_kind(t), node, e.rv0, e.rv1);
          if (bad) return panic(e, (u32)bad);
          e.wl_w(F_STEP, node)++;
          e.t_pc(t) = 2;
          continue;
        }
        case K_C3 * PCS + 6:  // observe: connect1, then tx.send(Observe{name})
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 7;
          e.yb = OP_OBSERVE | 0u << 4, e.yc = 0;
          return RQ_CH_SEND;
        case K_C3 * PCS + 7:
          if (e.ret < 0) return panic(e, 65);
          e.wl_w(F_STEP, e.t_node(t))++;
          e.t_pc(t) = 2;
          continue;
        case K_C3 * PCS + 8: {  // one message of the ObserveStream
          if (e.ret < 0) return panic(e, 90);
          u32 node = e.t_node(t), step = e.wl_w(F_STEP, node);
          u32 want = step == 1 ? 10u : step == 2 ? 11u : step == 3 ? 12u : 20u;
          u32 v0 = e.rv0, v1 = e.rv1;
          if ((v0 & 15) != ST_OK) return panic(e, 90);
          if (!((v0 >> 4) & 1) || (v1 & 0xffff) != want) return panic(e, 91);
          e.wl_s(S_SEEN)++;
          e.wl_s(S_SUM) = (e.wl_s(S_SUM) ^ (v1 + ((v0 >> 8) << 8))) * 0x9e3779b1u;
          e.wl_w(F_STEP, node) = step + 1;
          e.t_pc(t) = 2;
          continue;
        }
        default:
          return panic(e, 0xbad);
      }
  }
};

}  // namespace msim
