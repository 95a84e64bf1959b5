// MSIM_WL_ETCD — BASELINE config 4 / SURVEY §8(f) row 3: the simulated etcd server of
// madsim-etcd-client and its one-connection-per-request clients, with a linearizability check.
//
//   SimServer::serve        madsim-etcd-client/src/server.rs:34-101  (accept1 loop, task per request)
//   EtcdService::new        src/service.rs:20-39                     (1 s lease tick task)
//   EtcdService::timeout    src/service.rs:165-176                   (gen_bool(rate), gen_range(5 s..15 s), sleep)
//   ServiceInner            src/service.rs:282-492                   (put / get / delete / txn / lease_grant / tick)
//   KvClient / LeaseClient  src/kv.rs:26-75, src/lease.rs:30-88      (connect1 -> send -> recv per call; keep_alive: a ping / response loop)
//
// Oracle twin: oracle/workloads_etcd.hpp (same programs as coroutines on the reference's API, the
// store in std::map; here the store is a fixed table in the seed's state block).  Keys are
// 0..4 (key 4 is only ever written under lease 7), values are u16, one lease may be live.
#pragma once
#include "../engine/engine.h"

namespace msim {

/// ADMIN: the clients also use the rest of LeaseClient (revoke, keep_alive, time_to_live, leases: p7).  ADMIN = false
/// compiles those paths out; measured, the smaller kernel is 5 % SLOWER on config 4 (code layout against the 32 KB
/// instruction cache, profiles/r2e_etcd_variants.txt), so every ETCD run uses the full program.
template <bool ADMIN>
struct WL_EtcdT {
  static constexpr int NKEY = 5, LEASE_KEY = 4, LEASE_ID = 7, MAX_HIST = 128;
  enum : int {  // wl_s scalars
    S_REV = 0, S_MAXREV, S_NOK, S_NERR, S_NHIST, S_CNT, S_LEASE_ON, S_LEASE_ID, S_LEASE_TTL, S_LEASE_KEYS, S_LEASE_GRANTED,
    S_KV = 11,             // [NKEY*2] value | lease<<16 | present<<31 ; create_rev | mod_rev<<16
    S_CACHE = S_KV + 2 * NKEY,  // [NKEY] bug mode read cache: value | mod_rev<<16 (0 = empty)
    S_JH = S_CACHE + NKEY,      // [3] the clients' JoinHandles (main)
    S_HIST = S_JH + 3,          // [MAX_HIST*3]
    S_END = S_HIST + 3 * MAX_HIST
  };
  enum : int { F_I = 0, F_NEED_GRANT, F_SEEN01, F_SEEN23, F_SEEN4, F_REQ0, F_REQ1, F_FLOOR, F_KA_PHASE, F_NFIELD };  // wl_w per client node
  struct Caps {
    static constexpr int NT = 16, TQ = 48, NMSG = 2, NSOCK = 4, NREG = 2, NMBOX = 2, NNODE = 5, NPARK = 1;
    static constexpr int WQ = 0, WW = F_NFIELD, WS = S_END;
    static constexpr bool USES_JOIN = true;
    static constexpr bool USES_KILL = false;
    static constexpr bool USES_CTRLC = false;  // does any task await signal::ctrl_c()?
    static constexpr bool USES_CLOSED = false;  // does any task await Sender::closed()?
    static constexpr bool USES_SPLIT = false;   // does any task give the Sender half of its connection to another task?
    static constexpr bool USES_TIMEOUT = false;  // does any task wrap awaits in time::timeout (beyond call_timeout)?
    static constexpr bool USES_CONN2 = false;    // does any task hold two connection ends at once?
    static constexpr bool USES_PMATCH = false;   // does any node use restart_on_panic_matching?
    static constexpr unsigned PMATCH_BASE = 0;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    // spawn (once per request) out of line: +10 % here (profiles/r2d_cold_paths.txt); the other two pieces lose
    static constexpr bool COLD_NOW_SMEM = true;  // ... and with the clock passed through the state block around it: +7 % (profiles/r2s)
    static constexpr bool COLD_SPAWN = true, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;
    static constexpr int NCONN = 6, NCQ = 2, NACC = 4;
  };
  enum : u32 { K_MAIN = 0, K_SERVE, K_TICK, K_REQ, K_CLIENT };
  enum : u32 { OP_REVOKE = 0, OP_PUT = 1, OP_GET = 2, OP_DELETE = 3, OP_CAS = 4, OP_GRANT = 5, OP_LEASE_READ = 6, OP_KEEP_ALIVE = 7 };
  enum : u32 { ST_OK = 0, ST_TIMED_OUT = 1, ST_LEASE_NOT_FOUND = 2, ST_REFUSED = 3 };
  enum : u32 { BUG_STALE_READ = 1, BUG_LEASE_KEY_DELETE = 2 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 2379u;  // 10.0.0.1:2379
  static MSIM_HD int init_kind(int) { return -1; }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }

  // ---------------- ServiceInner (service.rs:282-492) on the fixed table ----------------
  struct Rsp {
    u32 status = ST_OK, flag = 0, revision = 0, value = 0, mod_revision = 0;
    int panic_code = 0;
    MSIM_HD u32 v0() const { return status | flag << 4 | revision << 16; }
    MSIM_HD u32 v1() const { return value | mod_revision << 16; }
  };
  template <class E>
  static MSIM_HD bool lease_exists(E& e, u32 id) {
    return e.wl_s(S_LEASE_ON) && e.wl_s(S_LEASE_ID) == id;
  }
  template <class E>
  static MSIM_HD Rsp sv_put(E& e, u32 key, u32 value, u32 lease_id) {
    Rsp r;
    u32 w0 = e.wl_s(S_KV + 2 * key), w1 = e.wl_s(S_KV + 2 * key + 1);
    bool present = w0 >> 31;
    u32 prev_lease = (w0 >> 16) & 0xff;
    if (lease_id != 0) {
      if (!lease_exists(e, lease_id)) {
        r.status = ST_LEASE_NOT_FOUND;
        return r;
      }
      e.wl_s(S_LEASE_KEYS) |= 1u << key;
    }
    if (present && prev_lease != 0) {
      if (!lease_exists(e, prev_lease)) {  // .expect("no lease")
        r.panic_code = 68;
        return r;
      }
      e.wl_s(S_LEASE_KEYS) &= ~(1u << key);
    }
    u32 rev = ++e.wl_s(S_REV);
    u32 create = present ? (w1 & 0xffff) : rev;
    e.wl_s(S_KV + 2 * key) = value | lease_id << 16 | 1u << 31;
    e.wl_s(S_KV + 2 * key + 1) = create | rev << 16;
    if (e.P->p[3] & BUG_STALE_READ) e.wl_s(S_CACHE + key) = value | rev << 16;
    r.revision = rev;
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_get(E& e, u32 key) {
    Rsp r;
    r.revision = e.wl_s(S_REV);
    u32 w0 = e.wl_s(S_KV + 2 * key), w1 = e.wl_s(S_KV + 2 * key + 1);
    if (w0 >> 31) r.flag = 1, r.value = w0 & 0xffff, r.mod_revision = w1 >> 16;
    if ((e.P->p[3] & BUG_STALE_READ) && e.wl_s(S_CACHE + key)) {
      u32 c = e.wl_s(S_CACHE + key);
      r.flag = 1, r.value = c & 0xffff, r.mod_revision = c >> 16;
    }
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_delete(E& e, u32 key) {
    Rsp r;
    u32 w0 = e.wl_s(S_KV + 2 * key);
    if (w0 >> 31) {
      e.wl_s(S_KV + 2 * key) = 0;
      e.wl_s(S_KV + 2 * key + 1) = 0;
      e.wl_s(S_REV)++;
      u32 lease = (w0 >> 16) & 0xff;
      if (lease != 0) {
        if (!lease_exists(e, lease)) {
          r.panic_code = 68;
          return r;
        }
        e.wl_s(S_LEASE_KEYS) &= ~(1u << key);
      }
      r.flag = 1;
    }
    r.revision = e.wl_s(S_REV);
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_txn_cas(E& e, u32 key, u32 expect, u32 value) {
    u32 w0 = e.wl_s(S_KV + 2 * key);
    bool succeeded = (w0 >> 31) && (w0 & 0xffff) == expect;
    u32 rev0 = e.wl_s(S_REV);
    Rsp inner = succeeded ? sv_put(e, key, value, 0) : sv_get(e, key);
    if (inner.panic_code) return inner;
    e.wl_s(S_REV) = rev0 + 1;
    Rsp r;
    r.revision = rev0 + 1;
    r.flag = succeeded;
    if (!succeeded) r.value = inner.flag ? inner.value : 0, r.mod_revision = inner.mod_revision;
    return r;
  }
  template <class E>
  static MSIM_HD Rsp sv_lease_grant(E& e, u32 ttl, u32 id) {
    Rsp r;
    if (e.wl_s(S_LEASE_ON)) {
      if (e.wl_s(S_LEASE_ID) == id) {
        r.panic_code = 62;  // "lease ID already exists"
      } else {
        e.fail_cap(MSIM_CAP_SOCKETS);  // one live lease per store in this build
      }
      return r;
    }
    e.wl_s(S_LEASE_ON) = 1, e.wl_s(S_LEASE_ID) = id, e.wl_s(S_LEASE_TTL) = ttl, e.wl_s(S_LEASE_KEYS) = 0;
    e.wl_s(S_LEASE_GRANTED) = ttl;
    r.revision = ++e.wl_s(S_REV);
    return r;
  }
  /// lease_revoke (service.rs:423-435)
  template <class E>
  static MSIM_HD Rsp sv_lease_revoke(E& e, u32 id) {
    Rsp r;
    if (!lease_exists(e, id)) {
      r.status = ST_LEASE_NOT_FOUND;
      return r;
    }
    u32 keys = e.wl_s(S_LEASE_KEYS);
    for (u32 k = 0; k < (u32)NKEY; k++) {
      if (!(keys >> k & 1)) continue;
      if (!(e.wl_s(S_KV + 2 * k) >> 31)) {
        r.panic_code = 63;  // .expect("no key")
        return r;
      }
      e.wl_s(S_KV + 2 * k) = 0;
      e.wl_s(S_KV + 2 * k + 1) = 0;
    }
    e.wl_s(S_LEASE_ON) = 0;
    r.revision = ++e.wl_s(S_REV);
    return r;
  }
  /// lease_keep_alive (service.rs:437-448): ttl = granted_ttl, and the revision moves
  template <class E>
  static MSIM_HD Rsp sv_lease_keep_alive(E& e, u32 id) {
    Rsp r;
    if (!lease_exists(e, id)) {
      r.status = ST_LEASE_NOT_FOUND;
      return r;
    }
    e.wl_s(S_LEASE_TTL) = e.wl_s(S_LEASE_GRANTED);
    r.revision = ++e.wl_s(S_REV);
    r.value = e.wl_s(S_LEASE_GRANTED);
    return r;
  }
  /// lease_time_to_live (what = 0) / lease_leases (what = 1) (service.rs:450-470)
  template <class E>
  static MSIM_HD Rsp sv_lease_read(E& e, u32 what, u32 id) {
    Rsp r;
    r.revision = e.wl_s(S_REV);
    if (what == 1) {
      r.value = e.wl_s(S_LEASE_ON) ? 1 : 0;
      return r;
    }
    if (!lease_exists(e, id)) {
      r.status = ST_LEASE_NOT_FOUND;
      r.revision = 0;
      return r;
    }
    r.value = e.wl_s(S_LEASE_TTL) & 0xffff;
    r.flag = 1;
    return r;
  }
  /// tick (service.rs:472-492).  0 or a panic code
  template <class E>
  static MSIM_HD int sv_tick(E& e) {
    if (!e.wl_s(S_LEASE_ON)) return 0;
    int ttl = (int)e.wl_s(S_LEASE_TTL) - 1;
    e.wl_s(S_LEASE_TTL) = (u32)ttl;
    if (ttl > 0) return 0;
    u32 keys = e.wl_s(S_LEASE_KEYS);
    for (u32 k = 0; k < (u32)NKEY; k++) {
      if (!(keys >> k & 1)) continue;
      if (!(e.wl_s(S_KV + 2 * k) >> 31)) return 63;  // .expect("no key")
      e.wl_s(S_KV + 2 * k) = 0;
      e.wl_s(S_KV + 2 * k + 1) = 0;
    }
    e.wl_s(S_LEASE_ON) = 0;
    e.wl_s(S_REV)++;
    return 0;
  }

  // ---------------- client side ----------------
  template <class E>
  static MSIM_HD u32 seen_get(E& e, u32 node, u32 key) {
    u32 wv = e.wl_w(F_SEEN01 + (key >> 1), node);
    return (key & 1) ? wv >> 16 : wv & 0xffff;
  }
  template <class E>
  static MSIM_HD void seen_set(E& e, u32 node, u32 key, u32 v) {
    u32& wv = e.wl_w(F_SEEN01 + (key >> 1), node);
    wv = (key & 1) ? ((wv & 0xffffu) | v << 16) : ((wv & 0xffff0000u) | v);
  }
  /// everything etcd_client does with a response (or a refused connect)
  template <class E>
  static MSIM_HD int client_complete(E& e, u32 node, u32 status, u32 flag, u32 revision, u32 value, u32 mod_revision) {
    u32 c = node - 2;
    u32 r0 = e.wl_w(F_REQ0, node), r1 = e.wl_w(F_REQ1, node);
    u32 op = r0 & 15, key = (r0 >> 4) & 15, req_value = r1 & 0xffff;
    bool read = op == OP_GET || (op == OP_CAS && !flag);
    bool carries_value = read || op == OP_LEASE_READ || op == OP_KEEP_ALIVE;
    if (status == ST_OK) {
      e.wl_s(S_NOK)++;
      if (revision > e.wl_s(S_MAXREV)) e.wl_s(S_MAXREV) = revision;
      if (op == OP_GRANT) e.wl_w(F_NEED_GRANT, node) = 0;
      if (op == OP_PUT || (op == OP_CAS && flag)) seen_set(e, node, key, req_value);
      if (read) seen_set(e, node, key, value);
      if (op == OP_DELETE) seen_set(e, node, key, 0);
      if (op == OP_REVOKE) e.wl_w(F_NEED_GRANT, node) = c == 0;
    } else {
      e.wl_s(S_NERR)++;
      if (status == ST_LEASE_NOT_FOUND && c == 0) e.wl_w(F_NEED_GRANT, node) = 1;
    }
    u32 n = e.wl_s(S_NHIST);
    if (n >= (u32)MAX_HIST) return 67;
    u32 hv = carries_value ? value : req_value;
    e.wl_s(S_HIST + 3 * n) = c | op << 2 | key << 5 | status << 8 | flag << 10 | hv << 16;
    e.wl_s(S_HIST + 3 * n + 1) = revision | mod_revision << 16;
    e.wl_s(S_HIST + 3 * n + 2) = e.wl_w(F_FLOOR, node);
    e.wl_s(S_NHIST) = n + 1;
    return 0;
  }
  /// etcd_check_history: 0 or the panic code
  template <class E>
  static MSIM_HD int check_history(E& e) {
    u32 n = e.wl_s(S_NHIST);
MSIM_NOUNROLL
    for (u32 i = 0; i < n; i++) {
      u32 w0 = e.wl_s(S_HIST + 3 * i), w1 = e.wl_s(S_HIST + 3 * i + 1), floor = e.wl_s(S_HIST + 3 * i + 2);
      u32 op = (w0 >> 2) & 7, key = (w0 >> 5) & 7, status = (w0 >> 8) & 3, flag = (w0 >> 10) & 1, value = w0 >> 16;
      u32 revision = w1 & 0xffff, mod_revision = w1 >> 16;
      if (status != ST_OK) continue;
      bool changes = op == OP_PUT || op == OP_GRANT || op == OP_CAS || (op == OP_DELETE && flag) || op == OP_REVOKE || op == OP_KEEP_ALIVE;
      if (revision < floor || (changes && revision <= floor)) return 60;
      bool is_read = op == OP_GET || (op == OP_CAS && !flag);
      if (!is_read || key == (u32)LEASE_KEY) continue;
      u32 at = op == OP_GET ? revision : revision - 1;
      u32 best_rev = 0, best_val = 0;
      bool best_found = false;
MSIM_NOUNROLL
      for (u32 j = 0; j < n; j++) {
        u32 x0 = e.wl_s(S_HIST + 3 * j), x1 = e.wl_s(S_HIST + 3 * j + 1);
        u32 wop = (x0 >> 2) & 7, wkey = (x0 >> 5) & 7, wstatus = (x0 >> 8) & 3, wflag = (x0 >> 10) & 1;
        u32 wrev = x1 & 0xffff;
        if (wstatus != ST_OK || wkey != key) continue;
        bool is_write = wop == OP_PUT || (wop == OP_CAS && wflag) || (wop == OP_DELETE && wflag);
        if (!is_write || wrev > at || wrev < best_rev) continue;
        best_rev = wrev;
        best_found = wop != OP_DELETE;
        best_val = x0 >> 16;
      }
      bool found = op == OP_GET ? flag != 0 : value != 0;
      if (found != best_found) return 61;
      if (found && (value != best_val || mod_revision != best_rev)) return 61;
    }
    return 0;
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const Params& P = *e.P;
    const u32 n_clients = P.n_nodes - 1;
    for (;;) switch (e.t_kind(t) * 16 + e.t_pc(t)) {
        // ---- wl_etcd_main ----
        case K_MAIN * 16 + 0:  // server node, then the client nodes
          if (e.wl_s(S_CNT) < 1 + n_clients) {
            e.wl_s(S_CNT)++;
            e.ya = NODE_IP;
            return RQ_CREATE_NODE;
          }
          e.t_pc(t) = 1;  // server_node.spawn(etcd_serve)
          e.ya = K_SERVE, e.yb = 1, e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_MAIN * 16 + 1:
          e.t_pc(t) = 2;
          e.wl_s(S_CNT) = 0;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        case K_MAIN * 16 + 3:  // the JoinHandle of client S_CNT
          e.wl_s(S_JH + e.wl_s(S_CNT)) = (u32)e.ret;
          e.wl_s(S_CNT)++;
          e.t_pc(t) = 2;
          // fallthrough
        case K_MAIN * 16 + 2:  // clients[c].spawn(etcd_client(c))
          if (e.wl_s(S_CNT) < n_clients) {
            e.t_pc(t) = 3;
            e.ya = K_CLIENT | SPAWN_JOINABLE, e.yb = 2 + e.wl_s(S_CNT), e.yc = e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          e.wl_s(S_CNT) = 0;
          e.t_pc(t) = 4;
          // fallthrough
        case K_MAIN * 16 + 4:  // for t in tasks { t.await.unwrap() }
          if (e.wl_s(S_CNT) < n_clients) {
            e.t_pc(t) = 5;
            e.ya = e.wl_s(S_JH + e.wl_s(S_CNT));
            return RQ_JOIN;
          }
          e.t_pc(t) = 6;
          continue;
        case K_MAIN * 16 + 5:
          if (e.ret == 2) return panic(e, 0);
          e.wl_s(S_CNT)++;
          e.t_pc(t) = 4;
          continue;
        case K_MAIN * 16 + 6: {
          int bad = check_history(e);
          if (bad) return panic(e, (u32)bad);
          u32 sum = 0, n = e.wl_s(S_NHIST);
MSIM_NOUNROLL
          for (u32 i = 0; i < n; i++) {
            sum = (sum ^ e.wl_s(S_HIST + 3 * i)) * 0x9e3779b1u;
            sum = (sum ^ e.wl_s(S_HIST + 3 * i + 1)) * 0x9e3779b1u;
          }
          e.w(L::W_USER0) = e.wl_s(S_NOK);
          e.w(L::W_USER1) = e.wl_s(S_MAXREV);
          e.w(L::W_USER2) = sum;
          e.w(L::W_USER3) = e.wl_s(S_NERR);
          return RQ_READY;
        }

        // ---- etcd_serve: bind; EtcdService::new (spawns the tick task); accept1 loop ----
        case K_SERVE * 16 + 0:
          e.t_pc(t) = 1;
          e.ya = SERVER_ADDR;
          return RQ_BIND;
        case K_SERVE * 16 + 1:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          e.ya = K_TICK, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_SERVE * 16 + 2:
          e.t_pc(t) = 3;
          return RQ_ACCEPT1;
        case K_SERVE * 16 + 3:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          e.ya = K_REQ | SPAWN_CONN, e.yb = e.t_node(t), e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;

        // ---- tick task: loop { inner.tick(); sleep(1 s) } ----
        case K_TICK * 16 + 0: {
          int bad = sv_tick(e);
          if (bad) return panic(e, (u32)bad);
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        }

        // ---- etcd_request_task ----
        case K_REQ * 16 + 0:
          e.t_pc(t) = 1;
          return RQ_CH_RECV;
        case K_REQ * 16 + 5:  // LeaseKeepAlive: rx.recv().await? — the keeper's next ping — then answer again
          if (!ADMIN) return panic(e, 0xbad);
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 1;
          e.rv0 = e.t_a(t), e.rv1 = e.t_b(t);
          // fallthrough
        case K_REQ * 16 + 1: {
          if (e.ret < 0) return RQ_READY;
          e.t_a(t) = e.rv0;
          e.t_b(t) = e.rv1;
          // EtcdService::timeout: gen_bool(timeout_rate) — p == 1.0 draws nothing
          bool timed_out = P.p[5] ? true : e.gen_bool_int(P.p[6]);
          if (timed_out) {
            e.t_pc(t) = 2;
            e.xa = e.sample_inclusive_u64(5 * NS_PER_SEC, 15 * NS_PER_SEC - 1);  // gen_range(5 s..15 s)
            return RQ_SLEEP_FOR;
          }
          u32 r0 = e.t_a(t), r1 = e.t_b(t);
          u32 op = r0 & 15, key = (r0 >> 4) & 15, lease = (r0 >> 8) & 255, ttl = r0 >> 16;
          u32 value = r1 & 0xffff, expect = r1 >> 16;
          Rsp r;
          if (!ADMIN && (op == OP_REVOKE || op == OP_KEEP_ALIVE || op == OP_LEASE_READ)) return panic(e, 64);  // no client sends these
          switch (op) {
            case OP_PUT: r = sv_put(e, key, value, lease); break;
            case OP_GET: r = sv_get(e, key); break;
            case OP_DELETE: r = sv_delete(e, key); break;
            case OP_CAS: r = sv_txn_cas(e, key, expect, value); break;
            case OP_GRANT: r = sv_lease_grant(e, ttl, lease); break;
            case OP_REVOKE: if (ADMIN) r = sv_lease_revoke(e, lease); break;
            case OP_KEEP_ALIVE: if (ADMIN) r = sv_lease_keep_alive(e, lease); break;
            case OP_LEASE_READ: if (ADMIN) r = sv_lease_read(e, key, lease); break;
            default: r.panic_code = 64;
          }
          if (r.panic_code) return panic(e, (u32)r.panic_code);
          e.t_pc(t) = 3;
          e.yb = r.v0(), e.yc = r.v1();
          return RQ_CH_SEND;
        }
        case K_REQ * 16 + 2:  // Err("etcdserver: request timed out")
          e.t_pc(t) = 3;
          e.yb = ST_TIMED_OUT, e.yc = 0;
          return RQ_CH_SEND;
        case K_REQ * 16 + 3:
          if (ADMIN && (e.t_a(t) & 15) == OP_KEEP_ALIVE) {  // server.rs:65-69: loop { send(response)?; rx.recv()?; }
            if (e.ret < 0) return RQ_READY;
            e.t_pc(t) = 5;
            return RQ_CH_RECV;
          }
          return RQ_READY;  // tx, rx drop with the future

        // ---- etcd_client ----
        case K_CLIENT * 16 + 0:  // Client::connect = Endpoint::connect = bind("0.0.0.0:0")
          e.t_pc(t) = 1;
          e.ya = mk_addr(IPK_UNSPEC, 0, 0);
          return RQ_BIND;
        case K_CLIENT * 16 + 1: {
          if (e.ret < 0) return panic(e, 0);
          u32 node = e.t_node(t);
          e.wl_w(F_I, node) = 0;
          e.wl_w(F_NEED_GRANT, node) = node == 2;
          e.t_a(t) = 0;  // n_ok
          e.t_pc(t) = 2;
        }
          // fallthrough
        case K_CLIENT * 16 + 2:  // for i in 0..ops { r = gen_range(0..100)
          if ((u64)e.wl_w(F_I, e.t_node(t)) >= P.p[0]) return RQ_READY;
          e.t_pc(t) = 3;
          e.xa = 0, e.xb = 100;
          return RQ_GEN_RANGE;
        case K_CLIENT * 16 + 3:  // key = gen_range(0..NKEY)
          e.t_b(t) = (u32)e.ret64;
          e.t_pc(t) = 4;
          e.xa = 0, e.xb = NKEY;
          return RQ_GEN_RANGE;
        case K_CLIENT * 16 + 4: {
          u32 node = e.t_node(t), c = node - 2, i = e.wl_w(F_I, node);
          u32 r = e.t_b(t), key = (u32)e.ret64;
          u32 op, lease = 0, ttl = 0, expect = 0, value = (c + 1) << 12 | (i + 1);
          if (e.wl_w(F_NEED_GRANT, node)) {
            op = OP_GRANT, key = 0, lease = LEASE_ID, ttl = (u32)P.p[4], value = 0;
          } else if (key == (u32)LEASE_KEY) {
            if (r < 50) {
              op = OP_PUT, lease = LEASE_ID;
            } else if (r < 90 || !(P.p[3] & BUG_LEASE_KEY_DELETE)) {
              op = OP_GET;
            } else {
              op = OP_DELETE;
            }
          } else if (r < 40) {
            op = OP_PUT;
          } else if (r < 75) {
            op = OP_GET;
          } else if (r < 85) {
            op = OP_DELETE;
          } else if (ADMIN && P.p[7] && r < 95) {  // the rest of LeaseClient (lease.rs:46-88)
            lease = LEASE_ID, value = 0;
            if (r < 88) {
              op = OP_LEASE_READ, key = 0;
            } else if (r < 91) {
              op = OP_LEASE_READ, key = 1;
            } else if (r < 94 || c != 0) {
              op = OP_KEEP_ALIVE, key = 0;
            } else {
              op = OP_REVOKE, key = 0;
            }
          } else {
            op = OP_CAS, expect = seen_get(e, node, key);
          }
          e.wl_w(F_REQ0, node) = op | key << 4 | lease << 8 | ttl << 16;
          e.wl_w(F_REQ1, node) = value | expect << 16;
          e.wl_w(F_FLOOR, node) = e.wl_s(S_MAXREV);
          e.wl_w(F_KA_PHASE, node) = 0;
          e.t_pc(t) = 5;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        }
        case K_CLIENT * 16 + 5: {
          u32 node = e.t_node(t);
          if (e.ret < 0) {  // ConnectionRefused: the request never left
            int bad = client_complete(e, node, ST_REFUSED, 0, 0, 0, 0);
            if (bad) return panic(e, (u32)bad);
            e.t_pc(t) = 8;
            e.xa = P.p[1];
            return RQ_SLEEP_FOR;
          }
          e.t_pc(t) = 6;
          e.yb = e.wl_w(F_REQ0, node), e.yc = e.wl_w(F_REQ1, node);
          return RQ_CH_SEND;
        }
        case K_CLIENT * 16 + 6:
          if (e.ret < 0) return panic(e, 65);
          e.t_pc(t) = 7;
          return RQ_CH_RECV;
        case K_CLIENT * 16 + 7: {
          if (e.ret < 0) return panic(e, 66);
          u32 node = e.t_node(t);
          u32 v0 = e.rv0, v1 = e.rv1;
          u32 status = v0 & 15;
          if (status == ST_OK) e.t_a(t)++;
          if (ADMIN && (e.wl_w(F_REQ0, node) & 15) == OP_KEEP_ALIVE && e.wl_w(F_KA_PHASE, node) == 0) {
            // LeaseClient::keep_alive's first response is logged on arrival; then LeaseKeeper::keep_alive pings again
            int bad1 = client_complete(e, node, status, (v0 >> 4) & 1, v0 >> 16, v1 & 0xffff, v1 >> 16);
            if (bad1) return panic(e, (u32)bad1);
            e.wl_w(F_FLOOR, node) = e.wl_s(S_MAXREV);
            e.wl_w(F_KA_PHASE, node) = 1;
            e.t_pc(t) = 6;
            e.yb = e.wl_w(F_REQ0, node), e.yc = e.wl_w(F_REQ1, node);
            return RQ_CH_SEND;
          }
          e.conn_drop_end(t);  // the call's tx, rx drop when it returns
          int bad = client_complete(e, node, status, (v0 >> 4) & 1, v0 >> 16, v1 & 0xffff, v1 >> 16);
          if (bad) return panic(e, (u32)bad);
          e.t_pc(t) = 8;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        }
        case K_CLIENT * 16 + 8:
          e.wl_w(F_I, e.t_node(t))++;
          e.t_pc(t) = 2;
          continue;
        default:
          return panic(e, 0xbad);
      }
  }
};

using WL_Etcd = WL_EtcdT<true>;     // p7 = 1: + lease administration
using WL_EtcdKv = WL_EtcdT<false>;  // BASELINE config 4

}  // namespace msim
