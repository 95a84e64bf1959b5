// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// MSIM_WL_ETCD — BASELINE config 4 / SURVEY §8(f) row 3: the simulated etcd server of
// madsim-etcd-client (src/server.rs:34-101 SimServer::serve, src/service.rs:20-39 EtcdService::new
// with its 1 s lease tick task, :165-176 `timeout()` draws, :282-470 ServiceInner) and its
// one-connection-per-request clients (src/kv.rs:26-75, src/lease.rs), driven by a randomised
// KV workload whose history is checked for linearizability against the revision order.
//
// Restated on the oracle's madsim API exactly as the crate is written on madsim's: every client
// call is connect1 -> tx.send(request) -> rx.recv(); the server accepts, spawns a task per
// request, optionally "times out" (gen_bool(timeout_rate), gen_range(5 s..15 s), sleep), applies
// the operation to the store and sends the response.  Keys are small integers, values u16.
#pragma once
#include <array>
#include <map>
#include <set>
#include <vector>

#include "rt/runtime.hpp"

namespace msim_oracle {

constexpr u32 ETCD_NKEY = 5;       // keys 0..3 are plain, key 4 is always written under the lease
constexpr u32 ETCD_LEASE_KEY = 4;
constexpr u32 ETCD_LEASE_ID = 7;
constexpr u32 ETCD_MAX_HIST = 128;
enum : u32 { ETCD_REVOKE = 0, ETCD_PUT = 1, ETCD_GET = 2, ETCD_DELETE = 3, ETCD_CAS = 4, ETCD_GRANT = 5, ETCD_LEASE_READ = 6, ETCD_KEEP_ALIVE = 7 };
// ETCD_LEASE_READ: key 0 = LeaseTimeToLive (value = remaining ttl), key 1 = LeaseLeases (value = number of leases)
enum : u32 { ETCD_OK = 0, ETCD_TIMED_OUT = 1, ETCD_LEASE_NOT_FOUND = 2, ETCD_REFUSED = 3 };
enum : u32 { ETCD_BUG_STALE_READ = 1, ETCD_BUG_LEASE_KEY_DELETE = 2 };

struct EtcdParams {
  u32 n_clients = 2;
  u64 ops = 32;
  u64 think_ns = 50 * MS;
  float timeout_rate = 0.0f;
  u32 bug = 0;
  u32 lease_ttl = 1;
  u32 lease_admin = 0;  // 1: the clients also revoke / keep alive / query the lease (lease.rs:46-88)
};

// ---- wire format (the crate moves Box<dyn Any>; here two u32 words) ----
struct EtcdRequest {
  u32 op = 0, key = 0, lease = 0, ttl = 0, value = 0, expect = 0;
  Payload encode() const {
    Payload p;
    p.v[0] = op | key << 4 | lease << 8 | ttl << 16;
    p.v[1] = value | expect << 16;
    return p;
  }
  static EtcdRequest decode(const Payload& p) {
    EtcdRequest r;
    r.op = p.v[0] & 15, r.key = (p.v[0] >> 4) & 15, r.lease = (p.v[0] >> 8) & 255, r.ttl = p.v[0] >> 16;
    r.value = p.v[1] & 0xffff, r.expect = p.v[1] >> 16;
    return r;
  }
};
struct EtcdResponse {
  u32 status = ETCD_OK;
  u32 flag = 0;      // get: found; delete: deleted; txn: succeeded
  u32 revision = 0;  // header.revision
  u32 value = 0, mod_revision = 0;
  Payload encode() const {
    Payload p;
    p.v[0] = status | flag << 4 | revision << 16;
    p.v[1] = value | mod_revision << 16;
    return p;
  }
  static EtcdResponse decode(const Payload& p) {
    EtcdResponse r;
    r.status = p.v[0] & 15, r.flag = (p.v[0] >> 4) & 1, r.revision = p.v[0] >> 16;
    r.value = p.v[1] & 0xffff, r.mod_revision = p.v[1] >> 16;
    return r;
  }
};

// ---- service.rs:187-470 ServiceInner ----
struct EtcdKeyValue {
  u32 value = 0, lease = 0, create_revision = 0, modify_revision = 0;
};
struct EtcdLease {
  i64 ttl = 0, granted_ttl = 0;
  std::set<u32> keys;
};
struct EtcdServiceInner {
  u32 revision = 0;
  std::map<u32, EtcdKeyValue> kv;   // BTreeMap<Key, KeyValue>
  std::map<u32, EtcdLease> lease;   // HashMap<LeaseId, Lease>; iteration order is not observable here
  std::map<u32, EtcdKeyValue> read_cache;  // bug mode only
  u32 bug = 0;

  /// service.rs:282-325
  EtcdResponse put(u32 key, u32 value, u32 lease_id) {
    EtcdResponse rsp;
    auto prev = kv.find(key);
    if (lease_id != 0) {
      auto l = lease.find(lease_id);
      if (l == lease.end()) {
        rsp.status = ETCD_LEASE_NOT_FOUND;
        return rsp;
      }
      l->second.keys.insert(key);
    }
    if (prev != kv.end() && prev->second.lease != 0) {
      auto old = lease.find(prev->second.lease);
      if (old == lease.end()) panic("no lease", 68);  // .expect("no lease")
      old->second.keys.erase(key);
    }
    revision += 1;
    EtcdKeyValue v;
    v.value = value, v.lease = lease_id;
    v.create_revision = prev != kv.end() ? prev->second.create_revision : revision;
    v.modify_revision = revision;
    kv[key] = v;
    if (bug & ETCD_BUG_STALE_READ) read_cache[key] = v;
    rsp.revision = revision;
    return rsp;
  }
  /// service.rs:327-341
  EtcdResponse get(u32 key) {
    EtcdResponse rsp;
    rsp.revision = revision;
    const EtcdKeyValue* v = nullptr;
    auto it = kv.find(key);
    if (it != kv.end()) v = &it->second;
    if (bug & ETCD_BUG_STALE_READ) {  // injected bug: a read cache that delete forgets to invalidate
      auto c = read_cache.find(key);
      if (c != read_cache.end()) v = &c->second;
    }
    if (v) rsp.flag = 1, rsp.value = v->value, rsp.mod_revision = v->modify_revision;
    return rsp;
  }
  /// service.rs:349-366
  EtcdResponse del(u32 key) {
    EtcdResponse rsp;
    auto it = kv.find(key);
    if (it != kv.end()) {
      EtcdKeyValue v = it->second;
      kv.erase(it);
      revision += 1;
      if (v.lease != 0) {
        auto old = lease.find(v.lease);
        if (old == lease.end()) panic("no lease", 68);  // .expect("no lease")
        old->second.keys.erase(key);
      }
      rsp.flag = 1;
    }
    rsp.revision = revision;
    return rsp;
  }
  /// service.rs:368-403 for Txn{compare: [value(key) == expect], success: [Put], failure: [Get]}
  EtcdResponse txn_cas(u32 key, u32 expect, u32 value) {
    auto it = kv.find(key);
    bool succeeded = it != kv.end() && it->second.value == expect;
    u32 rev0 = revision;
    EtcdResponse inner = succeeded ? put(key, value, 0) : get(key);
    revision = rev0 + 1;
    EtcdResponse rsp;
    rsp.revision = revision;
    rsp.flag = succeeded;
    if (!succeeded) rsp.value = inner.flag ? inner.value : 0, rsp.mod_revision = inner.mod_revision;
    return rsp;
  }
  /// service.rs:405-421 (the client chooses the id: LeaseGrantOptions::with_id)
  EtcdResponse lease_grant(u32 ttl, u32 id) {
    if (lease.count(id)) panic("lease ID already exists", 62);
    EtcdLease l;
    l.ttl = l.granted_ttl = ttl;
    lease[id] = l;
    revision += 1;
    EtcdResponse rsp;
    rsp.revision = revision;
    return rsp;
  }
  /// service.rs:423-435
  EtcdResponse lease_revoke(u32 id) {
    EtcdResponse rsp;
    auto l = lease.find(id);
    if (l == lease.end()) {
      rsp.status = ETCD_LEASE_NOT_FOUND;
      return rsp;
    }
    for (u32 key : l->second.keys)
      if (!kv.erase(key)) panic("no key", 63);
    lease.erase(l);
    revision += 1;
    rsp.revision = revision;
    return rsp;
  }
  /// service.rs:437-448: resets the ttl AND bumps the revision
  EtcdResponse lease_keep_alive(u32 id) {
    EtcdResponse rsp;
    auto l = lease.find(id);
    if (l == lease.end()) {
      rsp.status = ETCD_LEASE_NOT_FOUND;
      return rsp;
    }
    l->second.ttl = l->second.granted_ttl;
    revision += 1;
    rsp.revision = revision;
    rsp.value = u32(l->second.granted_ttl);
    return rsp;
  }
  /// service.rs:450-470 lease_time_to_live / lease_leases
  EtcdResponse lease_read(u32 what, u32 id) {
    EtcdResponse rsp;
    rsp.revision = revision;
    if (what == 1) {
      rsp.value = u32(lease.size());
      return rsp;
    }
    auto l = lease.find(id);
    if (l == lease.end()) {
      rsp.status = ETCD_LEASE_NOT_FOUND;
      rsp.revision = 0;
      return rsp;
    }
    rsp.value = u32(l->second.ttl) & 0xffff;
    rsp.flag = 1;
    return rsp;
  }
  /// service.rs:472-492: called every second
  void tick() {
    size_t origin_len = lease.size();
    for (auto it = lease.begin(); it != lease.end();) {
      it->second.ttl -= 1;
      if (it->second.ttl <= 0) {
        for (u32 key : it->second.keys) {
          if (!kv.erase(key)) panic("no key", 63);
        }
        it = lease.erase(it);
      } else {
        ++it;
      }
    }
    if (lease.size() != origin_len) revision += 1;
  }
};

/// service.rs:11-176 EtcdService
struct EtcdService {
  float timeout_rate;
  std::shared_ptr<EtcdServiceInner> inner;

  /// service.rs:20-39: spawns the lease tick task on the current node
  static std::shared_ptr<EtcdService> create(float timeout_rate, u32 bug) {
    auto s = std::make_shared<EtcdService>();
    s->timeout_rate = timeout_rate;
    s->inner = std::make_shared<EtcdServiceInner>();
    s->inner->bug = bug;
    std::weak_ptr<EtcdServiceInner> weak = s->inner;
    spawn(tick_task(weak));
    return s;
  }
  static Async<Unit> tick_task(std::weak_ptr<EtcdServiceInner> weak) {
    while (auto inner = weak.lock()) {
      inner->tick();
      inner.reset();
      co_await sleep(1 * SEC);
    }
    co_return {};
  }
  /// service.rs:165-176
  Async<bool> timeout() {
    if (thread_rng().gen_bool(double(timeout_rate))) {
      u64 t = thread_rng().gen_range_duration(5 * SEC, 15 * SEC);
      co_await sleep(t);
      co_return true;  // Err("etcdserver: request timed out")
    }
    co_return false;
  }
  /// put / get / delete / txn / lease_grant (service.rs:41-74): timeout().await?, then the store
  Async<EtcdResponse> handle(EtcdRequest req) {
    if (co_await timeout()) {
      EtcdResponse r;
      r.status = ETCD_TIMED_OUT;
      co_return r;
    }
    switch (req.op) {
      case ETCD_PUT: co_return inner->put(req.key, req.value, req.lease);
      case ETCD_GET: co_return inner->get(req.key);
      case ETCD_DELETE: co_return inner->del(req.key);
      case ETCD_CAS: co_return inner->txn_cas(req.key, req.expect, req.value);
      case ETCD_GRANT: co_return inner->lease_grant(req.ttl, req.lease);
      case ETCD_REVOKE: co_return inner->lease_revoke(req.lease);
      case ETCD_KEEP_ALIVE: co_return inner->lease_keep_alive(req.lease);
      case ETCD_LEASE_READ: co_return inner->lease_read(req.key, req.lease);
    }
    panic("unknown request", 64);
  }
};

/// server.rs:46-100: the task spawned per accepted connection
inline Async<Unit> etcd_request_task(Connection conn, std::shared_ptr<EtcdService> service) {
  auto r = co_await conn.rx.recv();
  if (r.is_err()) co_return {};
  EtcdRequest req = EtcdRequest::decode(*r.ok);
  if (req.op == ETCD_KEEP_ALIVE) {  // server.rs:65-69: answer, then wait for the keeper's next ping
    for (;;) {
      EtcdResponse ka = co_await service->handle(req);
      if (conn.tx.send(ka.encode()).is_err()) co_return {};
      if ((co_await conn.rx.recv()).is_err()) co_return {};
    }
  }
  EtcdResponse rsp = co_await service->handle(req);
  (void)conn.tx.send(rsp.encode());  // tx.send(response).await?; the task ends either way
  co_return {};
}
/// server.rs:34-45 SimServer::serve
inline Async<Unit> etcd_serve(SocketAddr addr, float timeout_rate, u32 bug) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  auto service = EtcdService::create(timeout_rate, bug);
  for (;;) {
    Connection conn = std::move((co_await ep.accept1()).unwrap());
    spawn(etcd_request_task(std::move(conn), service));
  }
}

// ---- the test: clients + history ----
struct EtcdHistEntry {
  u32 client, op, key, status, flag, value, revision, mod_revision, floor;
};
struct EtcdHistory {
  std::vector<EtcdHistEntry> entries;  // completion order
  u32 max_done_revision = 0;
  u32 n_ok = 0, n_err = 0;
};

/// kv.rs:26-75 / lease.rs: one connection per request
inline Async<EtcdResponse> etcd_call(const Endpoint& ep, SocketAddr server, EtcdRequest req) {
  auto c = co_await ep.connect1(server);
  if (c.is_err()) {
    EtcdResponse r;
    r.status = ETCD_REFUSED;  // io::ErrorKind::ConnectionRefused: the request never left
    co_return r;
  }
  Connection conn = std::move(*c.ok);
  if (conn.tx.send(req.encode()).is_err()) panic("etcd client: send failed", 65);
  auto r = co_await conn.rx.recv();
  if (r.is_err()) panic("etcd client: connection reset", 66);
  co_return EtcdResponse::decode(*r.ok);
}

/// LeaseClient::keep_alive (lease.rs:56-63) + one more LeaseKeeper::keep_alive / stream.message() round.
/// The first response is logged when it arrives (its own history entry); the second is returned.
inline Async<EtcdResponse> etcd_keep_alive_twice(const Endpoint& ep, SocketAddr server, EtcdRequest req, u32 c,
                                                 std::shared_ptr<EtcdHistory> h, u32* n_ok, u32* floor, bool* need_grant) {
  EtcdResponse refused;
  refused.status = ETCD_REFUSED;
  auto cn = co_await ep.connect1(server);
  if (cn.is_err()) co_return refused;
  Connection conn = std::move(*cn.ok);
  EtcdResponse rsp;
  for (int k = 0; k < 2; k++) {
    if (conn.tx.send(req.encode()).is_err()) panic("etcd client: send failed", 65);
    auto r = co_await conn.rx.recv();
    if (r.is_err()) panic("etcd client: connection reset", 66);
    rsp = EtcdResponse::decode(*r.ok);
    if (k == 1) break;
    if (rsp.status == ETCD_OK) {
      (*n_ok)++, h->n_ok++;
      if (rsp.revision > h->max_done_revision) h->max_done_revision = rsp.revision;
    } else {
      h->n_err++;
      if (rsp.status == ETCD_LEASE_NOT_FOUND && c == 0) *need_grant = true;
    }
    if (h->entries.size() >= ETCD_MAX_HIST) panic("history full", 67);
    h->entries.push_back({c, req.op, req.key, rsp.status, rsp.flag, rsp.value, rsp.revision, rsp.mod_revision, *floor});
    *floor = h->max_done_revision;  // the second ping is invoked now
  }
  co_return rsp;
}

inline Async<u32> etcd_client(u32 c, SocketAddr server, EtcdParams p, std::shared_ptr<EtcdHistory> h) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());  // Client::connect
  GlobalRng& rng = thread_rng();
  std::array<u32, ETCD_NKEY> last_seen{};
  bool need_grant = c == 0;
  u32 n_ok = 0;
  for (u64 i = 0; i < p.ops; i++) {
    u64 r = rng.gen_range_u64(0, 100);
    u32 key = u32(rng.gen_range_u64(0, ETCD_NKEY));
    EtcdRequest req;
    req.key = key;
    req.value = (c + 1) << 12 | u32(i + 1);
    if (need_grant) {
      req.op = ETCD_GRANT, req.key = 0, req.lease = ETCD_LEASE_ID, req.ttl = p.lease_ttl, req.value = 0;
    } else if (key == ETCD_LEASE_KEY) {  // only ever written under the lease
      if (r < 50) {
        req.op = ETCD_PUT, req.lease = ETCD_LEASE_ID;
      } else if (r < 90 || !(p.bug & ETCD_BUG_LEASE_KEY_DELETE)) {
        req.op = ETCD_GET;
      } else {
        req.op = ETCD_DELETE;  // reaches the crate's own `expect("no lease")` once the lease has expired
      }
    } else if (r < 40) {
      req.op = ETCD_PUT;
    } else if (r < 75) {
      req.op = ETCD_GET;
    } else if (r < 85) {
      req.op = ETCD_DELETE;
    } else if (p.lease_admin && r < 95) {  // the rest of LeaseClient
      req.lease = ETCD_LEASE_ID, req.value = 0;
      if (r < 88) {
        req.op = ETCD_LEASE_READ, req.key = 0;  // time_to_live
      } else if (r < 91) {
        req.op = ETCD_LEASE_READ, req.key = 1;  // leases
      } else if (r < 94 || c != 0) {
        req.op = ETCD_KEEP_ALIVE, req.key = 0;
      } else {
        req.op = ETCD_REVOKE, req.key = 0;
      }
    } else {
      req.op = ETCD_CAS, req.expect = last_seen[key];
    }
    u32 floor = h->max_done_revision;
    EtcdResponse rsp;
    if (req.op == ETCD_KEEP_ALIVE) {
      rsp = co_await etcd_keep_alive_twice(ep, server, req, c, h, &n_ok, &floor, &need_grant);
    } else {
      rsp = co_await etcd_call(ep, server, req);
    }
    if (rsp.status == ETCD_OK) {
      n_ok++;
      h->n_ok++;
      if (rsp.revision > h->max_done_revision) h->max_done_revision = rsp.revision;
      if (req.op == ETCD_GRANT) need_grant = false;
      if (req.op == ETCD_PUT || (req.op == ETCD_CAS && rsp.flag)) last_seen[key] = req.value;
      if (req.op == ETCD_GET || (req.op == ETCD_CAS && !rsp.flag)) last_seen[key] = rsp.value;
      if (req.op == ETCD_DELETE) last_seen[key] = 0;
      if (req.op == ETCD_REVOKE) need_grant = c == 0;
    } else {
      h->n_err++;
      if (rsp.status == ETCD_LEASE_NOT_FOUND && c == 0) need_grant = true;
    }
    if (h->entries.size() >= ETCD_MAX_HIST) panic("history full", 67);
    bool carries_value = req.op == ETCD_GET || (req.op == ETCD_CAS && !rsp.flag) || req.op == ETCD_LEASE_READ || req.op == ETCD_KEEP_ALIVE;
    h->entries.push_back({c, req.op, req.key, rsp.status, rsp.flag, carries_value ? rsp.value : req.value,
                          rsp.revision, rsp.mod_revision, floor});
    co_await sleep(p.think_ns);
  }
  co_return n_ok;
}

/// Linearizability against the revision order: (A) real time — an operation's revision is at
/// least every revision already returned when it was invoked, strictly larger if it changed the
/// store; (B) every read of a plain key returns the latest logged write at or below its revision.
inline void etcd_check_history(const EtcdHistory& h) {
  auto changes = [](const EtcdHistEntry& e) {
    return e.op == ETCD_PUT || e.op == ETCD_GRANT || e.op == ETCD_CAS || (e.op == ETCD_DELETE && e.flag) ||
           e.op == ETCD_REVOKE || e.op == ETCD_KEEP_ALIVE;
  };
  for (const auto& e : h.entries) {
    if (e.status != ETCD_OK) continue;
    if (e.revision < e.floor || (changes(e) && e.revision <= e.floor)) panic("stale revision", 60);
    bool is_read = e.op == ETCD_GET || (e.op == ETCD_CAS && !e.flag);
    if (!is_read || e.key == ETCD_LEASE_KEY) continue;
    u32 at = e.op == ETCD_GET ? e.revision : e.revision - 1;  // a failed txn reads before its own bump
    u32 best_rev = 0, best_val = 0;
    bool best_found = false;
    for (const auto& w : h.entries) {
      if (w.status != ETCD_OK || w.key != e.key) continue;
      bool is_write = w.op == ETCD_PUT || (w.op == ETCD_CAS && w.flag) || (w.op == ETCD_DELETE && w.flag);
      if (!is_write || w.revision > at || w.revision < best_rev) continue;
      best_rev = w.revision;
      best_found = w.op != ETCD_DELETE;
      best_val = w.value;
    }
    bool found = e.op == ETCD_GET ? e.flag != 0 : e.value != 0;
    if (found != best_found) panic("read does not match the latest write (presence)", 61);
    if (found && (e.value != best_val || e.mod_revision != best_rev)) panic("read does not match the latest write", 61);
  }
}

inline Async<u32> wl_etcd_main(EtcdParams p, std::array<u32, 4>* user) {
  auto& rt = current_runtime();
  SocketAddr server{ip4(10, 0, 0, 1), 2379};
  auto server_node = rt.create_node().ip(server.ip).build();
  std::vector<Runtime::NodeHandle> clients;
  for (u32 c = 0; c < p.n_clients; c++) clients.push_back(rt.create_node().ip(ip4(10, 0, 0, 2 + c)).build());
  server_node.spawn(etcd_serve(server, p.timeout_rate, p.bug));
  co_await sleep(1 * SEC);
  auto h = std::make_shared<EtcdHistory>();
  std::vector<JoinHandle<u32>> tasks;
  for (u32 c = 0; c < p.n_clients; c++) tasks.push_back(clients[c].spawn(etcd_client(c, server, p, h)));
  for (auto& t : tasks) {
    auto r = co_await t;
    r.unwrap();
  }
  etcd_check_history(*h);
  u32 sum = 0;
  for (const auto& e : h->entries) {
    u32 w = e.client | e.op << 2 | e.key << 5 | e.status << 8 | e.flag << 10 | e.value << 16;
    sum = (sum ^ w) * 0x9e3779b1u;
    sum = (sum ^ (e.revision | e.mod_revision << 16)) * 0x9e3779b1u;
  }
  (*user)[0] = h->n_ok;
  (*user)[1] = h->max_done_revision;
  (*user)[2] = sum;
  (*user)[3] = h->n_err;
  co_return 0;
}

}  // namespace msim_oracle
