// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// MSIM_WL_ELECTION — SURVEY §8(f) row 3, the part of madsim-etcd-client the KV program (workloads_etcd.hpp) leaves
// out: the election API and the watcher behind it.
//   ElectionClient::{campaign, proclaim, leader, observe, resign}   src/election.rs:27-100  (one connection per call;
//                                                                    observe keeps its connection as the stream)
//   SimServer::serve, Campaign / Observe arms                        src/server.rs:71-95    (select_biased!{ tx.closed(), .. })
//   EtcdService::{campaign, proclaim, leader, observe, resign}       src/service.rs:94-151
//   ServiceInner::{campaign, proclaim, leader, observe, resign}      src/service.rs:494-585
//   EventBus (watcher): subscribe / publish over mpsc::channel(4)    src/service.rs:196-245
//   lease_grant with a server-chosen id (random::<i64>().abs())      src/service.rs:398-414
// The program is the reference's own `election()` test (madsim-etcd-client/tests/test.rs:129-232): a first leader that
// campaigns, re-campaigns, proclaims, holds and revokes its lease; a second candidate that waits in campaign() until the
// first lease goes and then resigns; an observer that must see the leader values 1, 1.1, 1.2 and 2 in order.  Its
// assertions are the panics below.  Names are {0: "leader", 1: "invalid_lease"}; values are small integers
// (10 = "1", 11 = "1.1", 12 = "1.2", 13 = "1.3", 20 = "2"); a lease travels as the low 32 bits of its id.
#pragma once
#include <array>
#include <deque>
#include <map>
#include <set>
#include <vector>

#include "rt/runtime.hpp"

namespace msim_oracle {

enum : u32 { EL_GRANT = 1, EL_REVOKE = 2, EL_CAMPAIGN = 3, EL_PROCLAIM = 4, EL_LEADER = 5, EL_OBSERVE = 6, EL_RESIGN = 7, EL_GET_PREFIX = 8 };
enum : u32 { EL_OK = 0, EL_TIMED_OUT = 1, EL_LEASE_NOT_FOUND = 2, EL_SESSION_EXPIRED = 3 };

struct ElectionParams {
  u64 hold_s = 30;  // how long the first leader holds before it revokes its lease (tests/test.rs:177)
  float timeout_rate = 0.0f;
};

struct ElRequest {
  u32 op = 0, name = 0, value = 0, lease = 0;  // GRANT: value = ttl
  Payload encode() const {
    Payload p;
    p.v[0] = op | name << 4 | value << 8;
    p.v[1] = lease;
    return p;
  }
  static ElRequest decode(const Payload& p) {
    ElRequest r;
    r.op = p.v[0] & 15, r.name = (p.v[0] >> 4) & 15, r.value = p.v[0] >> 8, r.lease = p.v[1];
    return r;
  }
};
struct ElResponse {
  u32 status = EL_OK, flag = 0, revision = 0;  // flag: a kv / a leader key is present
  u32 data = 0;  // grant: lease; leader / observe: value | modify_revision << 16; campaign: the key's rev; get: count
  Payload encode() const {
    Payload p;
    p.v[0] = status | flag << 4 | revision << 8;
    p.v[1] = data;
    return p;
  }
  static ElResponse decode(const Payload& p) {
    ElResponse r;
    r.status = p.v[0] & 15, r.flag = (p.v[0] >> 4) & 1, r.revision = p.v[0] >> 8, r.data = p.v[1];
    return r;
  }
};

// ---- tokio::sync::mpsc::channel(4) as the EventBus uses it: try_send on one side, recv().await on the other ----
struct ElEventChan {
  u32 queued = 0;  // the receivers only count arrivals (they re-read the leader on every event)
  bool rx_alive = true, tx_alive = true;
  std::optional<Waker> rx_waker;
  bool try_send() {  // Err(Full) / Err(Closed)
    if (!rx_alive || queued >= 4) return false;
    queued++;
    if (rx_waker) {
      Waker w = std::move(*rx_waker);
      rx_waker.reset();
      w.wake();
    }
    return true;
  }
  void drop_tx() {  // the EventBus dropped its Sender (publish's retain): the receiver sees None after draining
    tx_alive = false;
    if (rx_waker) {
      Waker w = std::move(*rx_waker);
      rx_waker.reset();
      w.wake();
    }
  }
};
struct ElEventRx {
  std::shared_ptr<ElEventChan> ch;
  ElEventRx() = default;
  explicit ElEventRx(std::shared_ptr<ElEventChan> c) : ch(std::move(c)) {}
  ElEventRx(ElEventRx&&) = default;
  ElEventRx& operator=(ElEventRx&&) = default;
  ~ElEventRx() {
    if (ch) ch->rx_alive = false, ch->rx_waker.reset();
  }
  struct RecvFut {
    ElEventChan* ch;
    Poll<bool> poll(Context& cx) {  // true = Some(event), false = None
      if (ch->queued) {
        ch->queued--;
        return Poll<bool>::done(true);
      }
      if (!ch->tx_alive) return Poll<bool>::done(false);
      ch->rx_waker = cx.waker;
      return Poll<bool>::pending();
    }
  };
  RecvFut recv() { return RecvFut{ch.get()}; }
};

// ---- ServiceInner: the election part + what it needs of kv / lease ----
struct ElKeyValue {
  u32 name = 0, lease = 0;  // key = "{name}/{lease:016x}"
  u32 value = 0, create_revision = 0, modify_revision = 0;
  bool operator==(const ElKeyValue& o) const {
    return name == o.name && lease == o.lease && value == o.value && create_revision == o.create_revision &&
           modify_revision == o.modify_revision;
  }
};
struct ElLease {
  u64 id = 0;  // random::<i64>().abs()
  i64 ttl = 0;
  std::vector<std::pair<u32, u32>> keys;  // (name, lease) of the keys attached
};
struct ElServiceInner {
  u32 revision = 0;
  std::vector<ElKeyValue> kv;      // BTreeMap<Key, KeyValue>: looked up by (name, lease), scanned by prefix
  std::vector<ElLease> lease;      // HashMap<LeaseId, Lease>: two leases at most, never expiring in the same tick
  std::vector<std::pair<u32, std::shared_ptr<ElEventChan>>> watcher;  // EventBus::list: (Prefix(name), Sender)

  ElLease* find_lease(u32 lo32) {
    for (auto& l : lease)
      if (u32(l.id) == lo32) return &l;
    return nullptr;
  }
  ElKeyValue* find_kv(u32 name, u32 lease_lo) {
    for (auto& k : kv)
      if (k.name == name && k.lease == lease_lo) return &k;
    return nullptr;
  }
  /// EventBus::publish (service.rs:233-244): a subscriber stays only while try_send succeeds
  void publish(u32 name) {
    std::vector<std::pair<u32, std::shared_ptr<ElEventChan>>> keep;
    for (auto& s : watcher) {
      if (s.first != name || s.second->try_send())
        keep.push_back(s);
      else
        s.second->drop_tx();
    }
    watcher = std::move(keep);
  }
  ElEventRx subscribe(u32 name) {
    auto ch = std::make_shared<ElEventChan>();
    watcher.emplace_back(name, ch);
    return ElEventRx(ch);
  }
  /// service.rs:398-414, id == 0: the server draws one
  ElResponse lease_grant(u32 ttl) {
    u64 id = 0;
    while (id == 0 || find_lease(u32(id))) id = u64(std::llabs(i64(thread_rng().gen_u64())));
    ElLease l;
    l.id = id, l.ttl = ttl;
    lease.push_back(l);
    revision += 1;
    ElResponse r;
    r.revision = revision, r.data = u32(id);
    return r;
  }
  void remove_kv(u32 name, u32 lease_lo) {
    for (size_t i = 0; i < kv.size(); i++)
      if (kv[i].name == name && kv[i].lease == lease_lo) {
        kv.erase(kv.begin() + long(i));
        return;
      }
    panic("no key", 63);
  }
  /// service.rs:416-428
  ElResponse lease_revoke(u32 lo32) {
    ElResponse r;
    ElLease* l = find_lease(lo32);
    if (!l) {
      r.status = EL_LEASE_NOT_FOUND;
      return r;
    }
    auto keys = l->keys;
    for (size_t i = 0; i < lease.size(); i++)
      if (&lease[i] == l) {
        lease.erase(lease.begin() + long(i));
        break;
      }
    for (auto& k : keys) {
      remove_kv(k.first, k.second);
      publish(k.first);  // Event::Delete(kv)
    }
    revision += 1;
    r.revision = revision;
    return r;
  }
  /// service.rs:472-492
  void tick() {
    size_t origin_len = lease.size();
    for (size_t i = 0; i < lease.size();) {
      lease[i].ttl -= 1;
      if (lease[i].ttl <= 0) {
        auto keys = lease[i].keys;
        lease.erase(lease.begin() + long(i));
        for (auto& k : keys) {
          remove_kv(k.first, k.second);
          publish(k.first);
        }
      } else {
        i++;
      }
    }
    if (lease.size() != origin_len) revision += 1;
  }
  /// service.rs:547-556: the key of the prefix with the smallest create_revision
  const ElKeyValue* leader(u32 name) const {
    const ElKeyValue* best = nullptr;
    for (auto& k : kv)
      if (k.name == name && (!best || k.create_revision < best->create_revision)) best = &k;
    return best;
  }
  /// service.rs:494-538.  status OK + won = true: the caller leads now; won = false: subscribed, wait
  ElResponse campaign(u32 name, u32 value, u32 lease_lo, bool* won, ElEventRx* rx) {
    ElResponse r;
    ElKeyValue* cur = find_kv(name, lease_lo);
    if (!cur || cur->value != value) {  // put the kv if key not exists or value not match
      revision += 1;  // (before the lease lookup: a campaign under an unknown lease still moves the revision)
      ElLease* l = find_lease(lease_lo);
      if (!l) {
        r.status = EL_LEASE_NOT_FOUND;
        return r;
      }
      bool attached = false;
      for (auto& k : l->keys) attached = attached || (k.first == name && k.second == lease_lo);
      if (!attached) l->keys.emplace_back(name, lease_lo);
      if (cur) {  // BTreeMap::insert over the old entry: a fresh KeyValue (create_revision too)
        cur->value = value, cur->create_revision = cur->modify_revision = revision;
      } else {
        ElKeyValue k;
        k.name = name, k.lease = lease_lo, k.value = value, k.create_revision = k.modify_revision = revision;
        kv.push_back(k);
      }
      publish(name);  // Event::Put(kv)
    }
    const ElKeyValue* ld = leader(name);
    if (ld->lease == lease_lo) {
      *won = true;
      r.flag = 1, r.revision = revision, r.data = revision;  // LeaderKey.rev = self.revision
    } else {
      *won = false;
      *rx = subscribe(name);
    }
    return r;
  }
  /// service.rs:540-554
  ElResponse proclaim(u32 name, u32 lease_lo, u32 value) {
    ElResponse r;
    ElKeyValue* k = find_kv(name, lease_lo);
    if (!k) {
      r.status = EL_SESSION_EXPIRED;
      return r;
    }
    revision += 1;
    k->value = value, k->modify_revision = revision;
    publish(name);
    r.revision = revision;
    return r;
  }
  ElResponse leader_rsp(u32 name) const {
    ElResponse r;
    r.revision = revision;
    if (const ElKeyValue* k = leader(name)) r.flag = 1, r.data = k->value | k->modify_revision << 16;
    return r;
  }
  /// service.rs:571-583
  ElResponse resign(u32 name, u32 lease_lo) {
    ElResponse r;
    ElKeyValue* k = find_kv(name, lease_lo);
    if (!k) {
      r.status = EL_SESSION_EXPIRED;
      return r;
    }
    ElLease* l = find_lease(k->lease);
    if (!l) panic("no lease", 68);
    for (size_t i = 0; i < l->keys.size(); i++)
      if (l->keys[i].first == name && l->keys[i].second == lease_lo) {
        l->keys.erase(l->keys.begin() + long(i));
        break;
      }
    remove_kv(name, lease_lo);
    publish(name);  // Event::Delete(kv)
    revision += 1;
    r.revision = revision;
    return r;
  }
  u32 count_prefix(u32 name) const {
    u32 n = 0;
    for (auto& k : kv) n += k.name == name;
    return n;
  }
};

struct ElService {
  float timeout_rate = 0.0f;
  std::shared_ptr<ElServiceInner> inner;
  u32 handled = 0;
  static Async<Unit> tick_task(std::weak_ptr<ElServiceInner> weak) {  // service.rs:27-33
    while (auto inner = weak.lock()) {
      inner->tick();
      inner.reset();
      co_await sleep(1 * SEC);
    }
    co_return {};
  }
  static std::shared_ptr<ElService> create(float timeout_rate) {
    auto s = std::make_shared<ElService>();
    s->timeout_rate = timeout_rate;
    s->inner = std::make_shared<ElServiceInner>();
    spawn(tick_task(std::weak_ptr<ElServiceInner>(s->inner)));
    return s;
  }
  Async<bool> timeout() {  // service.rs:165-176
    if (thread_rng().gen_bool(double(timeout_rate))) {
      u64 t = thread_rng().gen_range_duration(5 * SEC, 15 * SEC);
      co_await sleep(t);
      co_return true;
    }
    co_return false;
  }
  /// service.rs:94-121
  Async<ElResponse> campaign(ElRequest req) {
    ElResponse to;
    to.status = EL_TIMED_OUT;
    if (co_await timeout()) co_return to;
    bool won = false;
    ElEventRx rx;
    ElResponse r = inner->campaign(req.name, req.value, req.lease, &won, &rx);
    if (r.status != EL_OK || won) co_return r;
    for (;;) {
      if (!co_await rx.recv()) panic("sender should not drop", 81);
      const ElKeyValue* ld = inner->leader(req.name);
      if (!ld) {
        ElResponse e;
        e.status = EL_SESSION_EXPIRED;
        co_return e;
      }
      if (ld->lease == req.lease) {
        ElResponse ok;
        ok.flag = 1, ok.revision = inner->revision, ok.data = ld->modify_revision;  // rev: kv.modify_revision
        co_return ok;
      }
    }
  }
  /// every other request: timeout().await?, then the store
  Async<ElResponse> handle(ElRequest req) {
    ElResponse to;
    to.status = EL_TIMED_OUT;
    if (co_await timeout()) co_return to;
    switch (req.op) {
      case EL_GRANT: co_return inner->lease_grant(req.value);
      case EL_REVOKE: co_return inner->lease_revoke(req.lease);
      case EL_PROCLAIM: co_return inner->proclaim(req.name, req.lease, req.value);
      case EL_LEADER: co_return inner->leader_rsp(req.name);
      case EL_RESIGN: co_return inner->resign(req.name, req.lease);
      case EL_GET_PREFIX: {
        ElResponse r;
        r.revision = inner->revision, r.data = inner->count_prefix(req.name), r.flag = r.data != 0;
        co_return r;
      }
    }
    panic("unknown request", 64);
  }
};

/// select_biased!{ _ = tx.closed() => Closed, ret = fut => ret }
template <class T>
struct ElClosedOr {
  UnboundedSender<std::pair<Payload, LinkState>>::ClosedFut closed;
  Async<T> fut;
  Poll<std::optional<T>> poll(Context& cx) {  // nullopt = the client went away
    if (closed.poll(cx).ready()) return Poll<std::optional<T>>::done(std::nullopt);
    auto r = fut.poll(cx);
    if (!r.ready()) return Poll<std::optional<T>>::pending();
    return Poll<std::optional<T>>::done(std::move(*r.v));
  }
};
struct ElClosedOrEvent {
  UnboundedSender<std::pair<Payload, LinkState>>::ClosedFut closed;
  ElEventRx::RecvFut recv;
  Poll<int> poll(Context& cx) {  // 2 closed, 1 event, 0 the stream ended
    if (closed.poll(cx).ready()) return Poll<int>::done(2);
    auto r = recv.poll(cx);
    if (!r.ready()) return Poll<int>::pending();
    return Poll<int>::done(*r.v ? 1 : 0);
  }
};

/// server.rs:46-100: the task spawned per accepted connection
inline Async<Unit> el_request_task(Connection conn, std::shared_ptr<ElService> service) {
  auto r = co_await conn.rx.recv();
  if (r.is_err()) co_return {};
  ElRequest req = ElRequest::decode(*r.ok);
  service->handled++;
  ElResponse rsp;
  if (req.op == EL_CAMPAIGN) {  // server.rs:71-76
    auto ret = co_await ElClosedOr<ElResponse>{conn.tx.closed(), service->campaign(req)};
    if (!ret) co_return {};
    rsp = *ret;
  } else if (req.op == EL_OBSERVE) {  // server.rs:81-99
    if (co_await service->timeout()) {
      rsp.status = EL_TIMED_OUT;
    } else {
      ElEventRx stream = service->inner->subscribe(req.name);
      const ElKeyValue* k0 = service->inner->leader(req.name);
      std::optional<ElKeyValue> leader = k0 ? std::optional<ElKeyValue>(*k0) : std::nullopt;
      for (;;) {
        int k = co_await ElClosedOrEvent{conn.tx.closed(), stream.recv()};
        if (k != 1) co_return {};
        const ElKeyValue* n = service->inner->leader(req.name);
        std::optional<ElKeyValue> now = n ? std::optional<ElKeyValue>(*n) : std::nullopt;
        if (now == leader) continue;
        leader = now;
        if (conn.tx.send(service->inner->leader_rsp(req.name).encode()).is_err()) co_return {};
      }
    }
  } else {
    rsp = co_await service->handle(req);
  }
  (void)conn.tx.send(rsp.encode());
  co_return {};
}
inline Async<Unit> el_serve(SocketAddr addr, float timeout_rate, std::shared_ptr<ElService>* out) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  auto service = ElService::create(timeout_rate);
  *out = service;
  for (;;) {
    Connection conn = std::move((co_await ep.accept1()).unwrap());
    spawn(el_request_task(std::move(conn), service));
  }
}

/// one call = one connection (election.rs:40-42 etc.); `.unwrap()` on the transport
inline Async<ElResponse> el_call(const Endpoint& ep, SocketAddr server, ElRequest req) {
  Connection conn = std::move((co_await ep.connect1(server)).unwrap());
  if (conn.tx.send(req.encode()).is_err()) panic("election client: send failed", 65);
  auto r = co_await conn.rx.recv();
  if (r.is_err()) panic("election client: connection reset", 66);
  co_return ElResponse::decode(*r.ok);
}
inline ElResponse el_ok(ElResponse r, u32 code) {  // .unwrap() on the response
  if (r.status != EL_OK) panic("election: call failed", code);
  return r;
}

struct ElShared {
  u32 seen = 0, sum = 0;
};

/// tests/test.rs:154-188
inline Async<u32> el_first_leader(SocketAddr server, ElectionParams p) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  co_await sleep(5 * SEC);  // make sure the observer is ready
  ElRequest g;
  g.op = EL_GRANT, g.value = 60;
  u32 lease = el_ok(co_await el_call(ep, server, g), 82).data;
  ElRequest c;
  c.op = EL_CAMPAIGN, c.name = 0, c.value = 10, c.lease = lease;
  ElResponse r = el_ok(co_await el_call(ep, server, c), 83);
  if (!r.flag) panic("campaign returned no leader key", 83);
  ElRequest l;
  l.op = EL_LEADER, l.name = 0;
  r = el_ok(co_await el_call(ep, server, l), 84);
  if (!r.flag || (r.data & 0xffff) != 10) panic("leader value != 1", 84);
  el_ok(co_await el_call(ep, server, c), 85);  // campaign again completes immediately
  c.value = 11;
  el_ok(co_await el_call(ep, server, c), 85);  // campaign with a new value
  ElRequest pr;
  pr.op = EL_PROCLAIM, pr.name = 0, pr.lease = lease, pr.value = 12;
  el_ok(co_await el_call(ep, server, pr), 86);
  r = el_ok(co_await el_call(ep, server, l), 84);
  if (!r.flag || (r.data & 0xffff) != 12) panic("leader value != 1.2", 84);
  co_await sleep(p.hold_s * SEC);
  ElRequest rv;
  rv.op = EL_REVOKE, rv.lease = lease;
  el_ok(co_await el_call(ep, server, rv), 87);
  pr.value = 13;
  r = co_await el_call(ep, server, pr);  // proclaim should fail
  if (r.status == EL_OK) panic("proclaim after revoke succeeded", 88);
  ElRequest bad;
  bad.op = EL_CAMPAIGN, bad.name = 1, bad.value = 10, bad.lease = 1;  // campaign with invalid lease
  r = co_await el_call(ep, server, bad);
  if (r.status == EL_OK) panic("campaign with an invalid lease succeeded", 88);
  co_return lease;
}
/// tests/test.rs:191-207
inline Async<u32> el_second_leader(SocketAddr server) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  co_await sleep(10 * SEC);  // make sure client1 became leader
  ElRequest g;
  g.op = EL_GRANT, g.value = 60;
  u32 lease = el_ok(co_await el_call(ep, server, g), 82).data;
  ElRequest c;
  c.op = EL_CAMPAIGN, c.name = 0, c.value = 20, c.lease = lease;
  ElResponse r = el_ok(co_await el_call(ep, server, c), 83);
  if (!r.flag) panic("campaign returned no leader key", 83);
  ElRequest rs;
  rs.op = EL_RESIGN, rs.name = 0, rs.lease = lease;
  el_ok(co_await el_call(ep, server, rs), 89);
  co_return lease;
}
/// tests/test.rs:210-232
inline Async<u32> el_observer(SocketAddr server, std::shared_ptr<ElShared> sh) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  ElRequest o;
  o.op = EL_OBSERVE, o.name = 0;
  Connection stream = std::move((co_await ep.connect1(server)).unwrap());  // client.observe("leader")
  if (stream.tx.send(o.encode()).is_err()) panic("election client: send failed", 65);
  auto message = [&](u32 want) -> Async<Unit> {
    auto m = co_await stream.rx.recv();
    if (m.is_err()) panic("observe stream ended", 90);
    ElResponse r = el_ok(ElResponse::decode(*m.ok), 90);
    if (!r.flag || (r.data & 0xffff) != want) panic("observed leader value is wrong", 91);
    sh->seen++;
    sh->sum = (sh->sum ^ (r.data + (r.revision << 8))) * 0x9e3779b1u;
    co_return {};
  };
  co_await message(10);
  co_await message(11);
  co_await message(12);
  co_await sleep(15 * SEC);  // make sure client2 has campaigned
  ElRequest gp;
  gp.op = EL_GET_PREFIX, gp.name = 0;
  ElResponse r = el_ok(co_await el_call(ep, server, gp), 92);
  if (r.data != 2) panic("should have 2 keys with prefix leader", 92);
  co_await message(20);
  co_return sh->seen;
}

inline Async<u32> wl_election_main(ElectionParams p, std::array<u32, 4>* user) {
  auto& rt = current_runtime();
  SocketAddr server{ip4(10, 0, 0, 1), 2379};
  auto server_node = rt.create_node().ip(server.ip).build();
  auto client1 = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
  auto client2 = rt.create_node().ip(ip4(10, 0, 0, 3)).build();
  auto client3 = rt.create_node().ip(ip4(10, 0, 0, 4)).build();
  auto service = std::make_shared<std::shared_ptr<ElService>>();
  server_node.spawn(el_serve(server, p.timeout_rate, service.get()));
  co_await sleep(1 * SEC);
  auto sh = std::make_shared<ElShared>();
  auto t1 = client1.spawn(el_first_leader(server, p));
  auto t2 = client2.spawn(el_second_leader(server));
  auto t3 = client3.spawn(el_observer(server, sh));
  (co_await t1).unwrap();
  (co_await t2).unwrap();
  (co_await t3).unwrap();
  (*user)[0] = sh->seen;
  (*user)[1] = (*service)->inner->revision;
  (*user)[2] = sh->sum;
  (*user)[3] = (*service)->handled;
  co_return 0;
}

}  // namespace msim_oracle
