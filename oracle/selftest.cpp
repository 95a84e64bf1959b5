// ORACLE — TEST INFRASTRUCTURE ONLY.
// The reference's own unit tests for the hot path, re-expressed against the oracle runtime.
// Each case cites the test it restates; assertions are the reference's (virtual-time
// outcomes, set cardinalities) — the reference holds no golden vectors.
#include <cstdio>
#include <set>

#include "workloads.hpp"

using namespace msim_oracle;

static int g_failed = 0;
#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) {                                                    \
      std::printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      g_failed++;                                                  \
    }                                                              \
  } while (0)
#define ASSERT_PANIC(c) \
  do {                  \
    if (!(c)) panic(std::string("assertion failed: ") + #c); \
  } while (0)

template <class Fn>
static void run_case(const char* name, Fn fn) {
  int before = g_failed;
  try {
    fn();
  } catch (std::exception& e) {
    std::printf("  EXCEPTION in %s: %s\n", name, e.what());
    g_failed++;
  }
  std::printf("%s %s\n", g_failed == before ? "ok  " : "FAIL", name);
}

using Flag = std::shared_ptr<u32>;

static Async<Unit> tick_loop(Flag flag, u64 period) {
  for (;;) {
    co_await sleep(period);
    *flag += 2;
  }
}

// task/mod.rs:849-856
static Async<Unit> one(int) { co_return {}; }
static Async<Unit> spawn_in_block_on_main() {
  (co_await spawn(one(1))).unwrap();
  (co_await spawn(one(2))).unwrap();
  co_return {};
}

// task/mod.rs:858-900 kill
static Async<Unit> kill_main(Runtime* rt, u64 node1, Flag flag1, Flag flag2) {
  u64 t0 = now();
  co_await sleep_until(t0 + 3 * SEC);
  ASSERT_PANIC(*flag1 == 2);
  ASSERT_PANIC(*flag2 == 2);
  rt->kill(node1);
  rt->kill(node1);
  ASSERT_PANIC(rt->is_exit(node1));
  co_await sleep_until(t0 + 5 * SEC);
  ASSERT_PANIC(*flag1 == 2);
  ASSERT_PANIC(*flag2 == 4);
  co_return {};
}

// task/mod.rs:902-937 restart
static Async<Unit> restart_worker(Flag flag) {
  *flag = 0;
  for (;;) {
    co_await sleep(2 * SEC);
    *flag += 2;
  }
}
static Async<Unit> restart_main(Runtime* rt, u64 node, Flag flag) {
  u64 t0 = now();
  co_await sleep_until(t0 + 3 * SEC);
  ASSERT_PANIC(*flag == 2);
  rt->kill(node);
  rt->restart(node);
  ASSERT_PANIC(!rt->is_exit(node));
  co_await sleep_until(t0 + 6 * SEC);
  ASSERT_PANIC(*flag == 2);
  co_await sleep_until(t0 + 8 * SEC);
  ASSERT_PANIC(*flag == 4);
  co_return {};
}

// task/mod.rs:939-963 restart_on_panic
static Async<Unit> rop_worker(Flag flag) {
  u32 old = (*flag)++;
  if (old < 3) panic("explicit panic");
  co_return {};
}
static Async<Unit> rop_main(Flag flag) {
  co_await sleep(60 * SEC);
  ASSERT_PANIC(*flag == 4);
  co_return {};
}
// task/mod.rs:965-983 restart_on_panic_matching (should panic with "2")
static Async<Unit> ropm_worker(Flag flag) {
  u32 old = (*flag)++;
  panic(std::to_string(old));
  co_return {};
}
static Async<Unit> pending_main() {
  co_await PendingForever<Unit>{};
  co_return {};
}

// task/mod.rs:985-1016 pause_resume
static Async<Unit> pause_main(Runtime* rt, u64 node, Flag flag) {
  u64 t0 = now();
  co_await sleep_until(t0 + 3 * SEC);
  ASSERT_PANIC(*flag == 2);
  rt->pause(node);
  rt->pause(node);
  co_await sleep_until(t0 + 5 * SEC);
  ASSERT_PANIC(*flag == 2);
  rt->resume(node);
  rt->resume(node);
  co_await sleep_until(t0 + 5500 * MS);
  ASSERT_PANIC(*flag == 4);
  co_return {};
}

// task/mod.rs:1018-1041 random_select_from_ready_tasks
static Async<Unit> seq_task(std::shared_ptr<std::vector<int>> out, int i) {
  for (int j = 0; j < 5; j++) {
    out->push_back(i * 10 + j);
    co_await YieldNow{};
  }
  co_return {};
}
static Async<Unit> seq_main(std::shared_ptr<std::vector<int>> out) {
  std::vector<JoinHandle<Unit>> tasks;
  for (int i = 0; i < 3; i++) tasks.push_back(spawn(seq_task(out, i)));
  for (auto& t : tasks) (co_await t).unwrap();  // join_all polls in order; outcome identical here
  co_return {};
}

// task/mod.rs:1113-1139 kill_drop_futures
struct DropProbe {
  std::shared_ptr<int> counter;
  ~DropProbe() {
    if (counter) (*counter)++;
  }
  DropProbe(std::shared_ptr<int> c) : counter(std::move(c)) {}
  DropProbe(DropProbe&& o) noexcept : counter(std::move(o.counter)) {}
};
static Async<Unit> kdf_worker(DropProbe probe) {
  co_await PendingForever<Unit>{};
  (void)probe;
  co_return {};
}
static Async<Unit> kdf_main(Runtime* rt, u64 node, std::shared_ptr<int> dropped) {
  co_await sleep(1 * SEC);
  rt->kill(node);
  ASSERT_PANIC(*dropped == 0);  // strong_count == 2
  co_await sleep(1 * SEC);
  ASSERT_PANIC(*dropped == 1);  // future dropped after a while
  co_return {};
}

// task/mod.rs:1185-1252 spawn_in_future_drop_by_aborting_task / _by_killing_node: a value whose Drop spawns a task; the
// future that owns it is dropped by the executor (abort / kill), under the dropped task's own context (:262-289)
struct SpawnOnDrop {
  std::shared_ptr<u64> dropping_node;  // where the task spawned in drop() found itself (0 = it never ran)
  std::shared_ptr<bool> dropped;
  bool armed = true;
  SpawnOnDrop(std::shared_ptr<u64> n, std::shared_ptr<bool> d) : dropping_node(std::move(n)), dropped(std::move(d)) {}
  SpawnOnDrop(SpawnOnDrop&& o) noexcept : dropping_node(std::move(o.dropping_node)), dropped(std::move(o.dropped)), armed(o.armed) {
    o.armed = false;
  }
  static Async<Unit> note_node(std::shared_ptr<u64> out) {
    *out = current_task()->node->id;
    co_return {};
  }
  ~SpawnOnDrop() {
    if (!armed) return;
    spawn(note_node(dropping_node));
    *dropped = true;
  }
};
static Async<Unit> sod_holder(SpawnOnDrop a) {  // node.spawn(async move { drop(a) })
  (void)a;
  co_return {};
}
static Async<Unit> sod_abort_main(Runtime::NodeHandle node, std::shared_ptr<u64> dn, std::shared_ptr<bool> dropped) {
  auto h = node.spawn(sod_holder(SpawnOnDrop(dn, dropped)));  // not polled yet: single-threaded runtime
  h.abort();
  auto r = co_await h;
  ASSERT_PANIC(r.is_err() && r.err == Err::Cancelled);
  co_await sleep(114514 * SEC);  // let the task spawned in A::drop run
  ASSERT_PANIC(*dropped && *dn == node.id());  // ... under the node of the dropped task
  co_return {};
}
static Async<Unit> sod_kill_main(Runtime* rt, Runtime::NodeHandle node, std::shared_ptr<u64> dn, std::shared_ptr<bool> dropped) {
  auto h = node.spawn(sod_holder(SpawnOnDrop(dn, dropped)));
  rt->kill(node.id());
  auto r = co_await h;
  ASSERT_PANIC(r.is_err() && r.err == Err::Cancelled);
  ASSERT_PANIC(*dropped);
  co_await sleep(114514 * SEC);
  ASSERT_PANIC(*dn == 0);  // spawning on a killed node succeeds, but the task never runs
  co_return {};
}

// task/mod.rs:1141-1155 join_cancelled
static Async<Unit> jc_main(JoinHandle<Unit> h) {
  h.abort();
  auto r = co_await h;
  ASSERT_PANIC(r.is_err() && r.err == Err::Cancelled);
  co_return {};
}

// task/mod.rs:1157-1187 exited
static Async<Unit> exited_init(Flag flag) {
  spawn(tick_loop(flag, 2 * SEC));
  co_await sleep(5 * SEC);
  co_return {};  // exit here: the spawned task is killed
}
static Async<Unit> exited_main(Runtime* rt, u64 node, Flag flag) {
  ASSERT_PANIC(!rt->is_exit(node));
  co_await sleep(10 * SEC);
  ASSERT_PANIC(rt->is_exit(node));
  ASSERT_PANIC(*flag == 4);
  co_return {};
}

// time/mod.rs:263-304
static Async<Unit> time_main() {
  u64 t0 = now();
  co_await sleep(1 * SEC);
  ASSERT_PANIC(now() - t0 >= 1 * SEC);
  co_await sleep_until(t0 + 2 * SEC);
  ASSERT_PANIC(now() - t0 >= 2 * SEC);
  auto r = co_await timeout(2 * SEC, sleep(1 * SEC));
  ASSERT_PANIC(r.is_ok());
  auto r2 = co_await timeout(1 * SEC, sleep(2 * SEC));
  ASSERT_PANIC(r2.is_err());
  // advance
  u64 t1 = now();
  advance(1 * SEC);
  ASSERT_PANIC(now() - t1 == 1 * SEC);
  co_return {};
}

// net/endpoint.rs:372-408 send_recv (Barrier → both sides sleep until t=1s)
static Async<Unit> sr_node1(SocketAddr a1, SocketAddr a2) {
  Endpoint net = std::move((co_await Endpoint::bind(a1)).unwrap());
  co_await sleep_until(1 * SEC);
  Payload p1;
  p1.v[0] = 1;
  (co_await net.send_to(a2, 1, p1)).unwrap();
  co_await sleep(1 * SEC);
  Payload p2;
  p2.v[0] = 2;
  (co_await net.send_to(a2, 2, p2)).unwrap();
  co_return {};
}
static Async<Unit> sr_node2(SocketAddr a1, SocketAddr a2) {
  Endpoint net = std::move((co_await Endpoint::bind(a2)).unwrap());
  co_await sleep_until(1 * SEC);
  auto [d2, from2] = (co_await net.recv_from(2)).unwrap();
  ASSERT_PANIC(from2 == a1 && d2.v[0] == 2);
  auto [d1, from1] = (co_await net.recv_from(1)).unwrap();
  ASSERT_PANIC(from1 == a1 && d1.v[0] == 1);
  co_return {};
}
// net/endpoint.rs:410-443 receiver_drop
static Async<Unit> rd_node1(SocketAddr a1, SocketAddr a2) {
  Endpoint net = std::move((co_await Endpoint::bind(a1)).unwrap());
  co_await sleep_until(2 * SEC);
  Payload p1;
  p1.v[0] = 1;
  (co_await net.send_to(a2, 1, p1)).unwrap();
  co_return {};
}
static Async<Unit> rd_node2(SocketAddr a1, SocketAddr a2) {
  Endpoint net = std::move((co_await Endpoint::bind(a2)).unwrap());
  auto r = co_await timeout(1 * SEC, net.recv_from(1));
  ASSERT_PANIC(r.is_err());  // timeout and receiver dropped here
  co_await sleep_until(2 * SEC);
  auto [d, from] = (co_await net.recv_from(1)).unwrap();  // receive again should succeed
  ASSERT_PANIC(from == a1 && d.v[0] == 1);
  co_return {};
}
// net/endpoint.rs:469-512 bind
static Async<Unit> bind_case() {
  u32 ip = ip4(10, 0, 0, 1);
  {
    Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{IP_UNSPECIFIED, 0})).unwrap());
    ASSERT_PANIC(ep.local_addr().ip == IP_UNSPECIFIED && ep.local_addr().port != 0);
  }
  {
    Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{IP_LOCALHOST, 0})).unwrap());
    ASSERT_PANIC(ep.local_addr().ip == IP_LOCALHOST && ep.local_addr().port != 0);
  }
  auto bad = co_await Endpoint::bind(SocketAddr{ip4(10, 0, 0, 2), 0});
  ASSERT_PANIC(bad.is_err() && bad.err == Err::AddrNotAvailable);
  {
    Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{ip, 100})).unwrap());
    ASSERT_PANIC(ep.local_addr() == (SocketAddr{ip, 100}));
    auto dup = co_await Endpoint::bind(SocketAddr{ip, 100});
    ASSERT_PANIC(dup.is_err() && dup.err == Err::AddrInUse);
  }  // drop and reuse port
  (co_await Endpoint::bind(SocketAddr{ip, 100})).unwrap();
  co_return {};
}
// net/endpoint.rs:514-548 localhost
static Async<Unit> lh_node1() {
  Endpoint ep1 = std::move((co_await Endpoint::bind(SocketAddr{IP_LOCALHOST, 1})).unwrap());
  Endpoint ep2 = std::move((co_await Endpoint::bind(SocketAddr{ip4(10, 0, 0, 1), 2})).unwrap());
  co_await sleep_until(1 * SEC);
  auto r = co_await timeout(1 * SEC, ep1.recv_from(1));
  ASSERT_PANIC(r.is_err());  // localhost endpoint should not receive from other nodes
  auto [d, from] = (co_await ep2.recv_from(1)).unwrap();
  ASSERT_PANIC(from == (SocketAddr{ip4(10, 0, 0, 2), 1}));
  co_return {};
}
static Async<Unit> lh_node2() {
  Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{IP_LOCALHOST, 1})).unwrap());
  co_await sleep_until(1 * SEC);
  Payload p;
  (co_await ep.send_to(SocketAddr{ip4(10, 0, 0, 1), 1}, 1, p)).unwrap();
  (co_await ep.send_to(SocketAddr{ip4(10, 0, 0, 1), 2}, 1, p)).unwrap();
  co_return {};
}
// net/rpc.rs doc example :15-61 — call / add_rpc_handler round trip
static Async<Payload> echo_plus_one(Payload q) {
  Payload r;
  r.v[0] = q.v[0] + 1;
  co_return r;
}
static Async<Unit> rpc_server(SocketAddr a) {
  Endpoint ep = std::move((co_await Endpoint::bind(a)).unwrap());
  ep.add_rpc_handler(77, [](Payload q) { return echo_plus_one(q); });
  co_await PendingForever<Unit>{};
  co_return {};
}
static Async<Unit> rpc_client(SocketAddr server) {
  co_await sleep(1 * SEC);
  Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{IP_UNSPECIFIED, 0})).unwrap());
  for (u32 i = 0; i < 5; i++) {
    Payload q;
    q.v[0] = i;
    Payload r = (co_await ep.call(server, 77, q)).unwrap();
    ASSERT_PANIC(r.v[0] == i + 1);
  }
  // a call to nobody times out
  Payload q;
  auto r = co_await ep.call_timeout(SocketAddr{ip4(10, 0, 0, 9), 1}, 77, q, 1 * SEC);
  ASSERT_PANIC(r.is_err() && r.err == Err::TimedOut);
  co_return {};
}

template <class T>
static Async<Unit> join_unwrap(JoinHandle<T> h) {
  Result<T> r = co_await h;
  r.unwrap();
  co_return {};
}


// ---- madsim-etcd-client/tests/test.rs:9-62 kv(), :64-128 lease() on the restated SimServer ----
static Async<Unit> etcd_kv_client(SocketAddr server) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  auto call = [&](u32 op, u32 key, u32 value, u32 lease = 0, u32 ttl = 0) {
    EtcdRequest q;
    q.op = op, q.key = key, q.value = value, q.lease = lease, q.ttl = ttl;
    return etcd_call(ep, server, q);
  };
  EtcdResponse r = co_await call(ETCD_PUT, 0, 0xba5);  // put foo bar
  CHECK(r.status == ETCD_OK);
  r = co_await call(ETCD_GET, 0, 0);
  u32 revision = r.revision;
  CHECK(r.flag == 1 && r.value == 0xba5 && r.mod_revision == revision);  // create == mod == header revision
  r = co_await call(ETCD_PUT, 0, 0x99);  // put again
  r = co_await call(ETCD_GET, 0, 0);
  CHECK(r.value == 0x99 && r.mod_revision == r.revision && r.revision == revision + 1);
  r = co_await call(ETCD_DELETE, 0, 0);
  CHECK(r.flag == 1);
  r = co_await call(ETCD_GET, 0, 0);
  CHECK(r.flag == 0);
  co_return {};
}
static Async<Unit> etcd_lease_client(SocketAddr server) {
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  auto call = [&](u32 op, u32 key, u32 value, u32 lease = 0, u32 ttl = 0) {
    EtcdRequest q;
    q.op = op, q.key = key, q.value = value, q.lease = lease, q.ttl = ttl;
    return etcd_call(ep, server, q);
  };
  EtcdResponse r = co_await call(ETCD_GRANT, 0, 0, 9, 60);  // grant(60)
  CHECK(r.status == ETCD_OK);
  r = co_await call(ETCD_PUT, 1, 7, 9);  // put with the lease
  CHECK(r.status == ETCD_OK);
  r = co_await call(ETCD_GET, 1, 0);
  CHECK(r.flag == 1);
  r = co_await call(ETCD_LEASE_READ, 1, 0);  // leases(): exactly one
  CHECK(r.status == ETCD_OK && r.value == 1);
  // keep alive for 90 s: the first response is produced when the keep-alive connection opens,
  // the second when the keeper pings again
  co_await sleep(45 * SEC);
  {
    auto cn = co_await ep.connect1(server);
    Connection conn = std::move(cn.unwrap());
    EtcdRequest ka;
    ka.op = ETCD_KEEP_ALIVE, ka.lease = 9;
    conn.tx.send(ka.encode()).unwrap();
    EtcdResponse k1 = EtcdResponse::decode((co_await conn.rx.recv()).unwrap());
    CHECK(k1.status == ETCD_OK && k1.value == 60);
    co_await sleep(45 * SEC);
    conn.tx.send(ka.encode()).unwrap();
    EtcdResponse k2 = EtcdResponse::decode((co_await conn.rx.recv()).unwrap());
    CHECK(k2.status == ETCD_OK && k2.revision == k1.revision + 1);
  }
  r = co_await call(ETCD_LEASE_READ, 0, 0, 9);  // time_to_live: just refreshed
  CHECK(r.status == ETCD_OK && r.value <= 60 && r.value > 50);
  r = co_await call(ETCD_GET, 1, 0);
  CHECK(r.flag == 1);
  co_await sleep(61 * SEC);  // wait for the lease to expire
  r = co_await call(ETCD_GET, 1, 0);
  CHECK(r.flag == 0);
  r = co_await call(ETCD_PUT, 1, 7, 1);  // put with an invalid lease
  CHECK(r.status == ETCD_LEASE_NOT_FOUND);
  r = co_await call(ETCD_REVOKE, 0, 0, 1);  // revoke an invalid lease
  CHECK(r.status == ETCD_LEASE_NOT_FOUND);
  r = co_await call(ETCD_LEASE_READ, 0, 0, 1);  // time_to_live of an invalid lease
  CHECK(r.status == ETCD_LEASE_NOT_FOUND);
  r = co_await call(ETCD_GRANT, 0, 0, 9, 60);  // a revoked lease takes its keys with it
  r = co_await call(ETCD_PUT, 2, 5, 9);
  r = co_await call(ETCD_REVOKE, 0, 0, 9);
  CHECK(r.status == ETCD_OK);
  r = co_await call(ETCD_GET, 2, 0);
  CHECK(r.flag == 0);
  co_return {};
}
template <class F>
static void etcd_case(F client) {
  Runtime rt;
  SocketAddr a{ip4(10, 0, 0, 1), 2379};
  auto server = rt.create_node().ip(a.ip).build();
  auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
  server.spawn(etcd_serve(a, 0.0f, 0));
  auto h = cl.spawn(client(a));
  rt.block_on(join_unwrap(std::move(h)));
  rt.teardown();
}

// ---- tonic-example/tests/test.rs:280-315 unimplemented_service(), :317-363 interceptor() ----
static Async<Unit> tonic_unimplemented_client(SocketAddr server) {
  co_await sleep(1 * SEC);
  (co_await tonic_connect_ep(server)).unwrap();  // GreeterClient::connect
  Payload q;
  q.v[0] = TONIC_UNKNOWN_SERVICE;  // the server only has AnotherGreeter: "/helloworld.Greeter/SayHello" is nobody's
  auto r = co_await tonic_unary(server, q);
  CHECK(r.is_ok() && r.ok->v[0] == TONIC_UNIMPLEMENTED);  // error.code() == Code::Unimplemented
  r = co_await tonic_unary(server, q);
  CHECK(r.is_ok() && r.ok->v[0] == TONIC_UNIMPLEMENTED);
  co_return {};
}
static Async<Unit> tonic_interceptor_client(SocketAddr server) {
  co_await sleep(1 * SEC);
  (co_await tonic_connect_ep(server)).unwrap();
  u32 i = 0;  // GreeterClient::with_interceptor(channel, move |req| { i += 1; if i % 2 == 0 { Err(unavailable) } else { Ok(req) } })
  auto say_hello = [&]() -> Async<u32> {
    if (++i % 2 == 0) co_return u32(TONIC_UNAVAILABLE);
    Payload q;
    q.v[0] = TONIC_NAME;
    auto r = co_await tonic_unary(server, q);
    CHECK(r.is_ok());
    co_return r.ok->v[0];
  };
  CHECK(co_await say_hello() == TONIC_OK);           // (1, 1)
  CHECK(co_await say_hello() == TONIC_UNAVAILABLE);  // (2, 1)
  CHECK(co_await say_hello() == TONIC_UNAVAILABLE);  // (3, 2)
  CHECK(co_await say_hello() == TONIC_UNAVAILABLE);  // (4, 2)
  CHECK(co_await say_hello() == TONIC_OK);           // (5, 3)
  co_return {};
}

// ---- tonic-example/tests/test.rs:22-70 basic() (unary part), :139-151 invalid_address() ----
static Async<Unit> tonic_basic_client(SocketAddr server) {
  co_await sleep(1 * SEC);
  auto ep = co_await tonic_connect_ep(server);  // GreeterClient::connect
  CHECK(ep.is_ok());
  ep.ok.reset();
  Payload q;
  q.v[0] = TONIC_NAME;
  auto r = co_await tonic_unary(server, q);
  CHECK(r.is_ok() && r.ok->v[0] == TONIC_OK);
  CHECK(r.ok->v[1] == tonic_reply_for(TONIC_NAME, SocketAddr{ip4(10, 0, 0, 2), 0}));  // "Hello Tonic! (10.0.0.2)"
  q.v[0] = TONIC_ERROR;
  r = co_await tonic_unary(server, q);
  CHECK(r.is_ok() && r.ok->v[0] == TONIC_INVALID_ARGUMENT);  // Status::invalid_argument("error!")
  co_return {};
}
static Async<Unit> tonic_invalid_address_client(SocketAddr server) {
  auto ep = co_await tonic_connect_ep(server);  // nobody listens on 10.0.0.1:50051
  CHECK(ep.is_err());
  co_return {};
}

// ---- signal.rs:23-66 ctrl_c_kill / ctrl_c_catch; interval.rs:73-160 ----
static Async<Unit> pending_task() {
  co_await PendingForever<Unit>{};
  co_return {};
}
static Async<Unit> ctrl_c_kill_main(Runtime* rt, Runtime::NodeHandle node) {
  auto h = node.spawn(pending_task());
  CHECK(!h.is_finished());
  rt->send_ctrl_c(node.id());  // no handler was ever installed: ctrl-c kills the node
  co_await sleep(1 * SEC);
  CHECK(h.is_finished());
  co_return {};
}
static Async<Unit> ctrl_c_catch_task(Flag flag) {
  co_await ctrl_c();
  *flag = 1;
  co_await PendingForever<Unit>{};
  co_return {};
}
static Async<Unit> ctrl_c_catch_main(Runtime* rt, Runtime::NodeHandle node) {
  for (int i = 0; i < 2; i++) {
    Flag flag = std::make_shared<u32>(0);
    auto h = node.spawn(ctrl_c_catch_task(flag));
    co_await sleep(1 * SEC);
    CHECK(*flag == 0 && !h.is_finished());
    rt->send_ctrl_c(node.id());  // caught: the node survives
    co_await sleep(1 * SEC);
    CHECK(*flag == 1 && !h.is_finished());
  }
  co_return {};
}
static Async<Unit> interval_main(MissedTickBehavior b, std::shared_ptr<std::vector<u64>> out) {
  u64 t0 = now();
  Interval iv = interval(100 * MS);
  iv.set_missed_tick_behavior(b);
  out->push_back(co_await iv.tick() - t0);  // the first tick: sleep_until(now) = now + 1 ms
  co_await sleep(350 * MS);                 // miss three ticks
  for (int i = 0; i < 4; i++) out->push_back(co_await iv.tick() - t0);
  co_return {};
}

// ---- tonic-example/tests/test.rs:72-83 basic() server stream; :204-230 client_drops_response_stream ----
static Async<Unit> tonic_server_stream_client(SocketAddr server, std::shared_ptr<TonicStats> st) {
  co_await sleep(1 * SEC);
  Payload q;
  q.v[0] = TONIC_NAME;
  u64 t0 = now();
  auto r = co_await tonic_server_streaming(server, q, false, tonic_reply_for(TONIC_NAME, SocketAddr{ip4(10, 0, 0, 2), 0}));
  CHECK(r.is_ok() && *r.ok == 3);             // "{i}: Hello Tonic! (10.0.0.2)" x3, then Status::unknown("EOF")
  CHECK(now() - t0 > 3 * SEC);                // the generator sleeps 1 s after every item
  // drop the response stream at once: the server's select sees tx.closed() and stops streaming
  r = co_await tonic_server_streaming(server, q, true, 0);
  CHECK(r.is_ok() && *r.ok == 0);
  co_await sleep(10 * SEC);
  CHECK(st->handled == 2);
  // client stream (tests/test.rs:85-94): "Hello Tonic0 Tonic1 Tonic2! (10.0.0.2)" after the three requests, 1 s apart
  t0 = now();
  r = co_await tonic_client_streaming(server, 0, SocketAddr{ip4(10, 0, 0, 2), 0});
  CHECK(r.is_ok() && *r.ok == 3 && now() - t0 > 3 * SEC);
  CHECK(st->handled == 3);
  // bi-directional stream (tests/test.rs:96-107): "Hello Tonic{i}! (10.0.0.2)" x3, then the trailer
  t0 = now();
  r = co_await tonic_bidi(server, 0, SocketAddr{ip4(10, 0, 0, 2), 0}, false);
  CHECK(r.is_ok() && *r.ok == 3 && now() - t0 > 3 * SEC);
  // dropping the response stream after one reply cancels the background sender (codec.rs:29, cancel_on_drop)
  t0 = now();
  r = co_await tonic_bidi(server, 1, SocketAddr{ip4(10, 0, 0, 2), 0}, true);
  CHECK(r.is_ok() && *r.ok == 1 && now() - t0 < 1 * SEC);
  co_await sleep(5 * SEC);
  CHECK(st->handled == 5);
  co_return {};
}

// ---- tonic-example/tests/test.rs:233-278 server_crash ----
static Async<Unit> tonic_server_crash_client(Runtime* rt, u64 server_node, SocketAddr server) {
  SocketAddr me{ip4(10, 0, 0, 2), 0};
  Payload q;
  q.v[0] = TONIC_NAME;
  auto r = co_await tonic_unary(server, q);
  CHECK(r.is_ok());
  // a bidi call is in flight when the server dies: the stream ends in "broken pipe", not in a trailer
  auto ep = co_await tonic_connect_ep(server);
  auto c = co_await ep.unwrap().connect1(server);
  Connection conn = std::move(c.unwrap());
  auto task = spawn(tonic_bidi_sender(std::move(conn.tx), 0));
  auto header = co_await conn.rx.recv();
  CHECK(header.is_ok() && header.ok->v[0] == TONIC_OK);
  co_await sleep(1 * SEC);
  rt->kill(server_node);
  co_await sleep(1 * SEC);
  bool broken = false;
  for (;;) {
    auto m = co_await conn.rx.recv();
    if (m.is_err()) {
      broken = true;
      break;
    }
    CHECK(m.ok->v[0] == TONIC_OK);  // replies that were already on their way
  }
  CHECK(broken);
  r = co_await tonic_unary(server, q);  // Code::Unavailable: nobody listens any more
  CHECK(r.is_err());
  (void)task;
  (void)me;
  co_return {};
}

// ---- tonic-example/tests/test.rs:369-410 request_timeout ----
static Async<Unit> tonic_timeout_client(SocketAddr server) {
  co_await sleep(1 * SEC);
  Payload q;
  q.v[0] = TONIC_DELAY;  // AnotherGreeter::delay answers after 10 s
  u64 t0 = now();
  {
    auto t = timeout(1 * SEC, tonic_unary(server, q));  // Endpoint::timeout(1 s)
    auto r = co_await t;
    CHECK(r.is_err());  // Code::DeadlineExceeded
  }
  CHECK(now() - t0 >= 1 * SEC && now() - t0 < 1 * SEC + 10 * MS);
  t0 = now();
  {
    auto t = timeout(5 * SEC, tonic_unary(server, q));  // Request::set_timeout(5 s) overrides the channel's
    auto r = co_await t;
    CHECK(r.is_err());
  }
  CHECK(now() - t0 >= 5 * SEC && now() - t0 < 5 * SEC + 10 * MS);
  q.v[0] = TONIC_NAME;
  {
    auto t = timeout(1 * SEC, tonic_unary(server, q));
    auto r = co_await t;
    CHECK(r.is_ok() && r.ok->is_ok());  // a prompt call is untouched by the timeout
  }
  co_return {};
}

int main() {
  run_case("task::spawn_in_block_on", [] {
    Runtime rt;
    rt.block_on(spawn_in_block_on_main());
  });
  run_case("task::kill", [] {
    Runtime rt;
    auto n1 = rt.create_node().build();
    auto n2 = rt.create_node().build();
    Flag f1 = std::make_shared<u32>(0), f2 = std::make_shared<u32>(0);
    n1.spawn(tick_loop(f1, 2 * SEC));
    n2.spawn(tick_loop(f2, 2 * SEC));
    rt.block_on(kill_main(&rt, n1.id(), f1, f2));
  });
  run_case("task::restart", [] {
    Runtime rt;
    Flag flag = std::make_shared<u32>(0);
    auto n = rt.create_node().init([flag] { return restart_worker(flag); }).build();
    rt.block_on(restart_main(&rt, n.id(), flag));
  });
  run_case("task::restart_on_panic", [] {
    Runtime rt;
    Flag flag = std::make_shared<u32>(0);
    rt.create_node().init([flag] { return rop_worker(flag); }).restart_on_panic().build();
    rt.block_on(rop_main(flag));
    CHECK(*flag == 4);
  });
  run_case("task::restart_on_panic_matching", [] {
    Runtime rt;
    Flag flag = std::make_shared<u32>(0);
    rt.create_node()
        .init([flag] { return ropm_worker(flag); })
        .restart_on_panic_matching("0")
        .restart_on_panic_matching("1")
        .build();
    bool panicked_with_2 = false;
    try {
      rt.block_on(pending_main());
    } catch (Panic& e) {
      panicked_with_2 = std::string(e.what()) == "2";
    }
    CHECK(panicked_with_2);
  });
  run_case("task::pause_resume", [] {
    Runtime rt;
    auto n = rt.create_node().build();
    Flag flag = std::make_shared<u32>(0);
    n.spawn(tick_loop(flag, 2 * SEC));
    rt.block_on(pause_main(&rt, n.id(), flag));
  });
  run_case("task::random_select_from_ready_tasks", [] {
    std::set<std::vector<int>> seqs;
    for (u64 seed = 0; seed < 10; seed++) {
      Runtime rt(seed);
      auto out = std::make_shared<std::vector<int>>();
      rt.block_on(seq_main(out));
      CHECK(out->size() == 15);
      seqs.insert(*out);
    }
    CHECK(seqs.size() == 10);
  });
  run_case("task::kill_drop_futures", [] {
    Runtime rt;
    auto n = rt.create_node().build();
    auto dropped = std::make_shared<int>(0);
    n.spawn(kdf_worker(DropProbe(dropped)));
    rt.block_on(kdf_main(&rt, n.id(), dropped));
  });
  run_case("task::spawn_in_future_drop_by_aborting_task (task/mod.rs:1185-1217)", [] {
    Runtime rt;
    auto n = rt.create_node().build();
    rt.block_on(sod_abort_main(n, std::make_shared<u64>(0), std::make_shared<bool>(false)));
  });
  run_case("task::spawn_in_future_drop_by_killing_node (task/mod.rs:1219-1252)", [] {
    Runtime rt;
    auto n = rt.create_node().build();
    rt.block_on(sod_kill_main(&rt, n, std::make_shared<u64>(0), std::make_shared<bool>(false)));
  });
  run_case("buggify::buggify (buggify.rs:39-63)", [] {
    Runtime rt;
    GlobalRng& g = rt.rand;
    CHECK(!g.is_buggify_enabled());
    CHECK(!g.buggify());  // disabled: false WITHOUT a draw
    u64 d0 = g.draws();
    g.enable_buggify();
    CHECK(g.is_buggify_enabled() && g.draws() == d0);
    int n25 = 0, n10 = 0;
    for (int i = 0; i < 1000; i++) n25 += g.buggify();
    for (int i = 0; i < 1000; i++) n10 += g.buggify_with_prob(0.1);
    CHECK(n25 >= 200 && n25 < 300);  // 25 %
    CHECK(n10 >= 50 && n10 < 150);   // 10 %
    CHECK(g.draws() == d0 + 2000);   // one draw per call while enabled
    g.disable_buggify();
    CHECK(!g.is_buggify_enabled() && !g.buggify() && g.draws() == d0 + 2000);
  });
  run_case("task::join_cancelled", [] {
    Runtime rt;
    auto n = rt.create_node().build();
    auto h = n.spawn(pending_main());
    rt.block_on(jc_main(h));
  });
  run_case("task::exited", [] {
    Runtime rt;
    Flag flag = std::make_shared<u32>(0);
    auto n = rt.create_node().init([flag] { return exited_init(flag); }).build();
    rt.block_on(exited_main(&rt, n.id(), flag));
  });
  run_case("task::deadlock_panics (runtime/mod.rs:118-125)", [] {
    Runtime rt;
    bool dl = false;
    try {
      rt.block_on(pending_main());
    } catch (Deadlock&) {
      dl = true;
    }
    CHECK(dl);
  });
  run_case("time::sleep/timeout/advance", [] {
    Runtime rt;
    rt.block_on(time_main());
  });
  run_case("time::time_limit (task/mod.rs:253-258)", [] {
    Runtime rt;
    rt.set_time_limit(5 * SEC);
    bool tl = false;
    try {
      rt.block_on(wl_sleep_main(10 * SEC));
    } catch (TimeLimitExceeded&) {
      tl = true;
    }
    CHECK(tl);
  });
  run_case("rand::deterministic per seed, distinct across seeds (rand.rs:312-354)", [] {
    std::set<u64> firsts;
    for (u64 seed = 0; seed < 3; seed++)
      for (int rep = 0; rep < 3; rep++) {
        Runtime rt(seed);
        firsts.insert(rt.rand.gen_u64());
      }
    CHECK(firsts.size() == 3);
  });
  run_case("time::system_time base ~2022 (time/mod.rs:26-33)", [] {
    Runtime rt(0);
    u64 y2022 = 60ull * 60 * 24 * 365 * 52;
    CHECK(rt.task.time.base_time_secs() >= y2022 && rt.task.time.base_time_secs() < y2022 + 31536000ull);
    CHECK(rt.task.time.base_time_secs() - y2022 == 10235805ull);  // SURVEY Appendix B [DERIVED]
  });
  auto two_nodes = [](auto f1, auto f2, bool join_first = false) {
    Runtime rt;
    SocketAddr a1{ip4(10, 0, 0, 1), 1}, a2{ip4(10, 0, 0, 2), 1};
    auto n1 = rt.create_node().ip(a1.ip).build();
    auto n2 = rt.create_node().ip(a2.ip).build();
    auto h1 = n1.spawn(f1(a1, a2));
    auto h2 = n2.spawn(f2(a1, a2));
    if (join_first) rt.block_on(join_unwrap(h1));
    rt.block_on(join_unwrap(h2));
  };
  run_case("net::send_recv", [&] { two_nodes(sr_node1, sr_node2); });
  run_case("net::receiver_drop", [&] { two_nodes(rd_node1, rd_node2); });
  run_case("net::bind", [] {
    Runtime rt;
    auto n = rt.create_node().ip(ip4(10, 0, 0, 1)).build();
    auto h = n.spawn(bind_case());
    rt.block_on(join_unwrap(h));
  });
  run_case("net::localhost", [&] {
    two_nodes([](SocketAddr, SocketAddr) { return lh_node1(); }, [](SocketAddr, SocketAddr) { return lh_node2(); },
              true);
  });
  run_case("net::connect_send_recv == pingpong(1 round)", [] {
    Runtime rt;
    u32 done = rt.block_on(wl_pingpong_main(1, 10 * MS));
    CHECK(done == 1);
  });
  run_case("rpc::call / call_timeout / add_rpc_handler", [] {
    Runtime rt;
    SocketAddr a1{ip4(10, 0, 0, 1), 1};
    auto n1 = rt.create_node().ip(a1.ip).build();
    auto n2 = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    n1.spawn(rpc_server(a1));
    auto h = n2.spawn(rpc_client(a1));
    rt.block_on(join_unwrap(h));
  });
  run_case("check_determinism log (runtime/mod.rs:178-202)", [] {
    std::vector<u8> log;
    {
      Runtime rt(7);
      rt.rand.enable_log();
      rt.block_on(wl_pingpong_main(3, 10 * MS));
      log = *rt.rand.take_log();
    }
    CHECK(!log.empty());
    {
      Runtime rt(7);
      rt.rand.enable_check(log);
      rt.block_on(wl_pingpong_main(3, 10 * MS));  // must not panic
    }
    bool detected = false;
    try {
      Runtime rt(8);
      rt.rand.enable_check(log);
      rt.block_on(wl_pingpong_main(3, 10 * MS));
    } catch (Panic&) {
      detected = true;
    }
    CHECK(detected);
  });
  run_case("etcd::kv (madsim-etcd-client/tests/test.rs:9-62)", [] { etcd_case(etcd_kv_client); });
  run_case("etcd::lease grant / leases / keep_alive / time_to_live / expiry / revoke / not found (tests/test.rs:64-128)", [] { etcd_case(etcd_lease_client); });
  run_case("etcd::election campaign / proclaim / leader / observe / resign (madsim-etcd-client/tests/test.rs:129-232)", [] {
    // the program IS the reference test (oracle/workloads_election.hpp): its assertions are panics; several seeds, and
    // shorter holds, must all end with the observer having seen 1, 1.1, 1.2, 2 and the store at revision 9
    for (u64 seed : {0ull, 1ull, 7ull, 123456789ull}) {
      for (u64 hold : {30ull, 20ull}) {
        Runtime rt(seed, Config{});
        std::array<u32, 4> user{};
        ElectionParams p;
        p.hold_s = hold;
        rt.block_on(wl_election_main(p, &user));
        CHECK(user[0] == 4 && user[1] == 9 && user[3] == 15);
        rt.teardown();
      }
    }
    // a server that times requests out makes some `unwrap()` of the test panic — as it would in the reference
    int panics = 0;
    for (u64 seed = 0; seed < 20; seed++) {
      Runtime rt(seed, Config{});
      std::array<u32, 4> user{};
      ElectionParams p;
      p.timeout_rate = 0.2f;
      try {
        rt.block_on(wl_election_main(p, &user));
      } catch (Panic&) {
        panics++;
      }
      rt.teardown();
    }
    CHECK(panics > 0);
  });
  run_case("tonic::basic unary + invalid_argument (tonic-example/tests/test.rs:22-70)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st));
    auto h = cl.spawn(tonic_basic_client(a));
    rt.block_on(join_unwrap(std::move(h)));
    CHECK(st->handled == 2);
    rt.teardown();
  });
  run_case("tonic::basic server / client / bidi streams + client_drops_response_stream (tests/test.rs:72-107,204-230)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st));
    auto h = cl.spawn(tonic_server_stream_client(a, st));
    rt.block_on(join_unwrap(std::move(h)));
    rt.teardown();
  });
  run_case("tonic::server_crash (tests/test.rs:233-278)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st));
    auto h = cl.spawn(tonic_server_crash_client(&rt, server.id(), a));
    rt.block_on(join_unwrap(std::move(h)));
    rt.teardown();
  });
  run_case("tonic::request_timeout (tests/test.rs:369-410)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st));
    auto h = cl.spawn(tonic_timeout_client(a));
    rt.block_on(join_unwrap(std::move(h)));
    rt.teardown();
  });
  run_case("tonic::unimplemented_service (tests/test.rs:280-315)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st, true));
    auto h = cl.spawn(tonic_unimplemented_client(a));
    rt.block_on(join_unwrap(std::move(h)));
    CHECK(st->handled == 0 && st->server_calls == 0);  // no service was called, no interceptor ran
    rt.teardown();
  });
  run_case("tonic::interceptor on both sides (tests/test.rs:317-363)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    auto server = rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto st = std::make_shared<TonicStats>();
    server.spawn(tonic_serve(a, st, true));
    auto h = cl.spawn(tonic_interceptor_client(a));
    rt.block_on(join_unwrap(std::move(h)));
    CHECK(st->server_calls == 3 && st->handled == 2);
    rt.teardown();
  });
  run_case("tonic::invalid_address (tests/test.rs:139-151)", [] {
    Runtime rt;
    SocketAddr a{ip4(10, 0, 0, 1), 50051};
    rt.create_node().ip(a.ip).build();
    auto cl = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
    auto h = cl.spawn(tonic_invalid_address_client(a));
    rt.block_on(join_unwrap(std::move(h)));
  });
  run_case("tonic::client_crash (tests/test.rs:155-202): 10 random restarts of the client node", [] {
    for (u64 seed = 0; seed < 8; seed++) {
      Runtime rt(seed);
      TonicParams p;
      p.n_clients = 1, p.flags = 1;
      std::array<u32, 4> user{};
      rt.block_on(wl_tonic_main(p, &user));
      CHECK(user[1] == 11 && user[0] > 0);  // the init task ran once per restart + once at build
      rt.teardown();
    }
  });
  run_case("signal::ctrl_c_kill (signal.rs:23-40)", [] {
    Runtime rt;
    auto node = rt.create_node().build();
    rt.block_on(ctrl_c_kill_main(&rt, node));
  });
  run_case("signal::ctrl_c_catch (signal.rs:42-66)", [] {
    Runtime rt;
    auto node = rt.create_node().build();
    rt.block_on(ctrl_c_catch_main(&rt, node));
  });
  run_case("time::interval MissedTickBehavior Burst / Delay / Skip (interval.rs:73-160)", [] {
    auto run = [](MissedTickBehavior b) {
      Runtime rt;
      auto out = std::make_shared<std::vector<u64>>();
      rt.block_on(interval_main(b, out));
      for (auto& v : *out) v /= MS;  // whole ms: the 50-100 ns poll jitter only matters for Delay's `now`
      return *out;
    };
    // scheduled instants: 1 ms, then after ~351 ms of work three ticks (101, 201, 301) are overdue
    CHECK((run(MissedTickBehavior::Burst) == std::vector<u64>{1, 101, 201, 301, 401}));
    CHECK((run(MissedTickBehavior::Delay) == std::vector<u64>{1, 101, 451, 551, 651}));
    CHECK((run(MissedTickBehavior::Skip) == std::vector<u64>{1, 101, 401, 501, 601}));
  });
  run_case("workload::raft elects exactly one leader", [] {
    for (u64 seed = 0; seed < 20; seed++) {
      Runtime rt(seed);
      auto c = std::make_shared<RaftCluster>();
      RaftOutcome out;
      rt.block_on(wl_raft_main(c, &out));
      CHECK(out.leader >= 1 && out.leader <= 5);
      CHECK(out.term >= 1);
      rt.teardown();
    }
  });
  std::printf("%s (%d failed)\n", g_failed ? "SELFTEST FAILED" : "SELFTEST PASSED", g_failed);
  return g_failed ? 1 : 0;
}
