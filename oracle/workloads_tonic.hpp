// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// MSIM_WL_TONIC — BASELINE config 5 / SURVEY §8(f) row 3: the tonic-example Greeter service under
// madsim-tonic with a kill/restart chaos schedule (tonic-example/tests/test.rs:155-202
// client_crash, widened from one client node to six; unary calls only).
//
//   Server::serve_with_shutdown   madsim-tonic/src/transport/server.rs:217-336
//       bind; loop { accept1; first message INLINE (Err = handshake: continue); spawn the
//       response task: select_biased!{ tx.closed() => return, stream.next() } ; tx.send }
//   Endpoint::connect / connect_ep  madsim-tonic/src/transport/channel.rs:75-111
//       a fresh madsim Endpoint per call + a throw-away handshake connection
//   MultiEndpoint::connect1         channel.rs:294-307
//   Grpc::unary                     madsim-tonic/src/client.rs:52-78
//   MyGreeter::say_hello            tonic-example/src/lib.rs:48-63 ("error" -> invalid_argument)
#pragma once
#include <array>

#include "rt/runtime.hpp"

namespace msim_oracle {

constexpr u32 TONIC_NAME = 0x696e6f54u;   // "Toni"
constexpr u32 TONIC_ERROR = 0x6f727265u;  // "erro"
enum : u32 { TONIC_OK = 0, TONIC_INVALID_ARGUMENT = 1 };

struct TonicParams {
  u32 n_clients = 6;
  u64 rounds = 10;
  u64 think_ns = 100 * MS;
  u32 flags = 0;  // 1: the supervisor never kills, only restarts (client_crash as written)
};
struct TonicStats {
  std::array<u32, 8> replies{}, starts{};
  u32 handled = 0;
};
inline u32 tonic_reply_for(u32 name, SocketAddr remote) { return name * 31u + (remote.ip & 0xff); }

/// the task server.rs:268 spawns per request
inline Async<Unit> tonic_response_task(Connection conn, Payload request, std::shared_ptr<TonicStats> st) {
  st->handled++;
  u32 name = request.v[0];
  Payload rsp;
  if (name == TONIC_ERROR) {  // Err(Status::invalid_argument("error!")): sent without the closed() race
    rsp.v[0] = TONIC_INVALID_ARGUMENT;
  } else {
    if (conn.tx.is_closed()) co_return {};  // select_biased!: tx.closed() is polled first
    rsp.v[0] = TONIC_OK;
    rsp.v[1] = tonic_reply_for(name, conn.peer);
  }
  (void)conn.tx.send(rsp);  // tx.send(..).await?
  co_return {};
}
inline Async<Unit> tonic_serve(SocketAddr addr, std::shared_ptr<TonicStats> st) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  for (;;) {
    Connection conn = std::move((co_await ep.accept1()).unwrap());
    auto msg = co_await conn.rx.recv();
    if (msg.is_err()) continue;  // maybe handshake or error
    spawn(tonic_response_task(std::move(conn), *msg.ok, st));
  }
}

/// Endpoint::connect_ep (channel.rs:94-111): Endpoint::connect + handshake connect1
inline Async<Result<Endpoint>> tonic_connect_ep(SocketAddr server) {
  auto r = co_await Endpoint::connect(server);
  if (r.is_err()) co_return r;
  auto hs = co_await r.ok->connect1(server);
  if (hs.is_err()) co_return Result<Endpoint>::Error(hs.err);
  co_return r;  // the handshake's (tx, rx) are gone by now
}
/// Grpc::unary (client.rs:52-78) over MultiEndpoint::connect1 (channel.rs:294-307)
inline Async<Result<Payload>> tonic_unary(SocketAddr server, Payload request) {
  using R = Result<Payload>;
  auto ep = co_await tonic_connect_ep(server);
  if (ep.is_err()) co_return R::Error(Err::ConnectionRefused);
  auto c = co_await ep.ok->connect1(server);
  if (c.is_err()) co_return R::Error(c.err);
  Connection conn = std::move(*c.ok);
  ep.ok.reset();  // madsim_ep drops; tx / rx keep the BindGuard alive
  if (conn.tx.send(request).is_err()) co_return R::Error(Err::ConnectionReset);
  auto rsp = co_await conn.rx.recv();
  if (rsp.is_err()) co_return R::Error(rsp.err);
  co_return R::Ok(*rsp.ok);
}

/// the client node's init task (tests/test.rs:176-194, unary part)
inline Async<Unit> tonic_client(u32 c, SocketAddr server, TonicParams p, std::shared_ptr<TonicStats> st) {
  st->starts[c]++;
  SocketAddr me{ip4(10, 0, 0, 2 + c), 0};
  {  // GreeterClient::connect: Endpoint::connect_inner only checks that the endpoint is reachable
    auto ep = co_await tonic_connect_ep(server);
    if (ep.is_err()) panic("GreeterClient::connect failed", 70);
  }
  GlobalRng& rng = thread_rng();
  for (u32 seq = 0;; seq++) {
    Payload q;
    q.v[0] = rng.gen_range_u64(0, 10) == 0 ? TONIC_ERROR : TONIC_NAME;
    q.v[1] = seq;
    auto r = co_await tonic_unary(server, q);
    if (r.is_err()) panic("say_hello failed", 71);
    if (q.v[0] == TONIC_ERROR) {
      if (r.ok->v[0] != TONIC_INVALID_ARGUMENT) panic("expected invalid_argument", 72);
    } else if (r.ok->v[0] != TONIC_OK || r.ok->v[1] != tonic_reply_for(TONIC_NAME, me)) {
      panic("unexpected reply", 73);
    }
    st->replies[c]++;
    co_await sleep(p.think_ns);
  }
}

inline Async<u32> wl_tonic_main(TonicParams p, std::array<u32, 4>* user) {
  auto& rt = current_runtime();
  SocketAddr server{ip4(10, 0, 0, 1), 50051};
  auto st = std::make_shared<TonicStats>();
  auto node0 = rt.create_node().ip(server.ip).build();
  node0.spawn(tonic_serve(server, st));
  co_await sleep(1 * SEC);
  std::vector<u64> ids;
  for (u32 c = 0; c < p.n_clients; c++) {
    auto nb = rt.create_node();
    nb.ip(ip4(10, 0, 0, 2 + c));
    nb.init([c, server, p, st]() { return tonic_client(c, server, p, st); });
    ids.push_back(nb.build().id());
  }
  GlobalRng& rng = thread_rng();
  for (u64 i = 0; i < p.rounds; i++) {
    co_await sleep(rng.gen_range_duration(0, 5 * SEC));
    u64 c = rng.gen_range_u64(0, p.n_clients);
    u64 k = rng.gen_range_u64(0, 4);
    if (k == 3 && !(p.flags & 1))
      rt.kill(ids[c]);
    else
      rt.restart(ids[c]);
  }
  co_await sleep(1 * SEC);
  u32 replies = 0, starts = 0;
  for (u32 c = 0; c < p.n_clients; c++) replies += st->replies[c], starts += st->starts[c];
  if (st->handled < replies) panic("more replies than requests handled", 74);
  (*user)[0] = replies;
  (*user)[1] = starts;
  (*user)[2] = st->handled;
  (*user)[3] = u32(current_net().msg_count());
  co_return 0;
}

}  // namespace msim_oracle
