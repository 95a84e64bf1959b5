// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// MSIM_WL_TONIC — BASELINE config 5 / SURVEY §8(f) row 3: the tonic-example Greeter service under
// madsim-tonic with a kill/restart chaos schedule (tonic-example/tests/test.rs:155-202
// client_crash, widened from one client node to six; unary calls and, with flag 2, server- and
// client-streaming and bidi calls).
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
enum : u32 { TONIC_OK = 0, TONIC_INVALID_ARGUMENT = 1, TONIC_EOF = 2, TONIC_TRAILER = 3, TONIC_UNAVAILABLE = 4, TONIC_UNIMPLEMENTED = 5 };
constexpr u32 TONIC_SERVER_STREAMING = 1u << 31;  // the `server_streaming: bool` of the first message (client.rs:128)
constexpr u32 TONIC_CLIENT_STREAMING = 1u << 30;  // the first message is the stream-start header Box<()> (client.rs:170-176)
enum : u32 { TONIC_RESTART_ONLY = 1, TONIC_STREAMING_MIX = 2, TONIC_SERVER_CRASH = 4, TONIC_TIMEOUTS = 8, TONIC_INTERCEPT = 16 };
constexpr u32 TONIC_UNKNOWN_SERVICE = 0x6e6b6e75u;  // "unkn": a path whose service the server does not have
constexpr u32 TONIC_DELAY = 0x616c6564u;  // "dela": AnotherGreeter::delay sleeps 10 s before it answers (lib.rs:36-43)

struct TonicParams {
  u32 n_clients = 6;
  u64 rounds = 10;
  u64 think_ns = 100 * MS;
  // 4: halfway through, the supervisor kills the SERVER (tests/test.rs:233-278 server_crash): calls in flight end in
  //    "broken pipe", later ones are refused; the clients count failures instead of unwrapping
  // 8: the channel has a 1 s timeout (Endpoint::timeout, tests/test.rs:369-410 request_timeout): unary calls run
  //    inside time::timeout; picks 7 / 8 call `delay` with the channel's 1 s / a 5 s request timeout and must
  //    come back DeadlineExceeded (streaming calls are not wrapped here)
  // 16: interceptors on both sides and an unknown service (tests/test.rs:280-363 unimplemented_service, interceptor):
  //    the server's GreeterServer::with_interceptor and every client's GreeterClient::with_interceptor let every
  //    second request through and answer the others Status::unavailable("intercepted") — the client's before any
  //    network activity —; pick 9 calls a service the server does not have and must come back Unimplemented
  u32 flags = 0;  // 1: the supervisor never kills, only restarts (client_crash as written); 2: clients also make
                  // server-streaming calls (lots_of_replies), sometimes dropping the response stream at once
};
struct TonicStats {
  std::array<u32, 8> replies{}, starts{};
  u32 handled = 0, failures = 0;
  u32 server_calls = 0;  // the `i` of the server-side interceptor closure (tests/test.rs:324-335)
};
inline u32 tonic_reply_for(u32 name, SocketAddr remote) { return name * 31u + (remote.ip & 0xff); }

/// the task server.rs:268 spawns per request
/// server.rs:259-272 (no such service) and an intercepted call: the spawned task only sends the error status
inline Async<Unit> tonic_status_task(Connection conn, u32 status) {
  Payload rsp;
  rsp.v[0] = status;
  (void)conn.tx.send(rsp);
  co_return {};
}
inline Async<Unit> tonic_response_task(Connection conn, Payload request, std::shared_ptr<TonicStats> st) {
  st->handled++;
  u32 name = request.v[0];
  if (name == TONIC_DELAY) co_await sleep(10 * SEC);  // inside rsp_future.await, before the closed() race
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
/// select_biased!{ _ = tx.closed() => Closed, item = stream.next() => .. } where the stream is
/// parked in `sleep(1 s)` (server.rs:289-300 around lots_of_replies' generator, lib.rs:72-81)
struct ClosedOrSleep {
  UnboundedSender<std::pair<Payload, LinkState>>::ClosedFut closed;
  Sleep delay;
  Poll<bool> poll(Context& cx) {  // true = the client closed
    if (closed.poll(cx).ready()) return Poll<bool>::done(true);
    if (delay.poll(cx).ready()) return Poll<bool>::done(false);
    return Poll<bool>::pending();
  }
};
/// the response task of a server-streaming call: header, then the stream's items, then the trailer
inline Async<Unit> tonic_stream_response_task(Connection conn, Payload request, std::shared_ptr<TonicStats> st) {
  st->handled++;
  u32 name = request.v[0];
  Payload header;
  header.v[0] = TONIC_OK;
  if (conn.tx.send(header).is_err()) co_return {};
  for (u32 i = 0; i < 3; i++) {  // lots_of_replies: for i in 0..3 { yield reply; sleep(1 s).await }
    if (conn.tx.is_closed()) co_return {};
    Payload item;
    item.v[0] = TONIC_OK;
    item.v[1] = tonic_reply_for(name, conn.peer) + i;
    if (conn.tx.send(item).is_err()) co_return {};
    if (conn.tx.is_closed()) co_return {};
    if (co_await ClosedOrSleep{conn.tx.closed(), sleep(1 * SEC)}) co_return {};
  }
  Payload eof;  // Err(Status::unknown("EOF"))?: the stream's last item
  eof.v[0] = TONIC_EOF;
  if (conn.tx.send(eof).is_err()) co_return {};
  if (conn.tx.is_closed()) co_return {};
  Payload trailer;
  trailer.v[0] = TONIC_TRAILER;
  (void)conn.tx.send(trailer);
  co_return {};
}
/// the response task of a client-streaming call (lots_of_greetings, lib.rs:85-103): the service reads the
/// request stream until the client's Sender drops, then answers like a unary call
inline u32 tonic_fold_names(u32 acc, u32 name) { return acc * 33u + name; }
inline Async<Unit> tonic_client_stream_response_task(Connection conn, std::shared_ptr<TonicStats> st) {
  st->handled++;
  u32 acc = 0;
  for (;;) {  // while let Some(request) = stream.message().await?
    auto m = co_await conn.rx.recv();
    if (m.is_err()) break;
    acc = tonic_fold_names(acc, m.ok->v[0]);
  }
  if (conn.tx.is_closed()) co_return {};
  Payload rsp;
  rsp.v[0] = TONIC_OK;
  rsp.v[1] = tonic_reply_for(acc, conn.peer);
  (void)conn.tx.send(rsp);
  co_return {};
}
/// select_biased!{ _ = tx.closed() => Closed, item = stream.next() => .. } where the response stream is
/// parked in the REQUEST stream's rx.recv() (bidi_hello, lib.rs:106-124)
struct ClosedOrRecv {
  UnboundedSender<std::pair<Payload, LinkState>>::ClosedFut closed;
  Async<Result<Payload>> recv;
  Poll<int> poll(Context& cx) {  // 2 = the client closed, 1 = a request (in *out), 0 = the request stream ended
    if (closed.poll(cx).ready()) return Poll<int>::done(2);
    auto r = recv.poll(cx);
    if (!r.ready()) return Poll<int>::pending();
    if (r.v->is_err()) return Poll<int>::done(0);
    *out = *r.v->ok;
    return Poll<int>::done(1);
  }
  Payload* out;
};
/// the response task of a bidi call: header, one reply per request, trailer when the request stream ends
inline Async<Unit> tonic_bidi_response_task(Connection conn, std::shared_ptr<TonicStats> st) {
  st->handled++;
  Payload header;
  header.v[0] = TONIC_OK;
  if (conn.tx.send(header).is_err()) co_return {};
  for (;;) {
    Payload req;
    int k = co_await ClosedOrRecv{conn.tx.closed(), conn.rx.recv(), &req};
    if (k == 2) co_return {};
    if (k == 0) break;
    Payload item;
    item.v[0] = TONIC_OK;
    item.v[1] = tonic_reply_for(req.v[0], conn.peer);
    if (conn.tx.send(item).is_err()) co_return {};
  }
  Payload trailer;
  trailer.v[0] = TONIC_TRAILER;
  (void)conn.tx.send(trailer);
  co_return {};
}
inline Async<Unit> tonic_serve(SocketAddr addr, std::shared_ptr<TonicStats> st, bool intercept = false) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  for (;;) {
    Connection conn = std::move((co_await ep.accept1()).unwrap());
    auto msg = co_await conn.rx.recv();
    if (msg.is_err()) continue;  // maybe handshake or error
    if ((msg.ok->v[1] & TONIC_CLIENT_STREAMING) && (msg.ok->v[1] & TONIC_SERVER_STREAMING))
      spawn(tonic_bidi_response_task(std::move(conn), st));
    else if (msg.ok->v[1] & TONIC_CLIENT_STREAMING)
      spawn(tonic_client_stream_response_task(std::move(conn), st));
    else if (msg.ok->v[1] & TONIC_SERVER_STREAMING)
      spawn(tonic_stream_response_task(std::move(conn), *msg.ok, st));
    else if (intercept && msg.ok->v[0] == TONIC_UNKNOWN_SERVICE)  // self.services.get_mut(svc_name) is None
      spawn(tonic_status_task(std::move(conn), TONIC_UNIMPLEMENTED));
    else if (intercept && ++st->server_calls % 2 == 0)  // InterceptedService::call runs the closure HERE, in the serve loop
      spawn(tonic_status_task(std::move(conn), TONIC_UNAVAILABLE));
    else
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

/// Grpc::server_streaming (client.rs:110-136) + Streaming::message (codec.rs:30-44): read all three
/// replies and the EOF status, or drop the response stream right after the header
/// (tests/test.rs:204-230 client_drops_response_stream)
inline Async<Result<u32>> tonic_server_streaming(SocketAddr server, Payload request, bool drop_at_once, u32 expect0) {
  using R = Result<u32>;
  auto ep = co_await tonic_connect_ep(server);
  if (ep.is_err()) co_return R::Error(Err::ConnectionRefused);
  auto c = co_await ep.ok->connect1(server);
  if (c.is_err()) co_return R::Error(c.err);
  Connection conn = std::move(*c.ok);
  ep.ok.reset();
  request.v[1] |= TONIC_SERVER_STREAMING;
  if (conn.tx.send(request).is_err()) co_return R::Error(Err::ConnectionReset);
  auto header = co_await conn.rx.recv();
  if (header.is_err()) co_return R::Error(header.err);
  if (header.ok->v[0] != TONIC_OK) panic("bad streaming header", 75);
  if (drop_at_once) co_return R::Ok(0);
  for (u32 i = 0; i < 3; i++) {
    auto m = co_await conn.rx.recv();
    if (m.is_err()) co_return R::Error(m.err);
    if (m.ok->v[0] != TONIC_OK || m.ok->v[1] != expect0 + i) panic("unexpected streamed reply", 76);
  }
  auto e = co_await conn.rx.recv();
  if (e.is_err()) co_return R::Error(e.err);
  if (e.ok->v[0] != TONIC_EOF) panic("expected the EOF status", 77);
  co_return R::Ok(3);
}

/// Grpc::client_streaming (client.rs:82-108) + send_request_stream (:171-196) over hello_stream
/// (tests/test.rs:119-129: three requests, 1 s apart); the Sender drops when the stream ends
inline Async<Result<u32>> tonic_client_streaming(SocketAddr server, u32 seq, SocketAddr me) {
  using R = Result<u32>;
  auto ep = co_await tonic_connect_ep(server);
  if (ep.is_err()) co_return R::Error(Err::ConnectionRefused);
  auto c = co_await ep.ok->connect1(server);
  if (c.is_err()) co_return R::Error(c.err);
  Connection conn = std::move(*c.ok);
  ep.ok.reset();
  u32 acc = 0;
  {
    ConnSender tx = std::move(conn.tx);  // moved into send_request_stream
    Payload header;
    header.v[1] = seq | TONIC_CLIENT_STREAMING;
    if (tx.send(header).is_err()) co_return R::Error(Err::ConnectionReset);
    for (u32 i = 0; i < 3; i++) {
      Payload item;
      item.v[0] = TONIC_NAME + i;
      acc = tonic_fold_names(acc, item.v[0]);
      if (tx.send(item).is_err()) break;  // "allows the server to prematurely close the stream"
      co_await sleep(1 * SEC);
    }
  }
  auto rsp = co_await conn.rx.recv();
  if (rsp.is_err()) co_return R::Error(rsp.err);
  if (rsp.ok->v[0] != TONIC_OK || rsp.ok->v[1] != tonic_reply_for(acc, me)) panic("unexpected client-streaming reply", 78);
  co_return R::Ok(3);
}

/// Grpc::streaming (client.rs:139-166): the requests go out from a BACKGROUND task that owns the Sender;
/// Streaming::new(rx, Some(task)) (codec.rs:19-47) cancels that task when the response stream is dropped
/// after its first poll (`let _task = task.cancel_on_drop()` is the generator's first statement).
inline Async<Unit> tonic_bidi_sender(ConnSender tx, u32 seq) {
  Payload header;
  header.v[1] = seq | TONIC_CLIENT_STREAMING | TONIC_SERVER_STREAMING;
  if (tx.send(header).is_err()) co_return {};
  for (u32 i = 0; i < 3; i++) {  // hello_stream (tests/test.rs:119-129)
    Payload item;
    item.v[0] = TONIC_NAME + i;
    if (tx.send(item).is_err()) break;
    co_await sleep(1 * SEC);
  }
  co_return {};
}
struct TonicBidiStream {
  ConnReceiver rx;  // destroyed AFTER the body below ran: `_task` drops before the captured Receiver
  std::optional<JoinHandle<Unit>> task;
  bool armed = false;
  TonicBidiStream(ConnReceiver r, JoinHandle<Unit> t) : rx(std::move(r)), task(std::move(t)) {}
  TonicBidiStream(const TonicBidiStream&) = delete;
  void end_of_stream() {  // the generator returns: `_task` drops (cancel is a no-op on a finished task)
    if (armed && task) task->task->cancel();
    task.reset();
  }
  ~TonicBidiStream() { end_of_stream(); }
};
inline Async<Result<u32>> tonic_bidi(SocketAddr server, u32 seq, SocketAddr me, bool drop_after_first) {
  using R = Result<u32>;
  auto ep = co_await tonic_connect_ep(server);
  if (ep.is_err()) co_return R::Error(Err::ConnectionRefused);
  auto c = co_await ep.ok->connect1(server);
  if (c.is_err()) co_return R::Error(c.err);
  Connection conn = std::move(*c.ok);
  ep.ok.reset();
  auto task = spawn(tonic_bidi_sender(std::move(conn.tx), seq));
  auto header = co_await conn.rx.recv();
  if (header.is_err()) co_return R::Error(header.err);
  if (header.ok->v[0] != TONIC_OK) panic("bad bidi header", 75);
  TonicBidiStream stream(std::move(conn.rx), std::move(task));
  u32 i = 0;
  for (;;) {  // while let Some(reply) = stream.message().await.unwrap()
    stream.armed = true;
    auto m = co_await stream.rx.recv();
    if (m.is_err()) co_return R::Error(m.err);
    if (m.ok->v[0] == TONIC_TRAILER) {
      stream.end_of_stream();
      break;
    }
    if (m.ok->v[0] != TONIC_OK || m.ok->v[1] != tonic_reply_for(TONIC_NAME + i, me)) panic("unexpected bidi reply", 79);
    i++;
    if (drop_after_first) co_return R::Ok(1);  // the stream drops here: cancel the sender, then close the Receiver
  }
  if (i != 3) panic("bidi stream ended early", 79);
  co_return R::Ok(3);
}

/// the client node's init task (tests/test.rs:176-194, unary part)
inline Async<Unit> tonic_client(u32 c, SocketAddr server, TonicParams p, std::shared_ptr<TonicStats> st) {
  st->starts[c]++;
  SocketAddr me{ip4(10, 0, 0, 2 + c), 0};
  const bool tolerant = p.flags & TONIC_SERVER_CRASH;
  for (;;) {  // GreeterClient::connect: Endpoint::connect_inner only checks that the endpoint is reachable
    auto ep = co_await tonic_connect_ep(server);
    if (ep.is_ok()) break;
    if (!tolerant) panic("GreeterClient::connect failed", 70);
    st->failures++;
    ep.ok.reset();
    co_await sleep(p.think_ns);
  }
  GlobalRng& rng = thread_rng();
  u32 icount = 0;  // the `i` of this client's interceptor closure: a restarted node builds a new client
  for (u32 seq = 0;; seq++) {
    Payload q;
    u64 pick = rng.gen_range_u64(0, 10);
    q.v[0] = pick == 0 ? TONIC_ERROR : TONIC_NAME;
    q.v[1] = seq;
    if (p.flags & TONIC_INTERCEPT) {
      if (++icount % 2 == 0) {  // request.intercept(..)? fails before connect1 (client.rs:65): nothing reaches the network
        st->replies[c]++;
        co_await sleep(p.think_ns);
        continue;
      }
      if (pick == 9) q.v[0] = TONIC_UNKNOWN_SERVICE;
    }
    if ((p.flags & TONIC_STREAMING_MIX) && pick >= 1 && pick <= 3) {  // lots_of_replies; pick 3 drops the stream at once
      auto sr = co_await tonic_server_streaming(server, q, pick == 3, tonic_reply_for(TONIC_NAME, me));
      if (sr.is_err()) {
        if (!tolerant) panic("lots_of_replies failed", 71);
        st->failures++;
        co_await sleep(p.think_ns);
        continue;
      }
      st->replies[c]++;
      co_await sleep(p.think_ns);
      continue;
    }
    if ((p.flags & TONIC_STREAMING_MIX) && (pick == 5 || pick == 6)) {  // bidi_hello; pick 6 drops the stream after one reply
      auto sr = co_await tonic_bidi(server, seq, me, pick == 6);
      if (sr.is_err()) {
        if (!tolerant) panic("bidi_hello failed", 71);
        st->failures++;
        co_await sleep(p.think_ns);
        continue;
      }
      st->replies[c]++;
      co_await sleep(p.think_ns);
      continue;
    }
    if ((p.flags & TONIC_STREAMING_MIX) && pick == 4) {  // lots_of_greetings
      auto sr = co_await tonic_client_streaming(server, seq, me);
      if (sr.is_err()) {
        if (!tolerant) panic("lots_of_greetings failed", 71);
        st->failures++;
        co_await sleep(p.think_ns);
        continue;
      }
      st->replies[c]++;
      co_await sleep(p.think_ns);
      continue;
    }
    u64 tmo = 0;
    if (p.flags & TONIC_TIMEOUTS) {
      tmo = 1 * SEC;
      if (pick == 7 || pick == 8) q.v[0] = TONIC_DELAY;
      if (pick == 8) tmo = 5 * SEC;  // Request::set_timeout overrides the channel's
    }
    Result<Payload> r = Result<Payload>::Error(Err::Elapsed);
    if (tmo) {  // with_timeout (client.rs:246-258): the Timeout future is consumed by the await
      auto t = timeout(tmo, tonic_unary(server, q));
      auto tr = co_await t;
      if (tr.is_ok()) r = std::move(*tr.ok);
    } else {
      r = co_await tonic_unary(server, q);
    }
    if (q.v[0] == TONIC_DELAY) {
      if (r.is_ok()) panic("delay must end in DeadlineExceeded", 80);
      if (r.err == Err::Elapsed) {
        st->replies[c]++;
        co_await sleep(p.think_ns);
        continue;
      }
    }
    if (r.is_err()) {
      if (!tolerant) panic("say_hello failed", 71);
      st->failures++;
      co_await sleep(p.think_ns);
      continue;
    }
    if (q.v[0] == TONIC_UNKNOWN_SERVICE) {
      if (r.ok->v[0] != TONIC_UNIMPLEMENTED) panic("expected Unimplemented", 93);
    } else if ((p.flags & TONIC_INTERCEPT) && r.ok->v[0] == TONIC_UNAVAILABLE) {
      // the server's interceptor turned this one down
    } else if (q.v[0] == TONIC_ERROR) {
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
  node0.spawn(tonic_serve(server, st, (p.flags & TONIC_INTERCEPT) != 0));
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
    if ((p.flags & TONIC_SERVER_CRASH) && i == p.rounds / 2)
      rt.kill(node0.id());
    else if (k == 3 && !(p.flags & TONIC_RESTART_ONLY))
      rt.kill(ids[c]);
    else
      rt.restart(ids[c]);
  }
  co_await sleep(1 * SEC);
  u32 replies = 0, starts = 0;
  for (u32 c = 0; c < p.n_clients; c++) replies += st->replies[c], starts += st->starts[c];
  if (!(p.flags & (TONIC_TIMEOUTS | TONIC_INTERCEPT)) && st->handled < replies) panic("more replies than requests handled", 74);
  (*user)[0] = replies;
  (*user)[1] = starts;
  (*user)[2] = st->handled;
  (*user)[3] = (p.flags & TONIC_SERVER_CRASH) ? st->failures : u32(current_net().msg_count());
  co_return 0;
}

}  // namespace msim_oracle
