// MSIM_WL_TONIC — BASELINE config 5 / SURVEY §8(f) row 3: the tonic-example Greeter service under
// madsim-tonic while a supervisor kills / restarts the client nodes
// (tonic-example/tests/test.rs:155-202 client_crash, widened to six client nodes; unary calls and, with
// flag 2, server-streaming lots_of_replies, client-streaming lots_of_greetings and bidi bidi_hello).
//
//   Server::serve_with_shutdown    madsim-tonic/src/transport/server.rs:217-336
//   Endpoint::connect / connect_ep madsim-tonic/src/transport/channel.rs:75-111  (fresh endpoint + handshake per call)
//   MultiEndpoint::connect1        channel.rs:294-307
//   Grpc::unary                    madsim-tonic/src/client.rs:52-78
//   MyGreeter::say_hello           tonic-example/src/lib.rs:48-63
//   Handle::{kill,restart}         madsim/src/sim/task/mod.rs:356-401: a restarted node's old sockets
//                                  stay bound (BindGuard::drop skips killed incarnations, net/mod.rs:483-493)
//
// Oracle twin: oracle/workloads_tonic.hpp.
#pragma once
#include "../engine/engine.h"

namespace msim {

/// BIG: the capacities of the request-timeout schedule (flag 8): every pending poll inside time::timeout adds a
/// timer that lingers 1-5 s, and `delay` response tasks live 10 s.
/// STREAMING: the program also makes server- / client-streaming and bidi calls (flag 2).  The plain unary program
/// (BASELINE config 5) is its own instantiation: none of the streaming state machines, no Sender::closed() waits,
/// no split connection ends, two queued payloads per direction instead of four — 25 seeds per CTA instead of 20
/// and a hot loop a third smaller (the kernels are instruction-fetch bound).
/// INTERCEPT: the interceptor / unknown-service schedule (flag 16; tests/test.rs:280-363): GreeterServer::with_interceptor
/// and every GreeterClient::with_interceptor turn every second request down with Status::unavailable — the client's
/// before anything reaches the network —, and pick 9 calls a service the server does not have (Unimplemented).
template <bool BIG, bool STREAMING = true, bool INTERCEPT = false>
struct WL_TonicT {
  enum : int { S_HANDLED = 0, S_CNT, S_ROUND, S_C, S_FAILURES, S_SERVER_CALLS, S_END };  // wl_s scalars
  enum : int { F_REPLIES = 0, F_STARTS, F_ICOUNT, F_NFIELD };           // wl_w per client node
  struct Caps {
    static constexpr bool HEAP16 = false;  // split timer arrays: the 16-byte records measured no gain here (profiles/r2k_footprint_experiments.txt)
    // The unary programs (config 5 and its interceptor variant) run on tables cut to what the BASELINE-size batch needs with
    // a 1.5x margin (32 768 seeds of config 5 peak at 10 tasks / 11 timers; no overflow at 16 / 32 / 16 sockets / 24
    // connections): 6 080 B per seed instead of 9 248, i.e. 28 warps per SM instead of 25 (+12 %, profiles/r2n).  A seed of
    // some other parameter set that needs more re-runs on the large variant (capacity escalation), which for these
    // programs is 8x / 4x the tables (BIG_SCALE), at least the size of the streaming program's.
    static constexpr bool SMALL = !BIG && !STREAMING;
    static constexpr int BIG_SCALE = SMALL ? 2 : 1;
    static constexpr int NT = BIG ? 80 : SMALL ? 16 : 32, TQ = BIG ? 768 : SMALL ? 32 : 64, NMSG = 2, NSOCK = SMALL ? 16 : 24, NREG = 1, NMBOX = 1, NNODE = 8, NPARK = 1;
    static constexpr int WQ = 0, WW = INTERCEPT ? F_NFIELD : F_ICOUNT, WS = INTERCEPT ? S_END : S_SERVER_CALLS;  // the interceptor counters exist only in that variant
    static constexpr bool BIND_INLINE = true;  // every call binds an ephemeral endpoint: Network::bind inline, +6 % (profiles/r2w_bind_modes.txt)
    static constexpr bool USES_JOIN = false;
    static constexpr bool USES_KILL = true;
    static constexpr bool USES_CTRLC = false;  // does any task await signal::ctrl_c()?
    static constexpr bool USES_CLOSED = STREAMING;  // does any task await Sender::closed()? (the streaming response tasks)
    static constexpr bool USES_SPLIT = STREAMING;   // does any task give the Sender half of its connection to another task? (bidi)
    static constexpr bool USES_TIMEOUT = BIG;    // time::timeout around the unary calls: only the request-timeout schedule (flag 8 = the BIG variant)
    static constexpr bool USES_CONN2 = false;    // does any task hold two connection ends at once?
    static constexpr bool USES_PMATCH = false;   // does any node use restart_on_panic_matching?
    static constexpr unsigned PMATCH_BASE = 0;  // the panic code bit 0 of a restart_on_panic_matching mask stands for
    static constexpr bool COLD_SPAWN = true, COLD_FINISH = MSIM_COLD_DEFAULT & 2, COLD_CH = MSIM_COLD_DEFAULT & 4;  // spawn out of line: +3 % (profiles/r2x_placement_sweep.txt)
    static constexpr int NCONN = BIG ? 64 : SMALL ? 24 : 32, NCQ = STREAMING ? 4 : 2, NACC = SMALL ? 16 : 32;  // a client-streaming call queues header + 3 items at a slow server
  };
  enum : u32 { K_MAIN = 0, K_SERVE, K_RSP, K_CLIENT, K_RSP_STREAM, K_RSP_CSTREAM, K_RSP_BIDI, K_SENDER };
  enum : u32 { NAME = 0x696e6f54u, ERROR_NAME = 0x6f727265u, DELAY_NAME = 0x616c6564u, UNKNOWN_SERVICE = 0x6e6b6e75u };
  enum : u32 { ST_OK = 0, ST_INVALID_ARGUMENT = 1, ST_EOF = 2, ST_TRAILER = 3, ST_UNAVAILABLE = 4, ST_UNIMPLEMENTED = 5 };
  static constexpr u32 STATUS_ONLY = 1u << 29;  // the spawned task only sends the status in its first local (server.rs:259-272)
  enum : u32 { SERVER_STREAMING = 1u << 31, CLIENT_STREAMING = 1u << 30, FLAG_RESTART_ONLY = 1, FLAG_STREAMING_MIX = 2, FLAG_SERVER_CRASH = 4, FLAG_TIMEOUTS = 8, FLAG_INTERCEPT = 16 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 50051u;  // 10.0.0.1:50051
  /// NodeBuilder::init(|| tonic_client(c)) — what Handle::restart re-runs
  static MSIM_HD int init_kind(int) { return K_CLIENT; }
  static MSIM_HD u32 reply_for(u32 name, u32 remote_addr) { return name * 31u + addr_node(remote_addr); }
  template <class E>
  static MSIM_HD int panic(E& e, u32 code) {
    e.ya = code;
    return RQ_PANIC;
  }

  /// A failed call: `.unwrap()` panics — except in the server-crash schedule (tests/test.rs:233-278), where the
  /// client counts the failure, drops what the call held (`_task`, the connection, the endpoint) and goes on
  template <class E>
  static MSIM_HD int call_failed(E& e, int t, u32 code) {
    const bool deadline = e.ret == -3 && code == 71 && e.t_a(t) == DELAY_NAME;  // what `delay` must end in
    if (!deadline && !(e.P->p[2] & FLAG_SERVER_CRASH)) return panic(e, code);
    if (Caps::USES_SPLIT && e.t_handle(t) != NONE8) {
      if (e.t_handle(t) & 0x80)
        e.join_detach(e.t_handle(t) & 0x7f);
      else
        e.task_cancel_detach(e.t_handle(t));
      e.t_handle(t) = NONE8;
    }
    e.conn_drop_end(t);
    e.sock_unref(t);
    if (Caps::USES_TIMEOUT) e.t_dl2(t) = E::NONE64;  // the Timeout future is gone either way
    if (deadline)
      e.wl_w(F_REPLIES, e.t_node(t))++;
    else
      e.wl_s(S_FAILURES)++;
    if (code == 70) {
      e.t_pc(t) = 18;  // GreeterClient::connect again
    } else {
      e.t_b(t)++;
      e.t_pc(t) = 3;
    }
    e.xa = e.P->p[1];
    return RQ_SLEEP_FOR;
  }

  template <class E>
  static MSIM_HD int fsm(E& e, int t) {
    using L = typename E::L;
    const Params& P = *e.P;
    const u32 n_clients = P.n_nodes - 1;
    // kinds K_RSP_STREAM.. exist only in the streaming program: masking the kind tells the compiler so (their cases go)
    for (;;) switch ((STREAMING ? (u32)e.t_kind(t) : ((u32)e.t_kind(t) & 3u)) * 32 + e.t_pc(t)) {
        // ---- wl_tonic_main ----
        case K_MAIN * 32 + 0:
          e.t_pc(t) = 1;
          e.ya = NODE_IP;
          return RQ_CREATE_NODE;
        case K_MAIN * 32 + 1:  // node0.spawn(tonic_serve)
          e.t_pc(t) = 2;
          e.ya = K_SERVE, e.yb = 1, e.yc = e.yd = 0, e.xa = e.xb = 0;
          return RQ_SPAWN;
        case K_MAIN * 32 + 2:
          e.t_pc(t) = 3;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        case K_MAIN * 32 + 3:  // client nodes: create_node().ip().init(tonic_client).build()
          if (e.wl_s(S_CNT) < n_clients) {
            e.wl_s(S_CNT)++;
            e.ya = NODE_IP | NODE_INIT;
            return RQ_CREATE_NODE;
          }
          e.t_pc(t) = 4;
          // fallthrough
        case K_MAIN * 32 + 4:  // for i in 0..rounds { sleep(gen_range(0 s..5 s))
          if ((u64)e.wl_s(S_ROUND) >= P.p[0]) {
            e.t_pc(t) = 8;
            e.xa = NS_PER_SEC;
            return RQ_SLEEP_FOR;
          }
          e.t_pc(t) = 5;
          e.xa = e.sample_inclusive_u64(0, 5 * NS_PER_SEC - 1);  // UniformDuration Medium
          return RQ_SLEEP_FOR;
        case K_MAIN * 32 + 5:  // c = gen_range(0..n_clients)
          e.t_pc(t) = 6;
          e.xa = 0, e.xb = n_clients;
          return RQ_GEN_RANGE;
        case K_MAIN * 32 + 6:  // k = gen_range(0..4)
          e.wl_s(S_C) = (u32)e.ret64;
          e.t_pc(t) = 7;
          e.xa = 0, e.xb = 4;
          return RQ_GEN_RANGE;
        case K_MAIN * 32 + 7: {
          int id = 2 + (int)e.wl_s(S_C);
          if ((P.p[2] & FLAG_SERVER_CRASH) && (u64)e.wl_s(S_ROUND) == P.p[0] / 2)
            e.kill(1);
          else if (e.ret64 == 3 && !(P.p[2] & FLAG_RESTART_ONLY))
            e.kill(id);
          else
            e.restart(id);
          e.wl_s(S_ROUND)++;
          e.t_pc(t) = 4;
          continue;
        }
        case K_MAIN * 32 + 8: {
          u32 replies = 0, starts = 0;
          for (u32 c = 0; c < n_clients; c++) replies += e.wl_w(F_REPLIES, 2 + c), starts += e.wl_w(F_STARTS, 2 + c);
          if (!(P.p[2] & (FLAG_TIMEOUTS | FLAG_INTERCEPT)) && e.wl_s(S_HANDLED) < replies) return panic(e, 74);
          e.w(L::W_USER0) = replies;
          e.w(L::W_USER1) = starts;
          e.w(L::W_USER2) = e.wl_s(S_HANDLED);
          e.w(L::W_USER3) = (P.p[2] & FLAG_SERVER_CRASH) ? e.wl_s(S_FAILURES) : e.w(L::W_MSG_COUNT);
          return RQ_READY;
        }

        // ---- tonic_serve ----
        case K_SERVE * 32 + 0:
          e.t_pc(t) = 1;
          e.ya = SERVER_ADDR;
          return RQ_BIND;
        case K_SERVE * 32 + 1:
          if (e.ret < 0) return panic(e, 0);
          e.t_pc(t) = 2;
          // fallthrough
        case K_SERVE * 32 + 2:
          e.t_pc(t) = 3;
          return RQ_ACCEPT1;
        case K_SERVE * 32 + 3:  // the connection's first message, awaited inline
          if (e.ret < 0) return panic(e, 0);
          e.t_b(t) = e.rfrom;
          e.t_pc(t) = 4;
          return RQ_CH_RECV;
        case K_SERVE * 32 + 4:
          if (e.ret < 0) {  // Err(_) => continue: a handshake (or a killed client); tx, rx drop
            e.conn_drop_end(t);
            e.t_pc(t) = 2;
            continue;
          }
          e.t_pc(t) = 2;  // spawn the response task with tx (and the request)
          if (INTERCEPT && (P.p[2] & FLAG_INTERCEPT)) {
            u32 status = ST_OK;
            if (e.rv0 == UNKNOWN_SERVICE)
              status = ST_UNIMPLEMENTED;  // self.services.get_mut(svc_name) is None
            else if (++e.wl_s(S_SERVER_CALLS) % 2 == 0)
              status = ST_UNAVAILABLE;  // InterceptedService::call runs the closure here, in the serve loop
            if (status != ST_OK) {
              e.ya = K_RSP | SPAWN_CONN, e.yb = e.t_node(t), e.yc = status, e.yd = e.t_b(t), e.xa = STATUS_ONLY, e.xb = 0;
              return RQ_SPAWN;
            }
          }
          e.ya = (!STREAMING ? (u32)K_RSP
                  : ((e.rv1 & CLIENT_STREAMING) && (e.rv1 & SERVER_STREAMING)) ? (u32)K_RSP_BIDI
                  : (e.rv1 & CLIENT_STREAMING) ? (u32)K_RSP_CSTREAM
                  : (e.rv1 & SERVER_STREAMING) ? (u32)K_RSP_STREAM
                                               : (u32)K_RSP) | SPAWN_CONN, e.yb = e.t_node(t), e.yc = e.rv0, e.yd = e.t_b(t), e.xa = e.rv1, e.xb = 0;
          return RQ_SPAWN;

        // ---- tonic_response_task ----
        case K_RSP * 32 + 0:
          if (INTERCEPT && (e.t_c(t) & STATUS_ONLY)) {  // no such service / turned down by the interceptor
            e.t_pc(t) = 1;
            e.yb = e.t_a(t), e.yc = 0;
            return RQ_CH_SEND;
          }
          e.wl_s(S_HANDLED)++;
          if (e.t_a(t) == DELAY_NAME) {  // AnotherGreeter::delay: sleep(10 s) inside rsp_future.await
            e.t_pc(t) = 2;
            e.xa = 10 * NS_PER_SEC;
            return RQ_SLEEP_FOR;
          }
          // fallthrough
        case K_RSP * 32 + 2:
          e.t_pc(t) = 1;
          if (e.t_a(t) == ERROR_NAME) {
            e.yb = ST_INVALID_ARGUMENT, e.yc = 0;
            return RQ_CH_SEND;
          }
          if (e.conn_tx_closed(t)) return RQ_READY;  // select_biased!{ tx.closed() => return, .. }
          e.yb = ST_OK, e.yc = reply_for(e.t_a(t), e.t_b(t));
          return RQ_CH_SEND;
        case K_RSP * 32 + 1:
          return RQ_READY;

        // ---- tonic_stream_response_task: header, lots_of_replies' items, trailer (server.rs:274-313) ----
        case K_RSP_STREAM * 32 + 0:
          e.wl_s(S_HANDLED)++;
          e.t_c(t) = 0;  // i
          e.t_pc(t) = 1;
          e.yb = ST_OK, e.yc = 0;
          return RQ_CH_SEND;
        case K_RSP_STREAM * 32 + 1:  // tx.send(header).await?
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 2;
          // fallthrough
        case K_RSP_STREAM * 32 + 2:  // select_biased!{ closed => return, stream.next() => reply i }
          if (e.conn_tx_closed(t)) return RQ_READY;
          e.t_pc(t) = 3;
          e.yb = ST_OK, e.yc = reply_for(e.t_a(t), e.t_b(t)) + e.t_c(t);
          return RQ_CH_SEND;
        case K_RSP_STREAM * 32 + 3:  // the next select finds the generator in `sleep(1 s)`
          if (e.ret < 0) return RQ_READY;
          if (e.conn_tx_closed(t)) return RQ_READY;
          e.t_pc(t) = 4;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_OR_CLOSED;
        case K_RSP_STREAM * 32 + 4:
          if (e.ret < 0) return RQ_READY;  // the client dropped the response stream
          e.t_c(t)++;
          if (e.t_c(t) < 3) {
            e.t_pc(t) = 2;
            continue;
          }
          e.t_pc(t) = 5;  // Err(Status::unknown("EOF"))?: the stream's last item
          e.yb = ST_EOF, e.yc = 0;
          return RQ_CH_SEND;
        case K_RSP_STREAM * 32 + 5:  // stream.next() == None: break; the trailer
          if (e.ret < 0) return RQ_READY;
          if (e.conn_tx_closed(t)) return RQ_READY;
          e.t_pc(t) = 6;
          e.yb = ST_TRAILER, e.yc = 0;
          return RQ_CH_SEND;
        case K_RSP_STREAM * 32 + 6:
          return RQ_READY;

        // ---- tonic_client_stream_response_task: lots_of_greetings reads the request stream to its end ----
        case K_RSP_CSTREAM * 32 + 0:
          e.wl_s(S_HANDLED)++;
          e.t_c(t) = 0;  // fold of the names
          e.t_pc(t) = 1;
          return RQ_CH_RECV;
        case K_RSP_CSTREAM * 32 + 1:
          if (e.ret >= 0) {  // while let Some(request) = stream.message().await?
            e.t_c(t) = e.t_c(t) * 33u + e.rv0;
            return RQ_CH_RECV;
          }
          if (e.conn_tx_closed(t)) return RQ_READY;  // select_biased!{ tx.closed() => return, .. }
          e.t_pc(t) = 2;
          e.yb = ST_OK, e.yc = reply_for(e.t_c(t), e.t_b(t));
          return RQ_CH_SEND;
        case K_RSP_CSTREAM * 32 + 2:
          return RQ_READY;

        // ---- tonic_bidi_response_task: bidi_hello answers every request until the request stream ends ----
        case K_RSP_BIDI * 32 + 0:
          e.wl_s(S_HANDLED)++;
          e.t_pc(t) = 1;
          e.yb = ST_OK, e.yc = 0;
          return RQ_CH_SEND;
        case K_RSP_BIDI * 32 + 1:
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 2;
          // fallthrough
        case K_RSP_BIDI * 32 + 2:  // select_biased!{ tx.closed() => return, stream.next() => .. } parked in rx.recv()
          e.t_pc(t) = 3;
          return RQ_CH_RECV_OR_CLOSED;
        case K_RSP_BIDI * 32 + 3:
          if (e.ret == -2) return RQ_READY;
          if (e.ret < 0) {  // the request stream ended: None => break; the trailer
            e.t_pc(t) = 5;
            e.yb = ST_TRAILER, e.yc = 0;
            return RQ_CH_SEND;
          }
          e.t_pc(t) = 4;
          e.yb = ST_OK, e.yc = reply_for(e.rv0, e.t_b(t));
          return RQ_CH_SEND;
        case K_RSP_BIDI * 32 + 4:
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 2;
          continue;
        case K_RSP_BIDI * 32 + 5:
          return RQ_READY;

        // ---- tonic_bidi_sender: the background task that owns the Sender (client.rs:152-155) ----
        case K_SENDER * 32 + 0:
          e.t_c(t) = 0;
          e.t_pc(t) = 1;
          e.yb = 0, e.yc = e.t_a(t) | CLIENT_STREAMING | SERVER_STREAMING;
          return RQ_CH_SEND;
        case K_SENDER * 32 + 1:
          if (e.ret < 0) return RQ_READY;
          e.t_pc(t) = 2;
          // fallthrough
        case K_SENDER * 32 + 2:
          if (e.t_c(t) >= 3) return RQ_READY;  // the Sender drops with the future
          e.t_pc(t) = 3;
          e.yb = NAME + e.t_c(t), e.yc = 0;
          return RQ_CH_SEND;
        case K_SENDER * 32 + 3:
          if (e.ret < 0) return RQ_READY;
          e.t_c(t)++;
          e.t_pc(t) = 2;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;

        // ---- tonic_client: the client node's init task ----
        case K_CLIENT * 32 + 0:  // GreeterClient::connect: connect_ep, then drop it
          e.wl_w(F_STARTS, e.t_node(t))++;
          // fallthrough
        case K_CLIENT * 32 + 18:
          e.t_pc(t) = 1;
          e.ya = mk_addr(IPK_UNSPEC, 0, 0);
          return RQ_BIND;
        case K_CLIENT * 32 + 1:
          if (e.ret < 0) return call_failed(e, t, 70);
          e.t_pc(t) = 2;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        case K_CLIENT * 32 + 2:
          if (e.ret < 0) return call_failed(e, t, 70);
          e.conn_drop_end(t);
          e.sock_unref(t);
          if (INTERCEPT) e.wl_w(F_ICOUNT, e.t_node(t)) = 0;  // a (re)started node builds a new client: its closure starts at 0
          e.t_b(t) = 0;  // seq
          e.t_pc(t) = 3;
          // fallthrough
        case K_CLIENT * 32 + 3:  // loop { name = gen_range(0..10) == 0 ? "error" : "Tonic"
          e.t_pc(t) = 4;
          e.xa = 0, e.xb = 10;
          return RQ_GEN_RANGE;
        case K_CLIENT * 32 + 4:  // Grpc::unary / server_streaming -> connect_ep: Endpoint::connect
          e.t_a(t) = e.ret64 == 0 ? (u32)ERROR_NAME : (u32)NAME;
          e.t_c(t) = (STREAMING && (P.p[2] & FLAG_STREAMING_MIX) && e.ret64 >= 1 && e.ret64 <= 6) ? (u32)e.ret64 : 0u;  // 1..3: lots_of_replies, 4: lots_of_greetings, 5..6: bidi_hello
          if (INTERCEPT && (P.p[2] & FLAG_INTERCEPT)) {
            u32 node = e.t_node(t);
            if (++e.wl_w(F_ICOUNT, node) % 2 == 0) {  // request.intercept(..)? fails before connect1 (client.rs:65)
              e.wl_w(F_REPLIES, node)++;
              e.t_b(t)++;
              e.t_pc(t) = 3;
              e.xa = P.p[1];
              return RQ_SLEEP_FOR;
            }
            if (e.ret64 == 9) e.t_a(t) = UNKNOWN_SERVICE;
          }
          if (Caps::USES_TIMEOUT && (P.p[2] & FLAG_TIMEOUTS) && e.t_c(t) == 0) {  // with_timeout (client.rs:246-258): unary calls only
            u64 tmo = NS_PER_SEC;  // Endpoint::timeout(1 s)
            if (e.ret64 == 7 || e.ret64 == 8) e.t_a(t) = DELAY_NAME;
            if (e.ret64 == 8) tmo = 5 * NS_PER_SEC;  // Request::set_timeout overrides the channel's
            e.t_dl2(t) = e.sleep_deadline(e.now + tmo);  // the Sleep is created now (time/mod.rs:128-140)
          }
          e.t_pc(t) = 5;
          e.ya = mk_addr(IPK_UNSPEC, 0, 0);
          return RQ_BIND;
        case K_CLIENT * 32 + 5:  // the handshake connection
          if (e.ret < 0) return call_failed(e, t, 71);
          e.t_pc(t) = 6;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        case K_CLIENT * 32 + 6:  // the request's connection
          if (e.ret < 0) return call_failed(e, t, 71);
          e.conn_drop_end(t);
          e.t_pc(t) = 7;
          e.ya = SERVER_ADDR;
          return RQ_CONNECT1;
        case K_CLIENT * 32 + 7:
          if (e.ret < 0) return call_failed(e, t, 71);
          if (STREAMING && e.t_c(t) >= 5) {  // bidi: spawn the task that sends the requests; it takes the Sender
            e.t_pc(t) = 15;
            e.ya = K_SENDER | SPAWN_CONN_TX | SPAWN_JOINABLE, e.yb = e.t_node(t), e.yc = e.t_b(t), e.yd = 0, e.xa = e.xb = 0;
            return RQ_SPAWN;
          }
          if (STREAMING && e.t_c(t) == 4) {  // client streaming: the stream-start header, then hello_stream's items
            e.t_c(t) = 32;      // 32 + items sent
            e.t_a(t) = 0;       // fold of the names sent
            e.t_pc(t) = 11;
            e.yb = 0, e.yc = e.t_b(t) | CLIENT_STREAMING;
            return RQ_CH_SEND;
          }
          e.t_pc(t) = 8;
          e.yb = e.t_a(t), e.yc = e.t_b(t) | ((STREAMING && e.t_c(t)) ? (u32)SERVER_STREAMING : 0u);
          return RQ_CH_SEND;
        case K_CLIENT * 32 + 8:
          if (e.ret < 0) return call_failed(e, t, 71);
          e.t_pc(t) = 9;
          return RQ_CH_RECV;
        case K_CLIENT * 32 + 9: {
          if (e.ret < 0) return call_failed(e, t, 71);
          u32 node = e.t_node(t);
          u32 st = e.rv0, reply = e.rv1;
          if (STREAMING && e.t_c(t)) {  // the header of a server-streaming call
            if (st != ST_OK) return panic(e, 75);
            if (e.t_c(t) == 3) {  // drop the response stream at once (tests/test.rs:204-230)
              e.conn_drop_end(t);
              e.sock_unref(t);
              e.wl_w(F_REPLIES, node)++;
              e.t_b(t)++;
              e.t_pc(t) = 3;
              e.xa = P.p[1];
              return RQ_SLEEP_FOR;
            }
            e.t_c(t) = 16;  // items read so far = t_c - 16
            e.t_pc(t) = 10;
            return RQ_CH_RECV;
          }
          e.conn_drop_end(t);  // rx, tx, then the last BindGuard reference
          e.sock_unref(t);
          if (Caps::USES_TIMEOUT) e.t_dl2(t) = E::NONE64;  // the call completed inside its timeout
          if (e.t_a(t) == DELAY_NAME) return panic(e, 80);  // delay must end in DeadlineExceeded
          if (INTERCEPT && e.t_a(t) == UNKNOWN_SERVICE) {
            if (st != ST_UNIMPLEMENTED) return panic(e, 93);
          } else if (INTERCEPT && (P.p[2] & FLAG_INTERCEPT) && st == ST_UNAVAILABLE) {
            // the server's interceptor turned this one down
          } else
          if (e.t_a(t) == ERROR_NAME) {
            if (st != ST_INVALID_ARGUMENT) return panic(e, 72);
          } else if (st != ST_OK || reply != NAME * 31u + node) {
            return panic(e, 73);
          }
          e.wl_w(F_REPLIES, node)++;
          e.t_b(t)++;
          e.t_pc(t) = 3;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        }
        case K_CLIENT * 32 + 10: {  // Streaming::message(): three replies, then the EOF status
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) return call_failed(e, t, 71);
          u32 node = e.t_node(t), i = e.t_c(t) - 16;
          if (i < 3) {
            if (e.rv0 != ST_OK || e.rv1 != NAME * 31u + node + i) return panic(e, 76);
            e.t_c(t)++;
            return RQ_CH_RECV;
          }
          if (e.rv0 != ST_EOF) return panic(e, 77);
          e.conn_drop_end(t);
          e.sock_unref(t);
          e.wl_w(F_REPLIES, node)++;
          e.t_b(t)++;
          e.t_pc(t) = 3;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        }
        case K_CLIENT * 32 + 11:  // header sent (tx.send(..).await?), then: yield item i; sleep(1 s)
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) return call_failed(e, t, 71);
          // fallthrough
        case K_CLIENT * 32 + 12: {
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          u32 i = e.t_c(t) - 32;
          if (i >= 3) {  // the request stream ended: send_request_stream returns, its Sender drops
            e.conn_drop_tx(t);
            e.t_pc(t) = 14;
            return RQ_CH_RECV;
          }
          e.t_a(t) = e.t_a(t) * 33u + (NAME + i);
          e.t_pc(t) = 13;
          e.yb = NAME + i, e.yc = 0;
          return RQ_CH_SEND;
        }
        case K_CLIENT * 32 + 13:
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) {  // the server closed the stream prematurely: stop sending
            e.t_c(t) = 35;
            e.t_pc(t) = 12;
            continue;
          }
          e.t_c(t)++;
          e.t_pc(t) = 12;
          e.xa = NS_PER_SEC;
          return RQ_SLEEP_FOR;
        case K_CLIENT * 32 + 14: {
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) return call_failed(e, t, 71);
          u32 node = e.t_node(t);
          if (e.rv0 != ST_OK || e.rv1 != e.t_a(t) * 31u + node) return panic(e, 78);
          e.conn_drop_end(t);
          e.sock_unref(t);
          e.wl_w(F_REPLIES, node)++;
          e.t_b(t)++;
          e.t_pc(t) = 3;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        }
        case K_CLIENT * 32 + 15:  // a plain JoinHandle until the response stream is first polled
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          e.t_handle(t) = (u8)((u32)e.ret | 0x80u);
          e.t_a(t) = e.t_c(t);  // 5, or 6 = drop the stream after the first reply
          e.t_c(t) = 0;         // replies read
          e.t_pc(t) = 16;
          return RQ_CH_RECV;
        case K_CLIENT * 32 + 16:  // the header: Result<Response<()>>
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) return call_failed(e, t, 71);
          if (e.rv0 != ST_OK) return panic(e, 75);
          e.t_handle(t) = e.t_handle(t) & 0x7f;  // first poll of the response stream: `_task = task.cancel_on_drop()`
          e.t_pc(t) = 17;
          return RQ_CH_RECV;
        case K_CLIENT * 32 + 17: {  // while let Some(reply) = stream.message().await.unwrap()
          if (!STREAMING) return panic(e, 0xbad);  // streaming calls only
          if (e.ret < 0) return call_failed(e, t, 71);
          u32 node = e.t_node(t);
          bool done = false;
          if (e.rv0 == ST_TRAILER) {
            if (e.t_c(t) != 3) return panic(e, 79);
            done = true;
          } else {
            if (e.rv0 != ST_OK || e.rv1 != (NAME + e.t_c(t)) * 31u + node) return panic(e, 79);
            e.t_c(t)++;
            if (e.t_a(t) == 6) done = true;  // drop the response stream after the first reply
          }
          if (!done) return RQ_CH_RECV;
          e.task_cancel_detach(e.t_handle(t));  // `_task` drops (a no-op on a finished task) before the Receiver
          e.t_handle(t) = NONE8;
          e.conn_drop_end(t);
          e.sock_unref(t);
          e.wl_w(F_REPLIES, node)++;
          e.t_b(t)++;
          e.t_pc(t) = 3;
          e.xa = P.p[1];
          return RQ_SLEEP_FOR;
        }
        default:
          return panic(e, 0xbad);
      }
  }
};
using WL_TonicUnary = WL_TonicT<false, false>;  // flags without 2, 8, 16: BASELINE config 5
using WL_TonicIntercept = WL_TonicT<false, false, true>;  // + interceptors and an unknown service (flag 16; excludes 2 and 8)
using WL_Tonic = WL_TonicT<false, true>;        // + streaming calls (flag 2)
using WL_TonicBig = WL_TonicT<true, true>;      // + request timeouts (flag 8)
/// the instantiation that runs a given flags word (p[2]): f(WL{}) with the chosen program type
template <class F>
inline auto tonic_variant(u64 flags, F&& f) {
  if (flags & 8) return f(WL_TonicBig{});
  if (flags & 2) return f(WL_Tonic{});
  if (flags & 16) return f(WL_TonicIntercept{});
  return f(WL_TonicUnary{});
}

}  // namespace msim
