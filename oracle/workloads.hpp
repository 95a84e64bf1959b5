// ORACLE — TEST INFRASTRUCTURE ONLY (see rt/future.hpp).
//
// The registered workloads (include/msim_b200.h: MSIM_WL_*), written the way a madsim test
// is written: async code against Endpoint / spawn / sleep / call_timeout.  The device
// library implements the SAME programs as explicit poll-level state machines
// (madsim_b200/csrc/workloads/*.cuh); parity = equal per-seed trace hash.
//
// MadRaft itself is not in /root/reference (README.md:78 only names it):
// the Raft election program below is authored here on the reference's API.
#pragma once
#include <array>

#include "rt/runtime.hpp"

namespace msim_oracle {

// ------------------------------------------------------------------------------------
// MSIM_WL_SLEEP — SURVEY Appendix B: block_on(async { sleep(d).await })
inline Async<u32> wl_sleep_main(u64 dur_ns) {
  co_await sleep(dur_ns);
  co_return 0;
}

// ------------------------------------------------------------------------------------
// MSIM_WL_PINGPONG — BASELINE config 1; endpoint.rs:550-584 (connect_send_recv) with the
// un-simulated tokio Barrier replaced by a sleep, repeated R times.
constexpr u32 PP_PING = 0x676e6970u;  // "ping"
constexpr u32 PP_PONG = 0x676e6f70u;  // "pong"
inline SocketAddr pp_addr(u32 i) { return SocketAddr{ip4(10, 0, 0, i), 1}; }

inline Async<Unit> pp_server(SocketAddr addr) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  if (ep.local_addr() != addr) panic("local_addr mismatch", 10);
  for (;;) {
    auto [data, from] = (co_await ep.recv_from(1)).unwrap();
    if (data.v[0] != PP_PING) panic("expected ping", 11);
    Payload pong;
    pong.v[0] = PP_PONG;
    (co_await ep.send_to(from, 1, pong)).unwrap();
  }
}
inline Async<u32> pp_client(SocketAddr server, u64 rounds, u64 think_ns) {
  co_await sleep(1 * SEC);
  Endpoint ep = std::move((co_await Endpoint::connect(server)).unwrap());
  if (*ep.peer_addr().ok != server) panic("peer_addr mismatch", 12);
  u32 done = 0;
  for (u64 r = 0; r < rounds; r++) {
    Payload ping;
    ping.v[0] = PP_PING;
    (co_await ep.send(1, ping)).unwrap();
    Payload p = (co_await ep.recv(1)).unwrap();
    if (p.v[0] != PP_PONG) panic("expected pong", 13);
    done++;
    co_await sleep(think_ns);
  }
  co_return done;
}
inline Async<u32> wl_pingpong_main(u64 rounds, u64 think_ns) {
  auto& rt = current_runtime();
  auto node1 = rt.create_node().ip(pp_addr(1).ip).build();
  auto node2 = rt.create_node().ip(pp_addr(2).ip).build();
  node1.spawn(pp_server(pp_addr(1)));
  auto f = node2.spawn(pp_client(pp_addr(1), rounds, think_ns));
  auto r = co_await f;
  co_return r.unwrap();
}

// ------------------------------------------------------------------------------------
// MSIM_WL_RAFT — BASELINE configs 2 and 3: 5-node Raft leader election (terms, votes,
// heartbeats) + optional supervisor fault schedule (partitions, clock jumps).
struct RaftParams {
  u32 n = 5;
  u64 dur_ms = 10000;
  u32 fault_mask = 0;
  u32 bug = 0;
  u64 rpc_timeout_ms = 40;
  u64 hb_ms = 50;
  u64 el_lo_ms = 150, el_hi_ms = 300;
};
enum : u32 { RAFT_FOLLOWER = 0, RAFT_CANDIDATE = 1, RAFT_LEADER = 2 };
struct Raft {
  u32 term = 0;
  u32 voted_for = 0;  // node id, 0 = none
  u32 role = RAFT_FOLLOWER;
  u32 votes = 0;
  u32 leader = 0;
  u64 last_hb = 0;  // Instant of the last heartbeat / granted vote / election start
};
struct RaftCluster {
  RaftParams p;
  std::array<std::shared_ptr<Raft>, 9> rafts{};  // by node id
  u32 elections = 0;
};
constexpr u64 RAFT_TAG_VOTE = hash_str("raft::RequestVote");
constexpr u64 RAFT_TAG_APPEND = hash_str("raft::AppendEntries");
inline SocketAddr raft_addr(u32 i) { return SocketAddr{ip4(10, 0, 0, i), 1}; }

inline Async<Payload> raft_on_request_vote(std::shared_ptr<Raft> r, std::shared_ptr<RaftCluster> c, Payload req) {
  u32 term = req.v[0], cand = req.v[1];
  if (term > r->term) {
    r->term = term;
    r->role = RAFT_FOLLOWER;
    r->voted_for = 0;
  }
  bool granted = term == r->term && (c->p.bug == 1 || r->voted_for == 0 || r->voted_for == cand);
  if (granted) {
    r->voted_for = cand;
    r->last_hb = now();
  }
  Payload rsp;
  rsp.v[0] = r->term;
  rsp.v[1] = granted ? 1 : 0;
  co_return rsp;
}
inline Async<Payload> raft_on_append(std::shared_ptr<Raft> r, std::shared_ptr<RaftCluster>, Payload req) {
  u32 term = req.v[0], leader = req.v[1];
  bool ok = term >= r->term;
  if (ok) {
    if (term > r->term) {
      r->term = term;
      r->voted_for = 0;
    }
    r->role = RAFT_FOLLOWER;
    r->leader = leader;
    r->last_hb = now();
  }
  Payload rsp;
  rsp.v[0] = r->term;
  rsp.v[1] = ok ? 1 : 0;
  co_return rsp;
}
inline Async<Unit> raft_append_task(Endpoint ep, std::shared_ptr<Raft> r, std::shared_ptr<RaftCluster> c, u32 me,
                                    u32 peer, u32 term0) {
  Payload req;
  req.v[0] = term0;
  req.v[1] = me;
  auto rsp = co_await ep.call_timeout(raft_addr(peer), RAFT_TAG_APPEND, req, c->p.rpc_timeout_ms * MS);
  if (rsp.is_ok() && rsp.ok->v[0] > r->term) {
    r->term = rsp.ok->v[0];
    r->role = RAFT_FOLLOWER;
    r->voted_for = 0;
  }
  co_return {};
}
inline Async<Unit> raft_leader_loop(Endpoint ep, std::shared_ptr<Raft> r, std::shared_ptr<RaftCluster> c, u32 me,
                                    u32 term0) {
  while (r->role == RAFT_LEADER && r->term == term0) {
    for (u32 peer = 1; peer <= c->p.n; peer++)
      if (peer != me) spawn(raft_append_task(ep, r, c, me, peer, term0));
    co_await sleep(c->p.hb_ms * MS);
  }
  co_return {};
}
inline Async<Unit> raft_vote_task(Endpoint ep, std::shared_ptr<Raft> r, std::shared_ptr<RaftCluster> c, u32 me,
                                  u32 peer, u32 term0) {
  Payload req;
  req.v[0] = term0;
  req.v[1] = me;
  auto rsp = co_await ep.call_timeout(raft_addr(peer), RAFT_TAG_VOTE, req, c->p.rpc_timeout_ms * MS);
  if (rsp.is_ok()) {
    u32 t = rsp.ok->v[0];
    bool granted = rsp.ok->v[1] != 0;
    if (t > r->term) {
      r->term = t;
      r->role = RAFT_FOLLOWER;
      r->voted_for = 0;
    } else if (r->role == RAFT_CANDIDATE && r->term == term0 && granted) {
      r->votes++;
      if (r->votes == c->p.n / 2 + 1) {
        r->role = RAFT_LEADER;
        r->leader = me;
        spawn(raft_leader_loop(ep, r, c, me, term0));
      }
    }
  }
  co_return {};
}
/// The node's init task (NodeBuilder::init): bind, serve RPCs, run the election ticker.
inline Async<Unit> raft_node_main(std::shared_ptr<RaftCluster> c, u32 me) {
  Endpoint ep = std::move((co_await Endpoint::bind(raft_addr(me))).unwrap());
  auto r = std::make_shared<Raft>();
  r->last_hb = now();
  c->rafts[me] = r;
  ep.add_rpc_handler(RAFT_TAG_VOTE, [r, c](Payload q) { return raft_on_request_vote(r, c, q); });
  ep.add_rpc_handler(RAFT_TAG_APPEND, [r, c](Payload q) { return raft_on_append(r, c, q); });
  for (;;) {
    u64 to_ms = thread_rng().gen_range_u64(c->p.el_lo_ms, c->p.el_hi_ms);
    co_await sleep(to_ms * MS);
    if (r->role != RAFT_LEADER && now() - r->last_hb >= to_ms * MS) {
      r->term++;
      r->role = RAFT_CANDIDATE;
      r->voted_for = me;
      r->votes = 1;
      r->last_hb = now();
      c->elections++;
      u32 term0 = r->term;
      for (u32 peer = 1; peer <= c->p.n; peer++)
        if (peer != me) spawn(raft_vote_task(ep, r, c, me, peer, term0));
    }
  }
}
/// Supervisor fault schedule (config 3): partitions via NetSim::clog_* (net/mod.rs:164-222),
/// clock jumps via time::advance (time/mod.rs:195-198); every choice is drawn from the
/// seed's own RNG.
inline Async<Unit> raft_fault_task(std::shared_ptr<RaftCluster> c) {
  GlobalRng& rng = thread_rng();
  NetSim& net = current_net();
  u32 kinds[3], nk = 0;
  for (u32 k = 0; k < 3; k++)
    if (c->p.fault_mask & (1u << k)) kinds[nk++] = k;
  for (;;) {
    co_await sleep(rng.gen_range_u64(500, 2000) * MS);
    u32 k = kinds[rng.gen_range_u64(0, nk)];
    if (k == 0) {
      u64 i = rng.gen_range_u64(1, c->p.n + 1);
      net.clog_node(i);
      co_await sleep(rng.gen_range_u64(200, 1000) * MS);
      net.unclog_node(i);
    } else if (k == 1) {
      u64 i = rng.gen_range_u64(1, c->p.n + 1);
      u64 j = rng.gen_range_u64(1, c->p.n);
      if (j >= i) j++;
      net.clog_link(i, j);
      net.clog_link(j, i);
      co_await sleep(rng.gen_range_u64(200, 1000) * MS);
      net.unclog_link(i, j);
      net.unclog_link(j, i);
    } else {
      advance(rng.gen_range_u64(0, 500) * MS);
    }
  }
}
struct RaftOutcome {
  u32 leader = 0, term = 0, elections = 0, msgs = 0;
};
inline Async<u32> wl_raft_main(std::shared_ptr<RaftCluster> c, RaftOutcome* out) {
  auto& rt = current_runtime();
  for (u32 i = 1; i <= c->p.n; i++)
    rt.create_node().ip(raft_addr(i).ip).init([c, i]() { return raft_node_main(c, i); }).build();
  if (c->p.fault_mask) spawn(raft_fault_task(c));
  u64 ticks = c->p.dur_ms / 100;
  for (u64 tick = 1; tick <= ticks; tick++) {
    co_await sleep(100 * MS);
    // safety: at most one leader per term
    bool any_leader = false;
    for (u32 a = 1; a <= c->p.n; a++) {
      if (!c->rafts[a] || c->rafts[a]->role != RAFT_LEADER) continue;
      any_leader = true;
      for (u32 b = a + 1; b <= c->p.n; b++)
        if (c->rafts[b] && c->rafts[b]->role == RAFT_LEADER && c->rafts[b]->term == c->rafts[a]->term)
          panic("two leaders in one term", 1);
    }
    // liveness (fault-free runs only): a leader exists by 2 s
    if (c->p.fault_mask == 0 && tick == 20 && !any_leader) panic("no leader elected by 2 s", 2);
  }
  for (u32 a = 1; a <= c->p.n; a++) {
    if (!c->rafts[a]) continue;
    out->term = std::max(out->term, c->rafts[a]->term);
  }
  u32 best = 0;
  for (u32 a = 1; a <= c->p.n; a++)
    if (c->rafts[a] && c->rafts[a]->role == RAFT_LEADER && c->rafts[a]->term >= best) {
      best = c->rafts[a]->term;
      out->leader = a;
    }
  out->elections = c->elections;
  out->msgs = u32(current_net().msg_count());
  co_return 0;
}

// ------------------------------------------------------------------------------------
// MSIM_WL_CHAOS — the fault surface of Handle (task/mod.rs:852-1016 kill / restart /
// pause_resume tests folded into one schedule): a worker node counts 2-second ticks while the
// supervisor kills, restarts, pauses and resumes it `rounds` times.
struct ChaosState {
  u32 flag = 0, starts = 0;
};
inline Async<Unit> chaos_worker(std::shared_ptr<ChaosState> st, u64 period_ns, u32 panics, u32 matching) {
  st->starts++;
  if (st->starts <= panics) {  // restart_on_panic (task/mod.rs:939-963) / _matching (:965-983): panic!("{}", starts - 1)
    if (matching) panic(std::to_string(st->starts - 1), 100 + st->starts - 1);
    panic("chaos worker boom", 23);
  }
  for (;;) {
    co_await sleep(period_ns);
    st->flag += 2;
  }
}
inline Async<u32> wl_chaos_main(u64 rounds, u64 period_ms, u32 panics, std::array<u32, 4>* user, u32 matching = 0) {
  auto& rt = current_runtime();
  auto st = std::make_shared<ChaosState>();
  u64 period = period_ms * MS;
  auto nb = rt.create_node();
  nb.init([st, period, panics, matching]() { return chaos_worker(st, period, panics, matching); });
  if (matching) {  // .restart_on_panic_matching("0") .. ("matching - 1"): single digits, so `contains` is equality
    for (u32 i = 0; i < matching; i++) nb.restart_on_panic_matching(std::to_string(i));
  } else if (panics) {
    nb.restart_on_panic();
  }
  auto node = nb.build();
  u64 id = node.id();
  GlobalRng& rng = thread_rng();
  for (u64 r = 0; r < rounds; r++) {
    co_await sleep(rng.gen_range_u64(1, 4) * period + period / 2);
    u64 k = rng.gen_range_u64(0, 3);
    if (k == 0) {
      rt.kill(id);
      if (!rt.is_exit(id)) panic("killed node must be exited", 20);
      co_await sleep(period);
      rt.restart(id);
      if (rt.is_exit(id)) panic("restarted node must be running", 21);
    } else if (k == 1) {
      rt.pause(id);
      u32 before = st->flag;
      co_await sleep(3 * period);
      if (st->flag != before) panic("paused node made progress", 22);
      rt.resume(id);
    } else {
      rt.restart(id);
    }
  }
  co_await sleep(period);
  (*user)[0] = st->flag;
  (*user)[1] = st->starts;
  co_return 0;
}

// ------------------------------------------------------------------------------------
// MSIM_WL_STREAM — SURVEY §8(f) row 1: the reliable channel layer every ecosystem simulator is
// built on (connect1 / accept1 / Sender / Receiver, net/mod.rs:337-430, endpoint.rs:178-274):
// an echo server and a client exchanging p[0] messages over ONE ordered connection while a
// supervisor task partitions the two nodes (in the spirit of tcp/mod.rs:104-179
// disconnect_and_recovery): lost packets become DELAY (exponential back-off), never loss.
constexpr u32 STREAM_FAULTS = 1;
inline Async<Unit> stream_echo(Connection conn) {
  for (;;) {
    auto r = co_await conn.rx.recv();
    if (r.is_err()) break;  // ConnectionReset: the peer dropped its sender
    Payload v = *r.ok;
    v.v[1] = v.v[0] ^ 0x5a5a5a5au;
    if (conn.tx.send(v).is_err()) break;
  }
  co_return {};
}
inline Async<Unit> stream_server(SocketAddr addr) {
  Endpoint ep = std::move((co_await Endpoint::bind(addr)).unwrap());
  for (;;) {
    Connection conn = std::move((co_await ep.accept1()).unwrap());
    spawn(stream_echo(std::move(conn)));
  }
}
inline Async<u32> stream_client(SocketAddr server, u64 msgs, u64 think_ns) {
  co_await sleep(1 * SEC);
  Endpoint ep = std::move((co_await Endpoint::bind(SocketAddr{IP_UNSPECIFIED, 0})).unwrap());
  u32 attempts = 0;
  std::optional<Connection> conn;
  for (;;) {  // connect1 surfaces ConnectionRefused when the request is lost / clogged: retry
    attempts++;
    auto r = co_await ep.connect1(server);
    if (r.is_ok()) {
      conn.emplace(std::move(*r.ok));
      break;
    }
    if (r.err != Err::ConnectionRefused) panic("unexpected connect1 error", 40);
    co_await sleep(100 * MS);
  }
  for (u64 i = 0; i < msgs; i++) {
    Payload p;
    p.v[0] = u32(i);
    conn->tx.send(p).unwrap();
    Payload r = (co_await conn->rx.recv()).unwrap();
    if (r.v[0] != u32(i) || r.v[1] != (u32(i) ^ 0x5a5a5a5au)) panic("echo mismatch / reordering", 41);
    co_await sleep(think_ns);
  }
  co_return attempts;
}
inline Async<Unit> stream_fault_task() {
  GlobalRng& rng = thread_rng();
  NetSim& net = current_net();
  for (;;) {
    co_await sleep(rng.gen_range_u64(300, 1500) * MS);
    u64 k = rng.gen_range_u64(0, 3);
    if (k == 0) {
      net.clog_link(1, 2);
      net.clog_link(2, 1);
    } else {
      net.clog_node(k);  // node 1 or 2
    }
    co_await sleep(rng.gen_range_u64(100, 3000) * MS);
    if (k == 0) {
      net.unclog_link(1, 2);
      net.unclog_link(2, 1);
    } else {
      net.unclog_node(k);
    }
  }
}
inline Async<u32> wl_stream_main(u64 msgs, u64 think_ns, u32 faults, std::array<u32, 4>* user) {
  auto& rt = current_runtime();
  SocketAddr a1{ip4(10, 0, 0, 1), 1};
  auto node1 = rt.create_node().ip(a1.ip).build();
  auto node2 = rt.create_node().ip(ip4(10, 0, 0, 2)).build();
  node1.spawn(stream_server(a1));
  auto f = node2.spawn(stream_client(a1, msgs, think_ns));
  if (faults & STREAM_FAULTS) spawn(stream_fault_task());
  auto r = co_await f;
  (*user)[0] = u32(msgs);
  (*user)[1] = r.unwrap();
  (*user)[3] = u32(current_net().msg_count());
  co_return 0;
}

}  // namespace msim_oracle

#include "workloads_etcd.hpp"
#include "workloads_tonic.hpp"
#include "workloads_signal.hpp"
#include "workloads_election.hpp"
