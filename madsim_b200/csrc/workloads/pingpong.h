// MSIM_WL_PINGPONG — BASELINE config 1: 2-node ping-pong over net::Endpoint + time::sleep
// (reference test net/endpoint.rs:550-584 `connect_send_recv`, repeated p[0] times with the
// un-simulated tokio Barrier replaced by a 1 s sleep).  Oracle twin: oracle/workloads.hpp
// pp_server / pp_client / wl_pingpong_main — each `case` below is one await point of that
// async code, expanded to the poll-level behaviour of SURVEY.md §3.4.
#pragma once
#include "../engine/engine.h"

namespace msim {

struct WL_PingPong {
  struct Caps {
    static constexpr int NT = 4, TQ = 16, NMSG = 4, NSOCK = 2, NREG = 2, NMBOX = 2, NNODE = 3, NPARK = 1;
    static constexpr int WQ = 0, WW = 0, WS = 0;
  };
  enum : u32 { K_MAIN = 0, K_SERVER = 1, K_CLIENT = 2 };
  static constexpr u32 MAIN_KIND = K_MAIN;
  static constexpr u32 PING = 0x676e6970u, PONG = 0x676e6f70u;
  static constexpr u32 SERVER_ADDR = (IPK_NODE << 24) | (1u << 16) | 1u;  // 10.0.0.1:1

  template <class E>
  static MSIM_HD int poll_main(E& e, int t) {
    using L = typename E::L;
    switch (e.t_pc(t)) {
      case 0: {
        int n1 = e.create_node(true, false);
        int n2 = e.create_node(true, false);
        e.spawn(K_SERVER, n1);                      // node1.spawn(server) — handle dropped (detached)
        int c = e.spawn(K_CLIENT, n2, TF_JOINABLE);  // let f = node2.spawn(client)
        if (c < 0) return POLL_PENDING;
        e.t_a(t) = (u32)c;
        e.t_pc(t) = 1;
      }
      // fallthrough
      case 1: {
        int target = (int)e.t_a(t);
        int r = e.join_poll(t, target);  // f.await
        if (r == 0) return POLL_PENDING;
        if (r == 2) return e.panic(14);  // .unwrap() on JoinError
        e.w(L::W_USER0) = e.t_a(target);
        e.join_release(target);
        return POLL_READY;
      }
    }
    return e.panic(0xbad);
  }

  template <class E>
  static MSIM_HD int poll_server(E& e, int t) {
    using L = typename E::L;
    for (;;) switch (e.t_pc(t)) {
        case 0:  // Endpoint::bind(addr): BindGuard::bind → rand_delay
          e.t_dl0(t) = e.rand_delay_deadline();
          e.t_pc(t) = 1;
          // fallthrough
        case 1: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int k = e.sock_bind(e.t_node(t), SERVER_ADDR);
          if (k < 0) return e.panic(10);
          e.sock_adopt(t, k);
          e.t_pc(t) = 2;
        }
        // fallthrough
        case 2:  // loop { recv_from(1): mailbox.recv(tag)
          e.mbox_recv(t, e.t_sock(t), 1);
          e.t_pc(t) = 3;
          // fallthrough
        case 3:  // oneshot::Receiver await
          if (e.t_rx(t) == NONE8) return POLL_PENDING;
          e.t_dl0(t) = e.rand_delay_deadline();  // rand_delay AFTER the receipt
          e.t_pc(t) = 4;
          // fallthrough
        case 4: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int m = e.t_rx(t);
          u32 from = e.w(L::W_M_FROM, m), v0 = e.w(L::W_M_V0, m);
          e.msg_free(m);
          e.t_rx(t) = NONE8;
          if (v0 != PING) return e.panic(11);
          e.t_a(t) = from;
          e.t_dl0(t) = e.rand_delay_deadline();  // send_to(from, 1, "pong"): rand_delay
          e.t_pc(t) = 5;
        }
        // fallthrough
        case 5: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int m = e.msg_alloc();
          if (m < 0) return POLL_PENDING;
          e.q(L::Q_M_TAG, m) = 1;
          e.q(L::Q_M_W, m) = 0;
          e.w(L::W_M_V0, m) = PONG;
          e.w(L::W_M_V1, m) = 0;
          int k = e.t_sock(t);
          e.net_send(e.b(L::B_SK_NODE, k), e.sk_port(k), e.t_a(t), m);
          e.t_pc(t) = 2;  // }
          continue;
        }
        default:
          return e.panic(0xbad);
      }
  }

  template <class E>
  static MSIM_HD int poll_client(E& e, int t) {
    using L = typename E::L;
    for (;;) switch (e.t_pc(t)) {
        case 0:  // sleep(1 s)
          e.t_dl0(t) = e.sleep_deadline(e.now + NS_PER_SEC);
          e.t_pc(t) = 1;
          // fallthrough
        case 1:
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          e.t_dl0(t) = e.rand_delay_deadline();  // Endpoint::connect = bind("0.0.0.0:0")
          e.t_pc(t) = 2;
          // fallthrough
        case 2: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int k = e.sock_bind(e.t_node(t), mk_addr(IPK_UNSPEC, 0, 0));
          if (k < 0) return e.panic(12);
          e.sock_adopt(t, k);
          e.t_a(t) = 0;  // done
          e.t_b(t) = 0;  // r
          e.t_pc(t) = 3;
        }
        // fallthrough
        case 3:  // for r in 0..rounds {
          if ((u64)e.t_b(t) >= e.P->p[0]) return POLL_READY;  // output = t_a; the Endpoint drops with the future
          e.t_dl0(t) = e.rand_delay_deadline();  // ep.send(1, "ping"): rand_delay
          e.t_pc(t) = 4;
          // fallthrough
        case 4: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int m = e.msg_alloc();
          if (m < 0) return POLL_PENDING;
          e.q(L::Q_M_TAG, m) = 1;
          e.q(L::Q_M_W, m) = 0;
          e.w(L::W_M_V0, m) = PING;
          e.w(L::W_M_V1, m) = 0;
          int k = e.t_sock(t);
          e.net_send(e.b(L::B_SK_NODE, k), e.sk_port(k), SERVER_ADDR, m);
          e.mbox_recv(t, k, 1);  // ep.recv(1): mailbox.recv(tag)
          e.t_pc(t) = 5;
        }
        // fallthrough
        case 5:
          if (e.t_rx(t) == NONE8) return POLL_PENDING;
          e.t_dl0(t) = e.rand_delay_deadline();
          e.t_pc(t) = 6;
          // fallthrough
        case 6: {
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          int m = e.t_rx(t);
          u32 from = e.w(L::W_M_FROM, m), v0 = e.w(L::W_M_V0, m);
          e.msg_free(m);
          e.t_rx(t) = NONE8;
          if (from != SERVER_ADDR) return e.panic(15);  // "not from the connected address"
          if (v0 != PONG) return e.panic(13);
          e.t_a(t) = e.t_a(t) + 1;
          e.t_dl0(t) = e.sleep_deadline(e.now + e.P->p[1]);  // sleep(think)
          e.t_pc(t) = 7;
        }
        // fallthrough
        case 7:
          if (!e.sleep_poll(t, e.t_dl0(t))) return POLL_PENDING;
          e.t_b(t) = e.t_b(t) + 1;
          e.t_pc(t) = 3;  // }
          continue;
        default:
          return e.panic(0xbad);
      }
  }

  template <class E>
  static MSIM_HD int poll(E& e, int t) {
    switch (e.t_kind(t)) {
      case K_MAIN: return poll_main(e, t);
      case K_SERVER: return poll_server(e, t);
      default: return poll_client(e, t);
    }
  }
  template <class E>
  static MSIM_HD void on_drop(E& e, int t) {
    e.drop_task_resources(t, e.t_kind(t) != K_MAIN);
  }
  template <class E>
  static MSIM_HD void node_init(E&, int) {}
  template <class E>
  static MSIM_HD void finish(E&) {}
};

}  // namespace msim
