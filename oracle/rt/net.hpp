// ORACLE — TEST INFRASTRUCTURE ONLY (see future.hpp).
//
// Network simulator: restatement of madsim/src/sim/net/network.rs:20-313 (Network),
// net/mod.rs:82-333,436-494 (NetSim::{rand_delay,send}, BindGuard), net/endpoint.rs:11-175,
// 290-363 (Endpoint, Mailbox) and net/rpc.rs:82-166 (call / call_timeout / add_rpc_handler),
// plus tokio::sync::oneshot as used by the mailbox.
// Reliable channels: connect1 / accept1 / Sender / Receiver,
// net/mod.rs:337-430 + net/endpoint.rs:178-274, over tokio's unbounded mpsc and an async stream.
#pragma once
#include <vector>
#include <deque>
#include <map>
#include <set>

#include "task.hpp"

namespace msim_oracle {

struct Runtime;
Runtime& current_runtime();
std::shared_ptr<TaskInfo> current_task();

// ---- addresses (IPv4 only; addr.rs resolution is always immediately ready) ----
constexpr u32 ip4(u32 a, u32 b, u32 c, u32 d) { return (a << 24) | (b << 16) | (c << 8) | d; }
constexpr u32 IP_UNSPECIFIED = 0;
constexpr u32 IP_LOCALHOST = ip4(127, 0, 0, 1);
inline bool ip_is_loopback(u32 ip) { return (ip >> 24) == 127; }
struct SocketAddr {
  u32 ip = 0;
  u16 port = 0;
  auto operator<=>(const SocketAddr&) const = default;
};
enum class IpProtocol : u8 { Tcp, Udp };

/// `Payload = Box<dyn Any>`: the simulator never looks inside (net/mod.rs:95).  Workloads here
/// exchange a fixed POD, which is also what the device library carries.
struct Payload {
  u64 w = 0;  // rpc: the response tag travels in front of the request (rpc.rs:120-123)
  u32 v[4] = {0, 0, 0, 0};
};

struct NetConfig {  // network.rs:66-89
  double packet_loss_rate = 0.0;
  u64 send_latency_lo = 1000000ull;   // 1 ms
  u64 send_latency_hi = 10000000ull;  // 10 ms
};

// ---- tokio::sync::mpsc::unbounded_channel (what net/mod.rs:373 and async_channel's single
// consumer need): FIFO, recv wakes through one registered waker ----
template <class T>
struct MpscState {
  std::deque<T> q;
  std::optional<Waker> rx_waker;
  int tx_count = 0;
  bool rx_closed = false;
  // Sender::closed() waiters (tokio: notify_rx_closed.notify_waiters() when the receiver closes)
  std::vector<std::pair<u64, Waker>> closed_waiters;
  u64 next_closed_key = 1;
  /// The queued items drop FRONT TO BACK, like the list of the Rust queue (tokio mpsc / concurrent-queue walk head -> tail).
  /// std::deque's own destructor destroys the full middle buffers first, then the first and the last partial ones — with
  /// more than ~33 queued connections (16-byte elements, 512-byte buffers) that is a different wake order of their peers.
  ~MpscState() {
    while (!q.empty()) q.pop_front();
  }
  void close_rx() {
    rx_closed = true;
    auto ws = std::move(closed_waiters);
    closed_waiters.clear();
    for (auto& kv : ws) kv.second.wake();
  }
  void wake_rx() {
    if (rx_waker) {
      Waker w = std::move(*rx_waker);
      rx_waker.reset();
      w.wake();
    }
  }
};
template <class T>
struct UnboundedSender {
  std::shared_ptr<MpscState<T>> st;
  UnboundedSender() = default;
  explicit UnboundedSender(std::shared_ptr<MpscState<T>> s) : st(std::move(s)) {
    if (st) st->tx_count++;
  }
  UnboundedSender(const UnboundedSender& o) : st(o.st) {
    if (st) st->tx_count++;
  }
  UnboundedSender(UnboundedSender&& o) noexcept : st(std::move(o.st)) {}
  UnboundedSender& operator=(UnboundedSender o) {
    std::swap(st, o.st);
    return *this;
  }
  ~UnboundedSender() {
    if (st && --st->tx_count == 0) st->wake_rx();  // closed: the receiver sees None after draining
  }
  bool send(T v) const {  // Err if the receiver is gone
    if (st->rx_closed) return false;
    st->q.push_back(std::move(v));
    st->wake_rx();
    return true;
  }
  bool is_closed() const { return st->rx_closed; }
  /// `tx.closed()`: completes when the receiver is dropped
  struct ClosedFut {
    std::shared_ptr<MpscState<T>> st;
    u64 key = 0;
    explicit ClosedFut(std::shared_ptr<MpscState<T>> s) : st(std::move(s)) {}
    ClosedFut(ClosedFut&& o) noexcept : st(std::move(o.st)), key(std::exchange(o.key, 0)) {}
    ClosedFut(const ClosedFut&) = delete;
    ~ClosedFut() {
      if (!key || !st) return;
      for (size_t i = 0; i < st->closed_waiters.size(); i++)
        if (st->closed_waiters[i].first == key) {
          st->closed_waiters.erase(st->closed_waiters.begin() + long(i));
          break;
        }
    }
    Poll<Unit> poll(Context& cx) {
      if (st->rx_closed) return Poll<Unit>::done({});
      if (key) {
        for (auto& kv : st->closed_waiters)
          if (kv.first == key) kv.second = cx.waker;
      } else {
        key = st->next_closed_key++;
        st->closed_waiters.emplace_back(key, cx.waker);
      }
      return Poll<Unit>::pending();
    }
  };
  ClosedFut closed() const { return ClosedFut(st); }
};
template <class T>
struct UnboundedReceiver {
  std::shared_ptr<MpscState<T>> st;
  UnboundedReceiver() = default;
  explicit UnboundedReceiver(std::shared_ptr<MpscState<T>> s) : st(std::move(s)) {}
  UnboundedReceiver(UnboundedReceiver&&) = default;
  UnboundedReceiver& operator=(UnboundedReceiver&&) = default;
  ~UnboundedReceiver() {
    if (st) st->close_rx();
  }
  /// `rx.recv().await`
  struct RecvFut {
    UnboundedReceiver* rx;
    Poll<std::optional<T>> poll(Context& cx) {
      auto& s = *rx->st;
      if (!s.q.empty()) {
        T v = std::move(s.q.front());
        s.q.pop_front();
        return Poll<std::optional<T>>::done(std::optional<T>(std::move(v)));
      }
      if (s.tx_count == 0) return Poll<std::optional<T>>::done(std::nullopt);
      s.rx_waker = cx.waker;
      return Poll<std::optional<T>>::pending();
    }
  };
  RecvFut recv() { return RecvFut{this}; }
};
template <class T>
std::pair<UnboundedSender<T>, UnboundedReceiver<T>> unbounded_channel() {
  auto st = std::make_shared<MpscState<T>>();
  return {UnboundedSender<T>(st), UnboundedReceiver<T>(st)};
}

/// The link state when sending a packet (net/mod.rs:411): Some(arrive time) or None (lost / clogged)
using LinkState = std::optional<u64>;
/// net/mod.rs:404-430
struct PayloadSender {
  std::shared_ptr<std::function<LinkState()>> test_link;
  UnboundedSender<std::pair<Payload, LinkState>> tx;
  bool send(Payload value) const {
    LinkState state = (*test_link)();
    return tx.send({value, state});
  }
  bool is_closed() const { return tx.is_closed(); }
  auto closed() const { return tx.closed(); }
};
using PayloadReceiver = Stream<Payload>;

/// network.rs:51-65
struct Socket {
  virtual ~Socket() = default;
  virtual void deliver(SocketAddr src, SocketAddr dst, u64 tag, Payload msg) = 0;
  virtual void new_connection(SocketAddr, SocketAddr, PayloadSender, PayloadReceiver) {}
};

/// network.rs:20-313
class Network {
 public:
  struct Node {
    std::optional<u32> ip;
    std::map<std::pair<SocketAddr, IpProtocol>, std::shared_ptr<Socket>> sockets;
  };
  Network(GlobalRng& r, NetConfig c) : rand(r), config(c) {}
  GlobalRng& rand;
  NetConfig config;
  u64 msg_count = 0;
  std::map<u64, Node> nodes;
  std::map<u32, u64> addr_to_node;
  std::set<u64> clogged_node_in, clogged_node_out;
  std::set<std::pair<u64, u64>> clogged_link;

  void insert_node(u64 id) { nodes[id] = Node{}; }
  void reset_node(u64 id) { node(id).sockets.clear(); }
  void set_ip(u64 id, u32 ip) {
    Node& n = node(id);
    if (n.ip) addr_to_node.erase(*n.ip);
    n.ip = ip;
    if (addr_to_node.count(ip)) panic("IP conflict");
    addr_to_node[ip] = id;
  }
  void clog_node(u64 id, bool in, bool out) {
    node(id);
    if (in) clogged_node_in.insert(id);
    if (out) clogged_node_out.insert(id);
  }
  void unclog_node(u64 id, bool in, bool out) {
    node(id);
    if (in) clogged_node_in.erase(id);
    if (out) clogged_node_out.erase(id);
  }
  void clog_link(u64 src, u64 dst) {
    node(src), node(dst);
    clogged_link.insert({src, dst});
  }
  void unclog_link(u64 src, u64 dst) {
    node(src), node(dst);
    clogged_link.erase({src, dst});
  }
  bool link_clogged(u64 src, u64 dst) const {
    return clogged_node_out.count(src) || clogged_node_in.count(dst) || clogged_link.count({src, dst});
  }
  /// :206-251
  Result<SocketAddr> bind(u64 node_id, SocketAddr addr, IpProtocol proto, std::shared_ptr<Socket> socket) {
    Node& n = node(node_id);
    if (addr.ip != IP_UNSPECIFIED && !ip_is_loopback(addr.ip) && n.ip && addr.ip != *n.ip)
      return Result<SocketAddr>::Error(Err::AddrNotAvailable);
    if (addr.port == 0) {
      u32 port = 1;
      for (; port <= 0xffff; port++)
        if (!n.sockets.count({SocketAddr{addr.ip, u16(port)}, proto})) break;
      if (port > 0xffff) return Result<SocketAddr>::Error(Err::AddrInUse);
      addr.port = u16(port);
    }
    if (n.sockets.count({addr, proto})) return Result<SocketAddr>::Error(Err::AddrInUse);
    n.sockets[{addr, proto}] = std::move(socket);
    return Result<SocketAddr>::Ok(addr);
  }
  void close(u64 node_id, SocketAddr addr, IpProtocol proto) { node(node_id).sockets.erase({addr, proto}); }
  /// :261-270  — `clogged || gen_bool(loss)`: no draw when clogged
  std::optional<u64> test_link(u64 src, u64 dst) {
    if (link_clogged(src, dst) || rand.gen_bool(config.packet_loss_rate)) return std::nullopt;
    msg_count += 1;
    return rand.gen_range_duration(config.send_latency_lo, config.send_latency_hi);
  }
  /// :273-291
  std::optional<u64> resolve_dest_node(u64 node_id, SocketAddr dst, IpProtocol proto) {
    Node& n0 = node(node_id);
    if (ip_is_loopback(dst.ip) || n0.sockets.count({dst, proto})) return node_id;
    if (!n0.ip) return std::nullopt;  // "ip not set" — before any draw
    auto it = addr_to_node.find(dst.ip);
    if (it != addr_to_node.end()) return it->second;
    return std::nullopt;
  }
  struct SendOk {
    u32 src_ip;
    u64 dst_node;
    std::shared_ptr<Socket> socket;
    u64 latency;
  };
  /// :296-313 — draws are spent before the socket lookup
  std::optional<SendOk> try_send(u64 node_id, SocketAddr dst, IpProtocol proto) {
    auto dst_node = resolve_dest_node(node_id, dst, proto);
    if (!dst_node) return std::nullopt;
    auto latency = test_link(node_id, *dst_node);
    if (!latency) return std::nullopt;
    auto nit = nodes.find(*dst_node);
    if (nit == nodes.end()) return std::nullopt;
    auto& sockets = nit->second.sockets;
    auto it = sockets.find({dst, proto});
    if (it == sockets.end()) it = sockets.find({SocketAddr{IP_UNSPECIFIED, dst.port}, proto});
    if (it == sockets.end()) return std::nullopt;
    u32 src_ip = ip_is_loopback(dst.ip) ? IP_LOCALHOST : *node(node_id).ip;
    return SendOk{src_ip, *dst_node, it->second, *latency};
  }

 private:
  Node& node(u64 id) {
    auto it = nodes.find(id);
    if (it == nodes.end()) panic("node not found");
    return it->second;
  }
};

/// net/mod.rs:82-333
class NetSim : public SimulatorHooks {
 public:
  NetSim(GlobalRng& r, TimeHandle& t, NetConfig c) : network(r, c), rand(r), time(t) {}
  Network network;
  GlobalRng& rand;
  TimeHandle& time;

  void create_node(u64 id) override { network.insert_node(id); }
  void reset_node(u64 id) override { network.reset_node(id); }
  void set_ip(u64 id, u32 ip) { network.set_ip(id, ip); }
  void update_config(const std::function<void(NetConfig&)>& f) { f(network.config); }
  u64 msg_count() const { return network.msg_count; }
  void clog_node(u64 id) { network.clog_node(id, true, true); }
  void clog_node_in(u64 id) { network.clog_node(id, true, false); }
  void clog_node_out(u64 id) { network.clog_node(id, false, true); }
  void unclog_node(u64 id) { network.unclog_node(id, true, true); }
  void unclog_node_in(u64 id) { network.unclog_node(id, true, false); }
  void unclog_node_out(u64 id) { network.unclog_node(id, false, true); }
  void clog_link(u64 s, u64 d) { network.clog_link(s, d); }
  void unclog_link(u64 s, u64 d) { network.unclog_link(s, d); }

  /// :287-295
  Async<Unit> rand_delay() {
    u64 delay = rand.gen_range_u64(0, 5) * 1000ull;  // Duration::from_micros
    if (rand.buggify_with_prob(0.1)) delay = rand.gen_range_u64(1, 5) * 1000000000ull;
    co_await time.sleep(delay);
    co_return {};
  }
  /// :298-333 (no RPC hooks, no IPVS: SURVEY §2 rows 9/11 out of scope)
  Async<Result<Unit>> send(u64 node, u16 port, SocketAddr dst, IpProtocol proto, u64 tag, Payload msg) {
    co_await rand_delay();
    if (auto ok = network.try_send(node, dst, proto)) {
      SocketAddr src{ok->src_ip, port};
      auto socket = ok->socket;  // delivery target captured at send time
      time.add_timer(ok->latency, TimerMeta{MSIM_EV_DELIVER, u32(network.msg_count)},
                     [socket, src, dst, tag, msg]() { socket->deliver(src, dst, tag, msg); });
    }
    co_return Result<Unit>::Ok({});
  }
};

/// net/mod.rs:366-402: the receiving half is an async stream that delays every message until
/// its arrival time and backs off exponentially (1 ms .. 10 s) while the link is down
inline PayloadReceiver channel_stream(UnboundedReceiver<std::pair<Payload, LinkState>> rx,
                                      std::shared_ptr<std::function<LinkState()>> test_link, TimeHandle* time) {
  for (;;) {
    auto item = co_await rx.recv();
    if (!item) break;
    Payload value = item->first;
    LinkState state = item->second;
    // wait until the link is ready
    u64 backoff = 1000000ull;
    u64 arrive_time;
    for (;;) {
      if (state) {
        arrive_time = *state;
        break;
      }
      co_await time->sleep(backoff);
      backoff = std::min<u64>(backoff * 2, 10000000000ull);
      state = (*test_link)();  // retry
    }
    co_await time->sleep_until(arrive_time);
    co_yield value;
  }
}

struct Connect1Ok {
  PayloadSender tx;
  PayloadReceiver rx;
  SocketAddr src;
};
inline std::pair<PayloadSender, PayloadReceiver> net_channel(NetSim* net, u64 node, SocketAddr dst, IpProtocol proto) {
  auto [tx, rx] = unbounded_channel<std::pair<Payload, LinkState>>();
  auto test_link = std::make_shared<std::function<LinkState()>>([net, node, dst, proto]() -> LinkState {
    auto ok = net->network.try_send(node, dst, proto);
    if (!ok) return std::nullopt;
    return net->time.now_instant() + ok->latency;
  });
  PayloadSender sender{test_link, std::move(tx)};
  return {std::move(sender), channel_stream(std::move(rx), test_link, &net->time)};
}
/// net/mod.rs:337-364: the connection request is delivered IMMEDIATELY (the reference's FIXME)
inline Async<Result<Connect1Ok>> net_connect1(NetSim* net, u64 node, u16 port, SocketAddr dst, IpProtocol proto) {
  co_await net->rand_delay();
  auto ok = net->network.try_send(node, dst, proto);
  if (!ok) co_return Result<Connect1Ok>::Error(Err::ConnectionRefused);
  SocketAddr src{ok->src_ip, port};
  auto [tx1, rx1] = net_channel(net, node, dst, proto);
  auto [tx2, rx2] = net_channel(net, ok->dst_node, src, proto);
  ok->socket->new_connection(src, dst, std::move(tx2), std::move(rx1));
  co_return Result<Connect1Ok>::Ok(Connect1Ok{std::move(tx1), std::move(rx2), src});
}

// ---- tokio::sync::oneshot (only what endpoint.rs uses) ----
template <class T>
struct OneshotState {
  std::optional<T> value;
  std::optional<Waker> rx_waker;
  bool rx_closed = false, tx_dropped = false;
};
template <class T>
struct OneshotSender {
  std::shared_ptr<OneshotState<T>> st;
  OneshotSender() = default;
  explicit OneshotSender(std::shared_ptr<OneshotState<T>> s) : st(std::move(s)) {}
  OneshotSender(OneshotSender&&) = default;
  OneshotSender& operator=(OneshotSender&& o) {
    drop();
    st = std::move(o.st);
    return *this;
  }
  ~OneshotSender() { drop(); }
  /// Err(value) if the receiver is gone
  std::optional<T> send(T v) {
    auto s = std::move(st);
    if (s->rx_closed) return v;
    s->value = std::move(v);
    if (s->rx_waker) s->rx_waker->wake();
    return std::nullopt;
  }

 private:
  void drop() {
    if (st) {
      st->tx_dropped = true;
      if (st->rx_waker) st->rx_waker->wake();
      st.reset();
    }
  }
};
template <class T>
struct OneshotReceiver {
  std::shared_ptr<OneshotState<T>> st;
  OneshotReceiver() = default;
  explicit OneshotReceiver(std::shared_ptr<OneshotState<T>> s) : st(std::move(s)) {}
  OneshotReceiver(OneshotReceiver&&) = default;
  OneshotReceiver& operator=(OneshotReceiver&&) = default;
  ~OneshotReceiver() {
    if (st) st->rx_closed = true;
  }
  Poll<Result<T>> poll(Context& cx) {
    if (st->value) {
      T v = std::move(*st->value);
      st->value.reset();
      return Poll<Result<T>>::done(Result<T>::Ok(std::move(v)));
    }
    if (st->tx_dropped) return Poll<Result<T>>::done(Result<T>::Error(Err::BrokenPipe));
    st->rx_waker = cx.waker;
    return Poll<Result<T>>::pending();
  }
};

struct Message {  // endpoint.rs:290-294
  u64 tag;
  Payload data;
  SocketAddr from;
};

/// endpoint.rs:298-363
inline thread_local size_t* tls_peak_registered = nullptr;  // capacity statistics only
inline thread_local size_t* tls_peak_msgs = nullptr;
struct Mailbox {
  std::vector<std::pair<u64, OneshotSender<Message>>> registered;
  std::vector<Message> msgs;
  size_t peak_registered = 0, peak_msgs = 0;
  ~Mailbox() {
    if (tls_peak_registered) *tls_peak_registered = std::max(*tls_peak_registered, peak_registered);
    if (tls_peak_msgs) *tls_peak_msgs = std::max(*tls_peak_msgs, peak_msgs);
  }
  void deliver(Message m) {
    size_t i = 0;
    std::optional<Message> msg = std::move(m);
    while (i < registered.size()) {
      if (msg->tag == registered[i].first) {
        OneshotSender<Message> sender = std::move(registered[i].second);  // swap_remove(i)
        if (i + 1 != registered.size()) registered[i] = std::move(registered.back());
        registered.pop_back();
        auto back = sender.send(std::move(*msg));
        if (!back) return;
        msg = std::move(back);
      } else {
        i += 1;
      }
    }
    msgs.push_back(std::move(*msg));
    peak_msgs = std::max(peak_msgs, msgs.size());
  }
  OneshotReceiver<Message> recv(u64 tag) {
    auto st = std::make_shared<OneshotState<Message>>();
    OneshotSender<Message> tx(st);
    OneshotReceiver<Message> rx(st);
    size_t idx = 0;
    for (; idx < msgs.size(); idx++)
      if (msgs[idx].tag == tag) break;
    if (idx < msgs.size()) {
      Message m = std::move(msgs[idx]);  // swap_remove(idx)
      if (idx + 1 != msgs.size()) msgs[idx] = std::move(msgs.back());
      msgs.pop_back();
      tx.send(std::move(m));
    } else {
      registered.emplace_back(tag, std::move(tx));
      peak_registered = std::max(peak_registered, registered.size());
    }
    return rx;
  }
};

struct BindGuard;
struct IncomingConn {
  PayloadSender tx;
  PayloadReceiver rx;
  SocketAddr peer;
};
struct EndpointSocket : Socket {
  Mailbox mailbox;
  /// async_channel::unbounded of incoming connections (endpoint.rs:18,306)
  std::pair<UnboundedSender<std::shared_ptr<IncomingConn>>, UnboundedReceiver<std::shared_ptr<IncomingConn>>> conns =
      unbounded_channel<std::shared_ptr<IncomingConn>>();
  void deliver(SocketAddr src, SocketAddr, u64 tag, Payload msg) override {
    mailbox.deliver(Message{tag, msg, src});
  }
  void new_connection(SocketAddr src, SocketAddr, PayloadSender tx, PayloadReceiver rx) override {
    conns.first.send(std::make_shared<IncomingConn>(IncomingConn{std::move(tx), std::move(rx), src}));  // try_send
  }
};

/// endpoint.rs:228-274 Sender / Receiver of a reliable connection
struct ConnSender {
  std::shared_ptr<BindGuard> guard;
  PayloadSender tx;
  Result<Unit> send(Payload v) const {  // `async fn send` has no await point
    if (!tx.send(v)) return Result<Unit>::Error(Err::ConnectionReset);
    return Result<Unit>::Ok({});
  }
  bool is_closed() const { return tx.is_closed(); }
  auto closed() const { return tx.closed(); }  // endpoint.rs:251-253
};
struct ConnReceiver {
  std::shared_ptr<BindGuard> guard;
  PayloadReceiver rx;
  Async<Result<Payload>> recv() {
    auto v = co_await StreamNext<Payload>{&rx};
    if (!v) co_return Result<Payload>::Error(Err::ConnectionReset);
    co_return Result<Payload>::Ok(*v);
  }
};
struct Connection {
  ConnSender tx;
  ConnReceiver rx;
  SocketAddr peer;
};

/// net/mod.rs:436-494
struct BindGuard {
  NetSim* net;
  std::shared_ptr<NodeInfo> node;
  SocketAddr addr;
  IpProtocol protocol;
  BindGuard(NetSim* n, std::shared_ptr<NodeInfo> nd, SocketAddr a, IpProtocol p)
      : net(n), node(std::move(nd)), addr(a), protocol(p) {}
  BindGuard(const BindGuard&) = delete;
  BindGuard& operator=(const BindGuard&) = delete;
  ~BindGuard() {
    if (node->is_killed()) return;  // avoid interfering with restarted node
    net->network.close(node->id, addr, protocol);
  }
};

NetSim& current_net();

/// net/endpoint.rs:11-175
class Endpoint {
 public:
  std::shared_ptr<BindGuard> guard;
  std::shared_ptr<EndpointSocket> socket;
  std::shared_ptr<std::optional<SocketAddr>> peer;

  /// endpoint.rs:23-36 + BindGuard::bind (net/mod.rs:446-480)
  static Async<Result<Endpoint>> bind(SocketAddr addr) {
    auto socket = std::make_shared<EndpointSocket>();
    NetSim& net = current_net();
    auto node = current_task()->node;
    // lookup_host yields exactly one address here
    co_await net.rand_delay();
    auto r = net.network.bind(node->id, addr, IpProtocol::Udp, socket);
    if (r.is_err()) co_return Result<Endpoint>::Error(r.err);
    Endpoint ep;
    ep.guard = std::make_shared<BindGuard>(&net, node, *r.ok, IpProtocol::Udp);
    ep.socket = socket;
    ep.peer = std::make_shared<std::optional<SocketAddr>>();
    co_return Result<Endpoint>::Ok(std::move(ep));
  }
  /// endpoint.rs:39-44
  static Async<Result<Endpoint>> connect(SocketAddr addr) {
    auto r = co_await bind(SocketAddr{IP_UNSPECIFIED, 0});
    if (r.is_err()) co_return r;
    *r.ok->peer = addr;
    co_return r;
  }
  SocketAddr local_addr() const { return guard->addr; }
  Result<SocketAddr> peer_addr() const {
    if (!*peer) return Result<SocketAddr>::Error(Err::NotConnected);
    return Result<SocketAddr>::Ok(**peer);
  }
  /// endpoint.rs:120-133
  Async<Result<Unit>> send_to_raw(SocketAddr dst, u64 tag, Payload data) const {
    auto g = guard;  // keep the port bound while the send is in flight
    co_return co_await g->net->send(g->node->id, g->addr.port, dst, IpProtocol::Udp, tag, data);
  }
  /// endpoint.rs:140-149
  Async<Result<std::pair<Payload, SocketAddr>>> recv_from_raw(u64 tag) const {
    using R = Result<std::pair<Payload, SocketAddr>>;
    auto g = guard;
    auto s = socket;
    auto recver = s->mailbox.recv(tag);
    auto msg = co_await recver;
    if (msg.is_err()) co_return R::Error(Err::BrokenPipe);
    co_await g->net->rand_delay();
    co_return R::Ok({msg.ok->data, msg.ok->from});
  }
  // send_to/recv_from (endpoint.rs:69-94) only add a byte copy around the raw calls
  Async<Result<Unit>> send_to(SocketAddr dst, u64 tag, Payload data) const { return send_to_raw(dst, tag, data); }
  Async<Result<std::pair<Payload, SocketAddr>>> recv_from(u64 tag) const { return recv_from_raw(tag); }
  Async<Result<Unit>> send(u64 tag, Payload data) const {
    auto p = peer_addr();
    if (p.is_err()) co_return Result<Unit>::Error(p.err);
    co_return co_await send_to(*p.ok, tag, data);
  }
  Async<Result<Payload>> recv(u64 tag) const {
    auto p = peer_addr();
    if (p.is_err()) co_return Result<Payload>::Error(p.err);
    auto r = co_await recv_from(tag);
    if (r.is_err()) co_return Result<Payload>::Error(r.err);
    if (r.ok->second != *p.ok) panic("receive a message but not from the connected address");
    co_return Result<Payload>::Ok(r.ok->first);
  }

  // ---- reliable connections (endpoint.rs:178-211) ----
  /// connect1: NetSim::connect1 with this endpoint's node and port
  Async<Result<Connection>> connect1(SocketAddr addr) const {
    auto g = guard;
    auto r = co_await net_connect1(g->net, g->node->id, g->addr.port, addr, IpProtocol::Udp);
    if (r.is_err()) co_return Result<Connection>::Error(r.err);
    co_return Result<Connection>::Ok(
        Connection{ConnSender{g, std::move(r.ok->tx)}, ConnReceiver{g, std::move(r.ok->rx)}, addr});
  }
  /// accept1: rand_delay FIRST, then the next incoming connection
  Async<Result<Connection>> accept1() const {
    auto g = guard;
    auto s = socket;
    co_await g->net->rand_delay();
    auto c = co_await s->conns.second.recv();
    if (!c) co_return Result<Connection>::Error(Err::ConnectionReset);
    co_return Result<Connection>::Ok(
        Connection{ConnSender{g, std::move((*c)->tx)}, ConnReceiver{g, std::move((*c)->rx)}, (*c)->peer});
  }

  // ---- net/rpc.rs:82-166 ----
  /// call_with_data (:107-131): rsp_tag = random::<u64>(); send; recv(rsp_tag)
  Async<Result<Payload>> call(SocketAddr dst, u64 req_tag, Payload request) const {
    u64 rsp_tag = g_rand().gen_u64();
    request.w = rsp_tag;
    auto s = co_await send_to_raw(dst, req_tag, request);
    if (s.is_err()) co_return Result<Payload>::Error(s.err);
    auto r = co_await recv_from_raw(rsp_tag);
    if (r.is_err()) co_return Result<Payload>::Error(r.err);
    if (r.ok->second != dst) panic("assertion failed: from == dst");
    co_return Result<Payload>::Ok(r.ok->first);
  }
  /// call_timeout (:96-105)
  Async<Result<Payload>> call_timeout(SocketAddr dst, u64 req_tag, Payload request, u64 timeout_ns) const;
  /// add_rpc_handler (:134-166): a loop task that spawns one task per request
  void add_rpc_handler(u64 req_tag, std::function<Async<Payload>(Payload)> f) const;

 private:
  static GlobalRng& g_rand();
};

/// rpc.rs:71-81 hash33
constexpr u64 hash_str(const char* s) {
  u64 h = 0;
  for (; *s; ++s) h = h * 33 + u64(u8(*s));
  return h;
}

}  // namespace msim_oracle
