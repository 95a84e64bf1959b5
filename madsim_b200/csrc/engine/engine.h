// msim_b200 engine — the per-seed executor step, one lane per seed.
//
// Replaces, for registered workloads, the reference's hot path:
//   Executor::block_on / run_all_ready        madsim/src/sim/task/mod.rs:220-323
//   mpsc::Receiver::try_recv_random            madsim/src/sim/utils/mpsc.rs:73-83
//   TimeRuntime::advance_to_next_event         madsim/src/sim/time/mod.rs:45-60
//   TimeHandle::{advance,sleep,add_timer}      madsim/src/sim/time/mod.rs:103-124,158-169
//   naive_timer::Timer (BinaryHeap)            crate naive-timer 0.2 + std BinaryHeap sift rules
//   GlobalRng / xoshiro256++ / gen_range       madsim/src/sim/rand.rs:42-89 + rand 0.8 / rand_xoshiro 0.6
//   NetSim::{rand_delay,send}, Network::*      madsim/src/sim/net/mod.rs:287-333, network.rs:142-313
//   Endpoint mailbox, rpc call/call_timeout    madsim/src/sim/net/endpoint.rs:298-363, net/rpc.rs:96-131
//   TaskHandle::{kill,restart,pause,resume}    madsim/src/sim/task/mod.rs:356-424
//
// Design (B200-first, not a translation).
//  * All per-seed state is ONE bounded block (Layout::BYTES) that lives in shared memory while a
//    CTA advances its seeds; the hot scalars (rng, clock, trace hash, queue lengths) are cached
//    in registers for the whole burst.  The block is seed-major (one base register + immediate
//    offsets per access): HBM<->smem staging is a bulk TMA copy of whole blocks, so the HBM image
//    needs no columnar interleave for coalescing, and address arithmetic is what the instruction
//    -issue-bound kernel cannot afford (profiles/: 0.42 IPC, 57 % I-cache hit rate when every
//    primitive was inlined at every await site).
//  * REQUEST / DISPATCH: a task's future is an explicit state machine (`WL::fsm`) that contains
//    only workload logic.  Whenever it needs a runtime primitive — sleep, rand_delay, bind, send,
//    recv, call_timeout, spawn, gen_range, join, time::advance — it RETURNS a request; the engine
//    executes it in one shared place (`poll_task`) and re-enters the state machine.  Each
//    primitive's poll-level behaviour therefore exists exactly once in the
//    instruction stream, and lanes of a warp that wait in the same primitive execute it together.
//  * Every bounded table reports overflow as MSIM_ST_FAIL_CAPACITY — never silently.
#pragma once
#include <type_traits>
#include "defs.h"

namespace msim {

constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- what a suspended future is waiting for (t_wait): state[3:0] | WF_* ----
enum : u8 {
  W_NONE = 0,
  W_SLEEP = 1,       // Sleep on dl0                                   -> resume the state machine
  W_BIND = 2,        // Endpoint::bind: rand_delay, then Network::bind -> ret = socket
  W_SEND = 3,        // send_to_raw: rand_delay, then NetSim::send of the prepared message
  W_RECV_RX = 4,     // recv_from_raw: oneshot not filled yet
  W_RECV_DELAY = 5,  // recv_from_raw: rand_delay AFTER the receipt   -> rv0/rv1/rfrom/rw
  W_CALL_SEND = 6,   // call_timeout: inside call's send_to_raw
  W_CALL_RX = 7,     // call_timeout: inside call's recv_from_raw (oneshot)
  W_CALL_DELAY = 8,  // call_timeout: inside call's post-receipt rand_delay
  W_JOIN = 9,        // JoinHandle.await
  W_CONNECT = 10,    // connect1: rand_delay, then the connection request (net/mod.rs:337-364)
  W_ACCEPT = 11,     // accept1: rand_delay FIRST, then the next incoming connection (endpoint.rs:197-211)
  W_CH_RECV = 12,    // Receiver::recv: the channel's delay/back-off stream (net/mod.rs:385-402)
  W_CTRLC = 13,      // signal::ctrl_c: watch::Receiver::changed() on the node's channel (signal.rs:4-8)
  WF_RD = 0x10,      // rand_delay not drawn yet: draw it now, then sleep on dl0
  WF_SLEEP = 0x20,   // sleeping on dl0
  WF_CLOSED = 0x40,  // select_biased!{ conn.tx.closed() => .., <the wait> }: closed() is polled first (W_SLEEP, W_CH_RECV)
};

// ---- requests a state machine returns to the engine ----
enum : int {
  RQ_PENDING = 0,   // Poll::Pending with nothing registered (woken by someone else)
  RQ_READY,         // Poll::Ready
  RQ_PANIC,         // ya = code
  RQ_SLEEP_FOR,     // time::sleep(xa ns)
  RQ_SLEEP_UNTIL,   // time::sleep_until(xa)
  RQ_RAND_DELAY,    // NetSim::rand_delay
  RQ_BIND,          // Endpoint::bind(ya = addr)                          -> ret = socket or <0
  RQ_SEND,          // ep.send_to_raw(ya = dst, tag xa, payload {w: xb, v: [yb, yc]})
  RQ_RECV,          // ep.recv_from_raw(tag xa)                           -> rv0, rv1, rfrom, rw
  RQ_CALL,          // ep.call_timeout(ya = dst, req tag xa, {yb, yc}, timeout xb ns) -> ret, rv0, rv1
  RQ_SPAWN,         // task::spawn: ya = kind | SPAWN_*; yb = node; locals a = yc, b = yd, c = lo32(xa), dl1 = xb -> ret
  RQ_CREATE_NODE,   // Handle::create_node().ip().init().build(): ya = NODE_* flags, yb = restart_on_panic_matching mask (NODE_PANIC_MATCHING) -> ret = node id
  RQ_GEN_RANGE,     // thread_rng().gen_range(xa..xb) (u64)                -> ret64
  RQ_JOIN,          // JoinHandle(ya = task slot).await                   -> ret (1 ok / 2 cancelled), rv0 = output
  RQ_ADVANCE,       // time::advance(xa ns)
  RQ_CONNECT1,      // ep.connect1(ya = addr)          -> ret 0 (the task now holds the connection end) / -1 refused
  RQ_ACCEPT1,       // ep.accept1()                    -> ret 0, rfrom = peer address
  RQ_CH_SEND,       // conn.tx.send({yb, yc})          -> ret 0 / -1 ConnectionReset   (no await point)
  RQ_CH_RECV,       // conn.rx.recv()                  -> ret 0, rv0, rv1 / -1 ConnectionReset
  RQ_SLEEP_RAW,     // Sleep::reset(xa) + await: the deadline is NOT clamped to now + 1 ms (time/sleep.rs:40-42, interval.rs:156)
  RQ_CTRL_C,        // signal::ctrl_c().await (installs the node's handler on first use)
  RQ_SLEEP_OR_CLOSED,  // select_biased!{ conn.tx.closed() => -2, sleep(xa ns) => 0 }  (madsim-tonic server.rs:289-300)
  RQ_CH_RECV_OR_CLOSED,  // select_biased!{ conn.tx.closed() => -2, conn.rx.recv() => 0 / -1 }
  RQ_TIMEOUT_BEGIN,  // time::timeout(xa ns, async { ..the awaits up to RQ_TIMEOUT_END.. }) (time/mod.rs:128-140): the Sleep is
                     // created now; at every pending poll the inner future goes first, then the Sleep; when it fires the
                     // pending await is dropped where it stands and the state machine continues with ret = -3 (Elapsed)
  RQ_TIMEOUT_END,
};
enum : u32 { SPAWN_EP = 0x100, SPAWN_JOINABLE = 0x200, SPAWN_CONN = 0x400, SPAWN_CONN_TX = 0x800 };  // clone the Endpoint / keep a JoinHandle / move the connection end / move only its Sender
enum : u32 { NODE_IP = 1, NODE_INIT = 2, NODE_RESTART_ON_PANIC = 4, NODE_PANIC_MATCHING = 8 };
enum : int { CALL_OK = 1, CALL_TIMEOUT = 2 };

/// Optional Caps members (absent = off).  HEAP16: timer heap entries are 16-byte {deadline, payload} records moved with
/// one 128-bit shared-memory access instead of a u64 array + a u32 array (12 bytes per entry, two accesses).
template <class C, class = void>
struct caps_heap16 { static constexpr bool value = MSIM_HEAP16_DEFAULT; };
template <class C>
struct caps_heap16<C, std::void_t<decltype(C::HEAP16)>> { static constexpr bool value = C::HEAP16; };

/// Compile-time layout of one seed's state block for capacities `C` (element offsets per section).
template <class C>
struct Layout {
  static constexpr bool H16 = MSIM_X_H16 < 0 ? caps_heap16<C>::value : (MSIM_X_H16 != 0);
  static_assert(C::NT <= 128, "task bitmap is 2 x u64");
  static_assert(C::NT < 255 && C::NMSG < 255 && C::NSOCK < 255, "u8 indices");
  // ---- u64 section ----
  static constexpr int Q_RNG = 0;      // 4
  static constexpr int Q_NOW = 4;
  static constexpr int Q_HASH = 5;
  static constexpr int Q_DRAWS = 6;
  static constexpr int Q_TFREE = 7;    // 2: free-task bitmap
  static constexpr int MW = (C::NMSG + 63) / 64;
  static constexpr int Q_MFREE = 9;    // MW: free-message bitmap (one word for every program's default tables)
  static constexpr int Q_HEAP_DL = (Q_MFREE + MW + 1) & ~1;  // (even: the 16-byte heap records stay aligned)                    // [TQ] timer deadlines (implicit binary heap); H16: [2*TQ] {deadline, payload} records
  static constexpr int Q_T_DL0 = Q_HEAP_DL + (H16 ? 2 : 1) * C::TQ;  // [NT] Sleep deadline of the task's current await
  static constexpr int Q_T_DL1 = Q_T_DL0 + C::NT;           // [NT] timeout deadline / generic u64 local
  static constexpr int NTT = C::USES_TIMEOUT ? C::NT : 0;
  static constexpr int Q_T_DL2 = Q_T_DL1 + C::NT;           // [NTT] deadline of the time::timeout scope the future is inside, or NONE64
  static constexpr int Q_M_TAG = Q_T_DL2 + NTT;             // [NMSG]
  static constexpr int Q_M_W = Q_M_TAG + C::NMSG;           // [NMSG] payload.w (rpc response tag)
  static constexpr int Q_RG_TAG = Q_M_W + C::NMSG;          // [NSOCK*NREG] registered receive tags
  static constexpr int Q_WL = Q_RG_TAG + C::NSOCK * C::NREG;  // [WQ*NNODE] workload per-node u64
  // reliable channels: direction d = 2*conn + side (0: connector->acceptor, 1: the reverse)
  static constexpr int ND = 2 * C::NCONN;
  static constexpr int Q_CH_ARRIVE = Q_WL + C::WQ * C::NNODE;  // [ND*NCQ] queued item's link state: arrive time or NONE64
  static constexpr int Q_CH_CUR = Q_CH_ARRIVE + ND * C::NCQ;   // [ND] link state of the item the stream is delaying
  static constexpr int Q_CH_DL = Q_CH_CUR + ND;                // [ND] the stream's Sleep deadline
  static constexpr int Q_CH_BACKOFF = Q_CH_DL + ND;            // [ND]
  static constexpr int NQ = Q_CH_BACKOFF + ND;
  // ---- u32 section ----
  enum : int {
    W_EV_POP = 0, W_EV_FIRE, W_HEAP_LEN, W_RQ_LEN, W_NEXT_SPAWN, W_MSG_COUNT, W_STATUS, W_FAIL_CODE,
    W_PEAK_TIMERS, W_LIVE_TASKS, W_PEAK_TASKS, W_USER0, W_USER1, W_USER2, W_USER3, W_CLOG_IN, W_CLOG_OUT,
    W_PK_LEN, W_BASE_SECS, W_NEXT_NODE, W_MAIN_SLOT, W_PANIC_CODE, W_RESUME, W_NSCALAR
  };
  static constexpr int W_CLOG_LINK = W_NSCALAR;                 // [NNODE] bit d set: link n->d clogged
  static constexpr int W_HEAP_PL = W_CLOG_LINK + C::NNODE;      // [TQ] timer payloads (H16: inside the records, nothing here)
  static constexpr int W_T_SPAWN = W_HEAP_PL + (H16 ? 0 : C::TQ);  // [NT] spawn index (NONE32 = free slot)
  static constexpr int W_T_AW = W_T_SPAWN + C::NT;              // [NT] awaiter (slot<<20|gen) or NONE32
  static constexpr int W_T_A = W_T_AW + C::NT;                  // [NT] locals
  static constexpr int W_T_B = W_T_A + C::NT;
  static constexpr int W_T_C = W_T_B + C::NT;
  static constexpr int W_M_V0 = W_T_C + C::NT;                  // [NMSG] payload v[0]
  static constexpr int W_M_V1 = W_M_V0 + C::NMSG;
  static constexpr int W_M_FROM = W_M_V1 + C::NMSG;             // source address (before the send: the destination)
  static constexpr int W_M_DST = W_M_FROM + C::NMSG;            // sock<<16 | gen16, captured at send
  static constexpr int W_M_SERIAL = W_M_DST + C::NMSG;
  static constexpr int W_RG_WAITER = W_M_SERIAL + C::NMSG;      // [NSOCK*NREG] slot<<20|gen
  static constexpr int W_SK_PORTGEN = W_RG_WAITER + C::NSOCK * C::NREG;  // [NSOCK] port | gen<<16
  static constexpr int W_WL = W_SK_PORTGEN + C::NSOCK;          // [WW*NNODE] workload per-node u32
  static constexpr int W_WLS = W_WL + C::WW * C::NNODE;         // [WS] workload scalars
  static constexpr int W_CH_V0 = W_WLS + C::WS;                 // [ND*NCQ] queued payloads
  static constexpr int W_CH_V1 = W_CH_V0 + ND * C::NCQ;
  static constexpr int W_CH_CUR_V0 = W_CH_V1 + ND * C::NCQ;     // [ND]
  static constexpr int W_CH_CUR_V1 = W_CH_CUR_V0 + ND;
  static constexpr int W_CH_DST = W_CH_CUR_V1 + ND;             // [ND] test_link destination address
  static constexpr int W_CH_WAITER = W_CH_DST + ND;             // [ND] mpsc rx_waker (slot<<20|gen) or NONE32
  static constexpr int NDC = C::USES_CLOSED ? ND : 0;
  static constexpr int W_CH_CLOSED_WAITER = W_CH_WAITER + ND;   // [NDC] the task parked in Sender::closed() of this direction
  static constexpr int W_CONN_SRC = W_CH_CLOSED_WAITER + NDC;   // [NCONN] connector's source address
  static constexpr int NSA = C::NCONN ? C::NSOCK : 0;           // accept state exists only with connections
  static constexpr int W_ACC_WAITER = W_CONN_SRC + C::NCONN;    // [NSA] task waiting in accept1
  static constexpr int NCC = C::USES_CTRLC ? C::NNODE : 0;     // ctrl-c state exists only where it is awaited
  static constexpr int W_N_CTRLC = W_ACC_WAITER + NSA;          // [NCC] NodeInfo::ctrl_c: installed<<31 | watch version
  static constexpr int W_N_CTRLC_WAITER = W_N_CTRLC + NCC;      // [NCC] the receiver parked in changed() (slot<<20|gen) or NONE32
  static constexpr int NPM = C::USES_PMATCH ? C::NNODE : 0;      // restart_on_panic_matching exists only where a program sets it
  static constexpr int W_N_PMATCH = W_N_CTRLC_WAITER + NCC;     // [NPM] bit c set: a panic with code Caps::PMATCH_BASE + c restarts the node
  static constexpr int NW = W_N_PMATCH + NPM;
  // ---- u8 section ----
  static constexpr int B_RQ = 0;                        // [NT] ready "queue": a Vec popped at a random index
  static constexpr int B_T_KIND = B_RQ + C::NT;
  static constexpr int B_T_PC = B_T_KIND + C::NT;
  static constexpr int B_T_FLAGS = B_T_PC + C::NT;
  static constexpr int B_T_NODE = B_T_FLAGS + C::NT;
  static constexpr int B_T_EPOCH = B_T_NODE + C::NT;
  static constexpr int B_T_WAIT = B_T_EPOCH + C::NT;    // W_* | WF_*: what the suspended future awaits
  static constexpr int B_T_RX = B_T_WAIT + C::NT;       // oneshot slot: delivered message or NONE8
  static constexpr int B_T_SOCK = B_T_RX + C::NT;       // endpoint this future holds (BindGuard ref) or NONE8
  static constexpr int B_T_SOCKGEN = B_T_SOCK + C::NT;
  static constexpr int B_T_MSG = B_T_SOCKGEN + C::NT;   // message slot owned by the future or NONE8
  static constexpr int B_MB = B_T_MSG + C::NT;          // [NSOCK*NMBOX] mailbox.msgs (message slots)
  static constexpr int B_SK_NODE = B_MB + C::NSOCK * C::NMBOX;
  static constexpr int B_SK_IPK = B_SK_NODE + C::NSOCK;
  static constexpr int B_SK_STATE = B_SK_IPK + C::NSOCK;
  static constexpr int B_SK_REGLEN = B_SK_STATE + C::NSOCK;
  static constexpr int B_SK_MBLEN = B_SK_REGLEN + C::NSOCK;
  static constexpr int B_SK_REFS = B_SK_MBLEN + C::NSOCK;
  static constexpr int B_SK_EPOCH = B_SK_REFS + C::NSOCK;
  static constexpr int B_N_FLAGS = B_SK_EPOCH + C::NSOCK;  // [NNODE]
  static constexpr int B_N_EPOCH = B_N_FLAGS + C::NNODE;
  static constexpr int B_PK = B_N_EPOCH + C::NNODE;        // [NPARK] parked runnables, park order
  static constexpr int B_T_CONN = B_PK + C::NPARK;         // [NT] connection end the future holds: conn<<1|side, or NONE8
  static constexpr int NTS = C::USES_SPLIT ? C::NT : 0;
  static constexpr int B_T_CONNHALF = B_T_CONN + (C::NCONN ? C::NT : 0);  // [NTS] halves of the end this future owns: 1 Sender, 2 Receiver
  static constexpr int B_T_HANDLE = B_T_CONNHALF + NTS;    // [NTS] task whose FallibleTask (cancel_on_drop) this future holds, or NONE8
  static constexpr int NT2 = C::USES_CONN2 ? C::NT : 0;
  static constexpr int B_T_CONN2 = B_T_HANDLE + NTS;       // [NT2] a second connection end the future holds (parked: conn_swap)
  static constexpr int B_CH_HEAD = B_T_CONN2 + NT2;        // [ND] FIFO head / length
  static constexpr int B_CH_LEN = B_CH_HEAD + ND;
  static constexpr int B_CH_PHASE = B_CH_LEN + ND;         // [ND] 0 idle, 1 back-off sleep, 2 arrival sleep
  static constexpr int B_CH_SRC = B_CH_PHASE + ND;         // [ND] test_link source node
  static constexpr int B_CH_FLAGS = B_CH_SRC + ND;         // [ND] CH_TX_ALIVE | CH_RX_ALIVE
  static constexpr int B_CONN_USED = B_CH_FLAGS + ND;      // [NCONN]
  static constexpr int B_ACC = B_CONN_USED + C::NCONN;     // [NSOCK*NACC] incoming connections (async_channel)
  static constexpr int B_ACC_LEN = B_ACC + NSA * C::NACC;
  static constexpr int NB = B_ACC_LEN + NSA;

  static constexpr int W_BASE = 8 * NQ;                 // byte offset of the u32 section
  static constexpr int B_BASE = W_BASE + 4 * NW;        // ... of the u8 section
  static constexpr int BYTES = round_up(B_BASE + NB, 16);  // one seed; a CTA's block is S * BYTES
};

/// COLD_NOW_SMEM: around an out-of-line call the clock is written to / re-read from the state block instead of staying live in
/// registers.  Pays where such a call sits on the hot path (ETCD's once-per-request spawn: +7 %), costs elsewhere (TONIC -6 %).
template <class C, class = void>
struct caps_cold_now_smem { static constexpr bool value = false; };
template <class C>
struct caps_cold_now_smem<C, std::void_t<decltype(C::COLD_NOW_SMEM)>> { static constexpr bool value = C::COLD_NOW_SMEM; };

/// PORT_FIRST: the socket lookup tests the port before state / node / protocol.  The same predicate; which order is faster is
/// a code-placement matter per program (profiles/r2ad_port_first.txt: Raft +2.7 %, ETCD -3.3 %, TONIC / ELECTION +-0.5 %).
template <class C, class = void>
struct caps_port_first { static constexpr bool value = false; };
template <class C>
struct caps_port_first<C, std::void_t<decltype(C::PORT_FIRST)>> { static constexpr bool value = C::PORT_FIRST; };
template <class C, class = void>
struct caps_bind_inline { static constexpr bool value = false; };
template <class C>
struct caps_bind_inline<C, std::void_t<decltype(C::BIND_INLINE)>> { static constexpr bool value = C::BIND_INLINE; };

/// CAPACITY ESCALATION.  Every table of a seed's state block is bounded (Caps); a seed that overflows one ends with
/// MSIM_ST_FAIL_CAPACITY instead of a wrong result.  BigWL<WL> is the same program over tables 2-4x the size (inside
/// the index widths the engine packs: task / message slots in 8 bits, the task bitmap, 64 connections); the
/// host re-runs the overflowed seeds of a batch on it (msim_b200.cu: escalate) — one seed per warp, the general
/// kernel only, so the batch kernels and their occupancy are untouched.
template <class C, class = void>
struct caps_big_scale { static constexpr int value = 1; };
template <class C>
struct caps_big_scale<C, std::void_t<decltype(C::BIG_SCALE)>> { static constexpr int value = C::BIG_SCALE; };
template <class C>
struct BigOf : C {
  static constexpr int cap_(int v, int m) { return v < m ? v : m; }
  static constexpr int K = caps_big_scale<C>::value;  // programs whose default tables are cut to the BASELINE size say 2
  static constexpr int NT = cap_(4 * K * C::NT, 128), TQ = 4 * K * C::TQ, NMSG = cap_(4 * K * C::NMSG, 192);
  static constexpr int NSOCK = cap_(2 * K * C::NSOCK, 64), NREG = cap_(4 * K * C::NREG, 64), NMBOX = cap_(4 * K * C::NMBOX, 64);
  static constexpr int NPARK = cap_(4 * K * C::NPARK, NT);
  static constexpr int NCONN = cap_(2 * K * C::NCONN, 64), NCQ = 2 * C::NCQ, NACC = cap_(2 * K * C::NACC, 64);
};
template <class WL>
struct BigWL : WL {
  using Caps = BigOf<typename WL::Caps>;
};

/// One lane's view of its seed.  `WL` supplies Caps + the state machines; `TRACE` = the variant that
/// also appends full trace records (single-seed replay, msim_batch_trace) — the batch kernels carry
/// neither its registers nor its code.
template <class WL, bool TRACE = true>
struct Engine {
  using C = typename WL::Caps;
  using L = Layout<C>;

  unsigned char* sb;  // this seed's state block (shared memory on the device)
  const Params* P;
  // register-cached hot state
  u64 r0, r1, r2, r3, now, hash, draws;
  u32 ev_pop, ev_fire, heap_len, rq_len, status, peak_timers;
  // request (state machine -> engine) and result (engine -> state machine) registers
  u64 xa, xb, ret64, rw;
  u32 ya, yb, yc, yd, rv0, rv1, rfrom;
  int ret;
  // optional full trace (single-seed replay)
  msim_trace_rec* tr;
  u32 tr_cap, tr_n;

  // ---------------- accessors ----------------
  MSIM_HD u64& q(int off, int i = 0) const { return reinterpret_cast<u64*>(sb)[off + i]; }
  MSIM_HD u32& w(int off, int i = 0) const { return reinterpret_cast<u32*>(sb + L::W_BASE)[off + i]; }
  MSIM_HD u8& b(int off, int i = 0) const { return (sb + L::B_BASE)[off + i]; }

  MSIM_HD u8& t_kind(int t) const { return b(L::B_T_KIND, t); }
  MSIM_HD u8& t_pc(int t) const { return b(L::B_T_PC, t); }
  MSIM_HD u8& t_flags(int t) const { return b(L::B_T_FLAGS, t); }
  MSIM_HD u8& t_node(int t) const { return b(L::B_T_NODE, t); }
  MSIM_HD u8& t_epoch(int t) const { return b(L::B_T_EPOCH, t); }
  MSIM_HD u8& t_wait(int t) const { return b(L::B_T_WAIT, t); }
  MSIM_HD u8& t_rx(int t) const { return b(L::B_T_RX, t); }
  MSIM_HD u8& t_sock(int t) const { return b(L::B_T_SOCK, t); }
  MSIM_HD u8& t_sockgen(int t) const { return b(L::B_T_SOCKGEN, t); }
  MSIM_HD u8& t_msg(int t) const { return b(L::B_T_MSG, t); }
  MSIM_HD u8& t_conn(int t) const { return b(L::B_T_CONN, t); }
  MSIM_HD u8& t_conn2(int t) const { return b(L::B_T_CONN2, t); }
  MSIM_HD u8& t_connhalf(int t) const { return b(L::B_T_CONNHALF, t); }
  MSIM_HD u8& t_handle(int t) const { return b(L::B_T_HANDLE, t); }
  MSIM_HD u32& t_spawn(int t) const { return w(L::W_T_SPAWN, t); }
  MSIM_HD u32& t_aw(int t) const { return w(L::W_T_AW, t); }
  MSIM_HD u32& t_a(int t) const { return w(L::W_T_A, t); }
  MSIM_HD u32& t_b(int t) const { return w(L::W_T_B, t); }
  MSIM_HD u32& t_c(int t) const { return w(L::W_T_C, t); }
  MSIM_HD u64& t_dl0(int t) const { return q(L::Q_T_DL0, t); }
  MSIM_HD u64& t_dl1(int t) const { return q(L::Q_T_DL1, t); }
  MSIM_HD u64& t_dl2(int t) const { return q(L::Q_T_DL2, t); }
  MSIM_HD u32 t_gen(int t) const { return t_spawn(t) & 0xfffffu; }
  MSIM_HD u8& n_flags(int n) const { return b(L::B_N_FLAGS, n); }
  MSIM_HD u8& n_epoch(int n) const { return b(L::B_N_EPOCH, n); }
  MSIM_HD u64& wl_q(int f, int n) const { return q(L::Q_WL, f * C::NNODE + n); }
  MSIM_HD u32& wl_w(int f, int n) const { return w(L::W_WL, f * C::NNODE + n); }
  MSIM_HD u32& wl_s(int i) const { return w(L::W_WLS, i); }

  // The per-program code-placement choices (Caps), or what a kernel-experiment build puts in their place (-DMSIM_X_<FLAG>=0/1)
  static constexpr bool X_SPAWN = MSIM_X_SPAWN < 0 ? (bool)C::COLD_SPAWN : (MSIM_X_SPAWN != 0);
  static constexpr bool X_FINISH = MSIM_X_FINISH < 0 ? (bool)C::COLD_FINISH : (MSIM_X_FINISH != 0);
  static constexpr bool X_CH = MSIM_X_CH < 0 ? (bool)C::COLD_CH : (MSIM_X_CH != 0);
  static constexpr bool X_NOW = MSIM_X_NOW < 0 ? caps_cold_now_smem<C>::value : (MSIM_X_NOW != 0);
  static constexpr bool X_BIND = MSIM_X_BIND < 0 ? caps_bind_inline<C>::value : (MSIM_X_BIND != 0);
  static constexpr bool X_PORT_FIRST = MSIM_X_PORTFIRST < 0 ? caps_port_first<C>::value : (MSIM_X_PORTFIRST != 0);

  // ---------------- rarely executed runtime paths ----------------
  /// kill / restart / create_node / bind / panic handling ... run OUT OF LINE: one copy per kernel keeps
  /// them off the hot loop's instruction-cache footprint.  The hot loop's registers are spilled to the
  /// state block around the call (store / load); the callee works on an Engine of its own.
  /// Which register-cached fields a path needs: G_RNG r0-r3 + draws; G_QUEUES rq_len, heap_len, status, peak_timers;
  /// G_CLOCK hash + event counters (`now` always travels as an argument).  A light call moves only its groups.
  enum : int { G_RNG = 1, G_QUEUES = 2, G_CLOCK = 4, G_ALL = 7 };
#if !defined(__CUDACC__)
  int dbg_valid = G_ALL;  // host builds (tests/enginecheck): which groups this Engine holds — a light callee that reached a
                          // full cold() would write garbage back, so store_g checks
#endif
  template <int G, class F>
  static MSIM_NI auto cold_run(unsigned char* sb_, const Params* P_, u64 now_, F f) {
    Engine e_;
    e_.sb = sb_;
    e_.P = P_;
#if !defined(__CUDACC__)
    e_.dbg_valid = 0;
#endif
    e_.tr = nullptr;  // nothing on these paths appends trace records
    e_.tr_cap = 0;
    e_.tr_n = 0;
    e_.template load_g<G>();
    e_.now = now_;
    if constexpr (std::is_void<decltype(f(e_))>::value) {
      f(e_);
      e_.template store_g<G>();
    } else {
      auto r_ = f(e_);
      e_.template store_g<G>();
      return r_;
    }
  }
  template <int G_, class F>
  MSIM_HD auto cold_g(F f) {
    constexpr int G = MSIM_COLD_LIGHT ? G_ : (int)G_ALL;
    store_g<G>();
    if (X_NOW) q(L::Q_NOW) = now;  // the clock goes through the state block too instead of staying live in registers across the call
    if constexpr (std::is_void<decltype(cold_run<G>(sb, P, now, f))>::value) {
      cold_run<G>(sb, P, now, f);
      load_g<G>();
      if (X_NOW) now = q(L::Q_NOW);
    } else {
      auto r_ = cold_run<G>(sb, P, now, f);
      load_g<G>();
      if (X_NOW) now = q(L::Q_NOW);
      return r_;
    }
  }
  template <class F>
  MSIM_HD auto cold(F f) {
    return cold_g<G_ALL>(f);
  }

  // ---------------- failure ----------------
  MSIM_HD void fail(u32 st, u32 code) {
    if (status == MSIM_ST_RUNNING) {
      status = st;
      w(L::W_FAIL_CODE) = code;
    }
  }
  MSIM_HD void fail_cap(u32 which) { fail(MSIM_ST_FAIL_CAPACITY, which); }
  MSIM_HD bool running() const { return status == MSIM_ST_RUNNING; }

  // ---------------- trace ----------------
  MSIM_HD static u64 mix(u64 h, u64 x) {
    h ^= x;
    h *= MSIM_HASH_MUL;
    h ^= h >> 32;
    return h;
  }
  MSIM_HD void trace(u64 t_now, u64 aux, u32 kind, u32 id24) {
    u32 kind_id = (kind << 24) | (id24 & 0xffffffu);
    u32 d = (u32)draws;
    u64 kd = ((u64)kind_id << 32) | d;
    hash = mix(hash, t_now ^ rotl64(aux, 17) ^ rotl64(kd, 41));
    if (TRACE && tr) {
      if (tr_n < tr_cap) {
        tr[tr_n].now_ns = t_now;
        tr[tr_n].aux = aux;
        tr[tr_n].kind_id = kind_id;
        tr[tr_n].draws = d;
      }
      tr_n++;
    }
  }

  // ---------------- RNG (rand.rs + rand 0.8 + rand_xoshiro 0.6) ----------------
  MSIM_HD u64 next_u64() {
    draws++;
    u64 result = rotl64(r0 + r3, 23) + r0;
    u64 t = r1 << 17;
    r2 ^= r0;
    r3 ^= r1;
    r1 ^= r2;
    r0 ^= r3;
    r2 ^= t;
    r3 = rotl64(r3, 45);
    return result;
  }
  /// The determinism log / check of GlobalRng::with (rand.rs:64-89): after EVERY call that goes through the global
  /// RNG — one gen_range, one gen_bool, one random() — the reference records
  ///     (next u8 of a CLONE of the rng) ^ (xor of the bytes of elapsed nanos as u128)
  /// and Runtime::check_determinism compares the second run's bytes with the first run's (runtime/mod.rs:178-202).
  /// `Standard` for u8 is `next_u32() as u8`, and next_u32 is `next_u64() >> 32`.  Compiled only into the replay
  /// kernel (TRACE); the batch kernels carry none of the code.  The log lives behind Params::dlog (single-seed
  /// replay): u32 count, u32 capacity, then the bytes.
  MSIM_HD void log_point() {
    if (TRACE && P->dlog) {
      u32* hdr = reinterpret_cast<u32*>(P->dlog);
      u64 peek = rotl64(r0 + r3, 23) + r0;  // what the clone's next_u64 would return
      u64 f = now ^ (now >> 32);
      f ^= f >> 16;
      f ^= f >> 8;
      u32 i = hdr[0];
      if (i < hdr[1]) reinterpret_cast<u8*>(hdr + 2)[i] = (u8)(peek >> 32) ^ (u8)f;
      hdr[0] = i + 1;
    }
  }
  /// random::<u64>() / one raw draw that is a whole `with` call of its own
  MSIM_HD u64 random_u64() {
    u64 v = next_u64();
    log_point();
    return v;
  }
  /// Rng::gen_bool(p) with p pre-digested to p_int = (p * 2^64) as u64 (rand 0.8 Bernoulli; p == 1.0 never gets here)
  MSIM_HD bool gen_bool_int(u64 p_int) {
    bool r = next_u64() < p_int;
    log_point();
    return r;
  }
  /// Rng::gen_range(lo..hi) for u64/usize: UniformInt::sample_single (widening multiply + zone)
  MSIM_HD u64 gen_range_u64(u64 lo, u64 hi) {
    u64 range = hi - lo;
    u64 zone = (range << clz64(range)) - 1;
    for (;;) {
      u64 v = next_u64();
      u64 l = v * range;
      if (l <= zone) {
        log_point();
        return lo + mulhi64(v, range);
      }
    }
  }
  /// the same sampler for ranges that fit 32 bits (ready-queue index, 50..100 ns, 0..5 us): the
  /// 64 x 64 widening multiply degenerates to two 32 x 32 products
  MSIM_HD u32 gen_range_small(u32 lo, u32 hi) {
    u32 range = hi - lo;
    u64 zone = ((u64)range << (32 + clz32(range))) - 1;
    for (;;) {
      u64 v = next_u64();
      u64 p_lo = (u64)(u32)v * range;
      u64 p_hi = (v >> 32) * range;
      u64 l = p_lo + (p_hi << 32);            // low 64 bits of v * range
      u64 h = (p_hi + (p_lo >> 32)) >> 32;    // high 64 bits (fits 32)
      if (l <= zone) {
        log_point();
        return lo + (u32)h;
      }
    }
  }
  /// Uniform::<u64>::new_inclusive(lo, hi).sample — modulo zone (UniformDuration Medium)
  MSIM_HD u64 sample_inclusive_u64(u64 lo, u64 hi) {
    u64 range = hi - lo + 1;
    u64 zone = ~0ull - ((~0ull - range + 1) % range);
    for (;;) {
      u64 v = next_u64();
      u64 l = v * range;
      if (l <= zone) {
        log_point();
        return lo + mulhi64(v, range);
      }
    }
  }
  /// gen_range(config.send_latency): sampler pre-digested on the host (Params.lat_*)
  MSIM_HD u64 sample_latency() {
    if (P->lat_mode == 0) {  // Small: Uniform<u32> over sub-second nanos; next_u32 = next_u64 >> 32
      u32 range = (u32)P->lat_range, zone = (u32)P->lat_zone;
      for (;;) {
        u32 v = (u32)(next_u64() >> 32);
        u64 m = (u64)v * range;
        if ((u32)m <= zone) {
          log_point();
          return P->lat_base + (u32)P->lat_lo + (u32)(m >> 32);
        }
      }
    } else {
      u64 range = P->lat_range, zone = P->lat_zone;
      for (;;) {
        u64 v = next_u64();
        u64 l = v * range;
        if (l <= zone) {
          log_point();
          return P->lat_lo + mulhi64(v, range);
        }
      }
    }
  }

  // ---------------- timer heap (naive_timer over Rust's BinaryHeap) ----------------
  struct HRec {
    u64 dl;
    u32 pl;
  };
  MSIM_HD u64& h_dl(int i) const { return q(L::Q_HEAP_DL, L::H16 ? 2 * i : i); }
  MSIM_HD HRec h_get(int i) const {
    if constexpr (L::H16) {
#if defined(__CUDA_ARCH__)
      ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&q(L::Q_HEAP_DL, 2 * i));
      return HRec{v.x, (u32)v.y};
#else
      return HRec{q(L::Q_HEAP_DL, 2 * i), (u32)q(L::Q_HEAP_DL, 2 * i + 1)};
#endif
    } else {
      return HRec{q(L::Q_HEAP_DL, i), w(L::W_HEAP_PL, i)};
    }
  }
  MSIM_HD void h_set(int i, u64 dl, u32 pl) const {
    if constexpr (L::H16) {
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<ulonglong2*>(&q(L::Q_HEAP_DL, 2 * i)) = make_ulonglong2(dl, (u64)pl);
#else
      q(L::Q_HEAP_DL, 2 * i) = dl, q(L::Q_HEAP_DL, 2 * i + 1) = pl;
#endif
    } else {
      q(L::Q_HEAP_DL, i) = dl, w(L::W_HEAP_PL, i) = pl;
    }
  }
  /// entry `from` moves to slot `to`; its deadline is already known to the caller
  MSIM_HD void h_move(int to, int from, u64 dl_from) const {
    if constexpr (L::H16) {
      HRec r = h_get(from);
      h_set(to, r.dl, r.pl);
    } else {
      q(L::Q_HEAP_DL, to) = dl_from, w(L::W_HEAP_PL, to) = w(L::W_HEAP_PL, from);
    }
  }
  /// Timer::add = BinaryHeap::push: append, sift_up while STRICTLY earlier than the parent
  MSIM_HD void timer_push(u64 dl, u32 pl) {
    u32 pos = heap_len;
    if (MSIM_UNLIKELY(pos >= (u32)C::TQ)) {
      fail_cap(MSIM_CAP_TIMERS);
      return;
    }
    heap_len = pos + 1;
    if (heap_len > peak_timers) peak_timers = heap_len;
    while (pos > 0) {
      u32 parent = (pos - 1) >> 1;
      if constexpr (L::H16) {
        HRec pr = h_get(parent);
        if (dl >= pr.dl) break;
        h_set(pos, pr.dl, pr.pl);
      } else {
        u64 pd = h_dl(parent);
        if (dl >= pd) break;
        h_move(pos, parent, pd);
      }
      pos = parent;
    }
    h_set(pos, dl, pl);
  }
  /// BinaryHeap::pop: move the last element to the root, sift_down_to_bottom (on equal
  /// deadlines the RIGHT child is promoted), then sift_up from the leaf.
  MSIM_HD void timer_pop(u64& dl, u32& pl) {
    HRec root = h_get(0);
    dl = root.dl;
    pl = root.pl;
    u32 end = --heap_len;
    if (end == 0) return;
    HRec e = h_get(end);
    u32 pos = 0, child = 1;
    while (end >= 2 && child <= end - 2) {
      if constexpr (L::H16) {
        HRec l = h_get(child), r = h_get(child + 1);
#if defined(__CUDA_ARCH__) && MSIM_HEAP_ASM_SEL
        u64 m_dl;  // ONE 64-bit compare feeds all three selects (the compiler re-derives the deadline select as a min)
        u32 m_pl, right;
        asm("{\n\t.reg .pred p;\n\tsetp.ge.u64 p, %3, %4;\n\tselp.b64 %0, %4, %3, p;\n\tselp.b32 %1, %6, %5, p;\n\tselp.u32 %2, 1, 0, p;\n\t}"
            : "=l"(m_dl), "=r"(m_pl), "=r"(right)
            : "l"(l.dl), "l"(r.dl), "r"(l.pl), "r"(r.pl));
        child += right;
        h_set(pos, m_dl, m_pl);
#else
        bool right = l.dl >= r.dl;
        child += right;
        h_set(pos, right ? r.dl : l.dl, right ? r.pl : l.pl);
#endif
      } else {
        u64 dl_l = h_dl(child), dl_r = h_dl(child + 1);
        if (dl_l >= dl_r) {
          child++;
          dl_l = dl_r;
        }
        h_move(pos, child, dl_l);
      }
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      h_move(pos, child, h_dl(child));
      pos = child;
    }
    while (pos > 0) {
      u32 parent = (pos - 1) >> 1;
      if constexpr (L::H16) {
        HRec pr = h_get(parent);
        if (e.dl >= pr.dl) break;
        h_set(pos, pr.dl, pr.pl);
      } else {
        u64 pd = h_dl(parent);
        if (e.dl >= pd) break;
        h_move(pos, parent, pd);
      }
      pos = parent;
    }
    h_set(pos, e.dl, e.pl);
  }
  /// Timer::expire(limit): fire every timer with deadline <= limit, in heap order.
  /// `now` is NOT touched: advance_to_next_event fires callbacks under the old clock.
  MSIM_HD bool timer_due(u64 limit) const { return heap_len > 0 && h_dl(0) <= limit && running(); }
  /// one `pop; callback(now)` of naive_timer::Timer::expire
  MSIM_HD void fire_one() {
    u64 dl;
    u32 pl;
    timer_pop(dl, pl);
    ev_fire++;
    u32 kind = pl >> 28, a = (pl >> 20) & 0xff, g = pl & 0xfffffu;
    u32 id = kind == TK_WAKE ? g : (kind == TK_DELIVER ? w(L::W_M_SERIAL, a) : a);
    trace(now, dl, MSIM_EV_WAKE + kind, id);  // MSIM_EV_WAKE / DELIVER / RESTART = 5 + TK_*
    if (kind == TK_WAKE)
      wake(a, g);
    else if (MSIM_LIKELY(kind == TK_DELIVER))
      deliver(a);
    else
      restart(a);
  }
  MSIM_HD void expire(u64 limit) {
    while (timer_due(limit)) fire_one();
  }
  /// sleep()/sleep_until(): deadline = max(deadline, now + 1 ms)  (time/mod.rs:111-124)
  MSIM_HD u64 sleep_deadline(u64 deadline) const {
    u64 m = now + NS_PER_MS;
    return deadline > m ? deadline : m;
  }

  // ---------------- tasks (async-task state machine) ----------------
  MSIM_HD void rq_push(int t) {
    if (MSIM_UNLIKELY(rq_len >= (u32)C::NT)) {
      fail_cap(MSIM_CAP_READY);
      return;
    }
    b(L::B_RQ, rq_len++) = (u8)t;
  }
  /// Waker::wake for the task that lived in slot `t` with generation `gen`
  MSIM_HD void wake(int t, u32 gen) {
    u32 sp = t_spawn(t);
    if ((sp & 0xfffffu) != gen || sp == NONE32) return;
    u8 f = t_flags(t);
    if (f & (TF_COMPLETED | TF_CLOSED | TF_SCHEDULED)) return;
    t_flags(t) = f | TF_SCHEDULED;
    if (!(f & TF_RUNNING)) rq_push(t);
  }
  MSIM_HD void wake_task(int t) { wake(t, t_gen(t)); }
  MSIM_HD void notify_awaiter(int t) {
    if (!C::USES_JOIN) return;  // nobody ever awaits a JoinHandle in this program
    u32 aw = t_aw(t);
    if (aw != NONE32) {
      t_aw(t) = NONE32;
      wake(aw >> 20, aw & 0xfffffu);
    }
  }
  MSIM_HD void maybe_free(int t) {
    u8 f = t_flags(t);
    if ((f & (TF_ALIVE | TF_JOINABLE | TF_SCHEDULED | TF_RUNNING)) == 0 && (f & (TF_COMPLETED | TF_CLOSED))) {
      t_flags(t) = 0;
      t_spawn(t) = NONE32;
      q(L::Q_TFREE, t >> 6) |= 1ull << (t & 63);
    }
  }
  /// Dropping the task's future (and the TaskInfo inside it): whatever the suspended future holds
  /// goes with it — oneshot receiver (registered or filled), prepared message, Endpoint clone.
  MSIM_HD void drop_future(int t) {
    u8 f = t_flags(t);
    if (f & TF_ALIVE) {
      u8 st = t_wait(t) & 0x0f;
      if (st == W_RECV_RX || st == W_CALL_RX) mbox_unregister(t);
      if (t_rx(t) != NONE8) {
        msg_free(t_rx(t));
        t_rx(t) = NONE8;
      }
      if (t_msg(t) != NONE8) {
        msg_free(t_msg(t));
        t_msg(t) = NONE8;
      }
      if (C::USES_JOIN && st == W_JOIN) join_detach((int)t_c(t));
      if (C::USES_SPLIT && t_handle(t) != NONE8) {  // `_task` drops before the stream's Receiver
        if (t_handle(t) & 0x80)
          join_detach(t_handle(t) & 0x7f);  // still a plain JoinHandle (the stream was never polled): detach
        else
          task_cancel_detach(t_handle(t));
        t_handle(t) = NONE8;
      }
      if (C::NCONN) conn_drop_end(t);
      if (C::USES_CONN2 && t_conn2(t) != NONE8) {  // the parked end: declared earlier, dropped later
        conn_swap(t);
        conn_drop_end(t);
      }
      sock_unref(t);
      t_wait(t) = W_NONE;
      t_flags(t) = f & ~TF_ALIVE;
      w(L::W_LIVE_TASKS)--;
    }
  }
  /// impl Drop for Runnable: the kill / cancel path (task/mod.rs:273)
  MSIM_HD void drop_runnable_inl(int t) {
    u8 f = t_flags(t);
    if (!(f & (TF_COMPLETED | TF_CLOSED))) t_flags(t) = f | TF_CLOSED;
    drop_future(t);
    t_flags(t) = t_flags(t) & ~TF_SCHEDULED;
    notify_awaiter(t);
    maybe_free(t);
  }
  MSIM_HD void drop_runnable(int t) {
    cold([t](Engine& e_) { e_.drop_runnable_inl(t); });
  }
  /// Spawner::spawn_inner (task/mod.rs:627-654): returns the slot or -1.  Programs whose hot loop is larger than the
  /// SM's instruction cache (profiles/r2_etcd, r2_tonic_unary: 63-67 % hit rate) run the once-per-request pieces —
  /// spawn, the end of a task, the link test of a channel send, connect1 — out of line (Caps::COLD_*).
  MSIM_HD int spawn(u32 kind, u32 node, u8 extra_flags) {
    if (X_SPAWN) return cold_g<G_QUEUES>([kind, node, extra_flags](Engine& e_) { return e_.spawn_inl(kind, node, extra_flags); });
    return spawn_inl(kind, node, extra_flags);
  }
  MSIM_HD int spawn_inl(u32 kind, u32 node, u8 extra_flags) {
    int t = -1;
MSIM_NOUNROLL
    for (int wd = 0; wd < (C::NT + 63) / 64; wd++) {
      u64 m = q(L::Q_TFREE, wd);
      if (m) {
        int bit = ffs64(m);
        q(L::Q_TFREE, wd) = m & (m - 1);
        t = wd * 64 + bit;
        break;
      }
    }
    u32 idx = w(L::W_NEXT_SPAWN);
    if (MSIM_UNLIKELY(t < 0 || t >= C::NT || idx >= 0xfffffu)) {
      fail_cap(MSIM_CAP_TASKS);
      return -1;
    }
    w(L::W_NEXT_SPAWN) = idx + 1;
    t_kind(t) = (u8)kind;
    t_pc(t) = 0;
    t_wait(t) = W_NONE;
    t_flags(t) = TF_SCHEDULED | TF_ALIVE | extra_flags;
    t_node(t) = (u8)node;
    t_epoch(t) = n_epoch(node);
    t_rx(t) = NONE8;
    t_sock(t) = NONE8;
    t_msg(t) = NONE8;
    if (C::NCONN) t_conn(t) = NONE8;
    if (C::USES_CONN2) t_conn2(t) = NONE8;
    if (C::USES_SPLIT) t_connhalf(t) = 3, t_handle(t) = NONE8;
    if (C::USES_TIMEOUT) t_dl2(t) = NONE64;
    t_spawn(t) = idx;
    t_aw(t) = NONE32;
    u32 live = ++w(L::W_LIVE_TASKS);
    if (live > w(L::W_PEAK_TASKS)) w(L::W_PEAK_TASKS) = live;
    rq_push(t);
    return t;
  }
  /// JoinHandle::poll (async_task::Task::poll_task).  0 pending, 1 Ok (output = t_a(target)), 2 cancelled
  MSIM_HD int join_poll(int self, int target) {
    u8 f = t_flags(target);
    if (f & TF_CLOSED) {
      if (f & (TF_SCHEDULED | TF_RUNNING)) {
        t_aw(target) = ((u32)self << 20) | t_gen(self);
        return 0;
      }
      t_flags(target) = f & ~TF_JOINABLE;
      return 2;
    }
    if (!(f & TF_COMPLETED)) {
      t_aw(target) = ((u32)self << 20) | t_gen(self);
      return 0;
    }
    t_flags(target) = (f | TF_CLOSED) & ~TF_JOINABLE;
    return 1;  // caller reads t_a(target) then maybe_free(target)
  }
  /// Dropping a JoinHandle = detach
  MSIM_HD void join_detach(int target) {
    t_flags(target) = t_flags(target) & ~TF_JOINABLE;
    t_aw(target) = NONE32;
    maybe_free(target);
  }
  /// Drop of the FallibleTask made by JoinHandle::cancel_on_drop (task/join.rs:54-56; async-task Task::drop):
  /// close the task; if it is neither scheduled nor running, schedule it so that the executor drops its future
  MSIM_HD void task_cancel_detach(int x) {
    u8 f = t_flags(x);
    if (!(f & (TF_COMPLETED | TF_CLOSED))) {
      f |= TF_CLOSED;
      if (!(f & (TF_SCHEDULED | TF_RUNNING))) {
        t_flags(x) = f | TF_SCHEDULED;
        rq_push(x);
      } else {
        t_flags(x) = f;
      }
    }
    join_detach(x);
  }
  /// AbortHandle::abort (task/join.rs:158-163)
  MSIM_HD void abort_task(int t) {
    if (t_flags(t) & TF_ALIVE) {
      t_flags(t) = t_flags(t) | TF_CANCELLED;
      wake_task(t);
    }
  }

  // ---------------- nodes: kill / restart / pause / resume (task/mod.rs:356-424) ----------------
  MSIM_HD bool task_node_killed(int t) const {
    int n = t_node(t);
    return t_epoch(t) != n_epoch(n) || (n_flags(n) & NF_KILLED);
  }
  /// NodeInfo::kill for the incarnation (n, epoch): wake every live task in spawn order
  MSIM_HD void info_kill(int n, u8 epoch) {
    u32 last = 0;
    bool first = true;
    for (;;) {
      u32 best = NONE32;
      int bt = -1;
MSIM_NOUNROLL
      for (int t = 0; t < C::NT; t++) {
        u32 sp = t_spawn(t);
        if (sp == NONE32 || !(t_flags(t) & TF_ALIVE)) continue;
        if (t_node(t) != n || t_epoch(t) != epoch) continue;
        if (!first && sp <= last) continue;
        if (sp < best) {
          best = sp;
          bt = t;
        }
      }
      if (bt < 0) break;
      wake_task(bt);
      last = best;
      first = false;
    }
  }
  MSIM_HD void drop_parked(int n) {
    u32 len = w(L::W_PK_LEN), keep = 0;
MSIM_NOUNROLL
    for (u32 i = 0; i < len; i++) {
      int t = b(L::B_PK, i);
      if (t_node(t) == n) {
        drop_runnable(t);
      } else {
        b(L::B_PK, keep++) = (u8)t;
      }
    }
    w(L::W_PK_LEN) = keep;
  }
  MSIM_HD void kill_inl(int n) {
    drop_parked(n);
    n_flags(n) = n_flags(n) | NF_KILLED;
    info_kill(n, n_epoch(n));
    reset_node_sockets(n);  // NetSim::reset_node: all sockets of the node closed (network.rs:142-147)
  }
  MSIM_HD void kill(int n) {
    cold([n](Engine& e_) { e_.kill_inl(n); });
  }
  /// Spawner::exit (task/mod.rs:656-661): info.kill() only — no reset_node, paused list kept
  MSIM_HD void node_exit_inl(int n, u8 epoch) {
    if (n_epoch(n) == epoch) n_flags(n) = n_flags(n) | NF_KILLED;
    info_kill(n, epoch);
  }
  MSIM_HD void node_exit(int n, u8 epoch) {
    cold([n, epoch](Engine& e_) { e_.node_exit_inl(n, epoch); });
  }
  MSIM_HD void restart_inl(int n) {
    drop_parked(n);
    info_kill(n, n_epoch(n));  // old_info.kill()
    n_epoch(n) = n_epoch(n) + 1;
    n_flags(n) = n_flags(n) & ~(NF_KILLED | NF_PAUSED);
    if (C::USES_CTRLC) w(L::W_N_CTRLC, n) = 0, w(L::W_N_CTRLC_WAITER, n) = NONE32;  // the new NodeInfo has no handler
    int kind = WL::init_kind(n);
    if (kind >= 0) spawn_inl((u32)kind, n, TF_INIT);
  }
  MSIM_HD void restart(int n) {
    cold([n](Engine& e_) { e_.restart_inl(n); });
  }
  /// Handle::send_ctrl_c (task/mod.rs:427-441): watch::Sender::send if a handler was ever installed
  /// on this NodeInfo (an error for "no receiver" is ignored), otherwise the node is killed
  MSIM_HD void send_ctrl_c_inl(int n) {
    if (!C::USES_CTRLC) {
      kill_inl(n);
      return;
    }
    u32 c = w(L::W_N_CTRLC, n);
    if (c >> 31) {
      w(L::W_N_CTRLC, n) = 0x80000000u | ((c + 1) & 0x7fffffffu);
      wake_waiter(w(L::W_N_CTRLC_WAITER, n));
    } else {
      kill_inl(n);
    }
  }
  MSIM_HD void send_ctrl_c(int n) {
    cold([n](Engine& e_) { e_.send_ctrl_c_inl(n); });
  }
  MSIM_HD void pause(int n) { n_flags(n) = n_flags(n) | NF_PAUSED; }
  MSIM_HD void resume_inl(int n) {
    n_flags(n) = n_flags(n) & ~NF_PAUSED;
    u32 len = w(L::W_PK_LEN), keep = 0;
MSIM_NOUNROLL
    for (u32 i = 0; i < len; i++) {
      int t = b(L::B_PK, i);
      if (t_node(t) == n)
        rq_push(t);
      else
        b(L::B_PK, keep++) = (u8)t;
    }
    w(L::W_PK_LEN) = keep;
  }
  MSIM_HD void resume(int n) {
    cold([n](Engine& e_) { e_.resume_inl(n); });
  }
  MSIM_HD bool is_exit(int n) const { return n_flags(n) & NF_KILLED; }
  /// Handle::create_node + NodeBuilder::build (runtime/mod.rs:405-418): ids start at 1; the init
  /// closure runs (spawns the init task) at build time
  MSIM_HD int create_node_inl(u32 flags, u32 match_mask) {
    u32 n = w(L::W_NEXT_NODE);
    if (n >= (u32)C::NNODE) {
      fail_cap(MSIM_CAP_TASKS);
      return -1;
    }
    w(L::W_NEXT_NODE) = n + 1;
    n_flags(n) = NF_EXISTS | ((flags & NODE_IP) ? NF_HAS_IP : 0) | ((flags & NODE_RESTART_ON_PANIC) ? NF_RESTART_ON_PANIC : 0);
    n_epoch(n) = 0;
    if (C::USES_CTRLC) w(L::W_N_CTRLC, n) = 0, w(L::W_N_CTRLC_WAITER, n) = NONE32;
    if (C::USES_PMATCH) w(L::W_N_PMATCH, n) = (flags & NODE_PANIC_MATCHING) ? match_mask : 0u;  // NodeBuilder::restart_on_panic_matching
    int kind = (flags & NODE_INIT) ? WL::init_kind((int)n) : -1;
    if (kind >= 0) spawn_inl((u32)kind, n, TF_INIT);
    return (int)n;
  }
  MSIM_HD int create_node(u32 flags, u32 match_mask) {
    return cold([flags, match_mask](Engine& e_) { return e_.create_node_inl(flags, match_mask); });
  }

  // ---------------- network (net/network.rs, net/mod.rs, net/endpoint.rs) ----------------
  MSIM_HD u8& sk_state(int k) const { return b(L::B_SK_STATE, k); }
  MSIM_HD u32 sk_gen(int k) const { return w(L::W_SK_PORTGEN, k) >> 16; }
  MSIM_HD u32 sk_port(int k) const { return w(L::W_SK_PORTGEN, k) & 0xffff; }
  MSIM_HD int msg_alloc() {  // the lowest free slot
    if constexpr (L::MW == 1) {
      u64 m = q(L::Q_MFREE);
      if (MSIM_UNLIKELY(!m)) {
        fail_cap(MSIM_CAP_MSGS);
        return -1;
      }
      int i = ffs64(m);
      q(L::Q_MFREE) = m & (m - 1);
      return i;
    } else {  // the large-capacity variants
      for (int wd = 0; wd < L::MW; wd++) {
        u64 m = q(L::Q_MFREE, wd);
        if (m) {
          q(L::Q_MFREE, wd) = m & (m - 1);
          return wd * 64 + ffs64(m);
        }
      }
      fail_cap(MSIM_CAP_MSGS);
      return -1;
    }
  }
  MSIM_HD void msg_free(int m) { q(L::Q_MFREE, L::MW == 1 ? 0 : m >> 6) |= 1ull << (m & 63); }
  MSIM_HD int sock_find(int node, u32 ipk, u32 port) const {
MSIM_NOUNROLL
    for (int k = 0; k < C::NSOCK; k++)
      if (X_PORT_FIRST) {  // same conjunction, port tested first
        if (sk_port(k) == port && sk_state(k) == 1 && b(L::B_SK_NODE, k) == node && b(L::B_SK_IPK, k) == ipk) return k;
      } else {
        if (sk_state(k) == 1 && b(L::B_SK_NODE, k) == node && b(L::B_SK_IPK, k) == ipk && sk_port(k) == port) return k;
      }
    return -1;
  }
  /// Network::bind (network.rs:206-251).  >= 0 socket slot; -1 AddrNotAvailable; -2 AddrInUse
  MSIM_HD int sock_bind_inl(int node, u32 addr) {
    u32 ipk = addr_ipk(addr), port = addr_port(addr);
    if (ipk == IPK_NODE && (int)addr_node(addr) != node) return -1;
    if (port == 0) {
      port = 1;
      while (port <= 0xffff && sock_find(node, ipk, port) >= 0) port++;
      if (port > 0xffff) return -2;
    } else if (sock_find(node, ipk, port) >= 0) {
      return -2;
    }
    int k = 0;
    while (k < C::NSOCK && sk_state(k)) k++;
    if (k >= C::NSOCK) {
      fail_cap(MSIM_CAP_SOCKETS);
      return -2;
    }
    sk_state(k) = 1;
    b(L::B_SK_NODE, k) = (u8)node;
    b(L::B_SK_IPK, k) = (u8)ipk;
    b(L::B_SK_REGLEN, k) = 0;
    b(L::B_SK_MBLEN, k) = 0;
    b(L::B_SK_REFS, k) = 1;
    b(L::B_SK_EPOCH, k) = n_epoch(node);
    w(L::W_SK_PORTGEN, k) = port | ((sk_gen(k) + 1) << 16);
    return k;
  }
  /// out of line for the programs that bind once per node; inline (Caps::BIND_INLINE) where every client call binds an
  /// ephemeral endpoint: TONIC +6 %, ELECTION +3 %, but ETCD -7 % (profiles/r2w_bind_modes.txt)
  MSIM_HD int sock_bind(int node, u32 addr) {
    if constexpr (X_BIND) return sock_bind_inl(node, addr);
    return cold([node, addr](Engine& e_) { return e_.sock_bind_inl(node, addr); });
  }
  MSIM_HD void sock_release_storage(int k) {
    u32 ml = b(L::B_SK_MBLEN, k);
MSIM_NOUNROLL
    for (u32 i = 0; i < ml; i++) msg_free(b(L::B_MB, k * C::NMBOX + i));
    b(L::B_SK_MBLEN, k) = 0;
    b(L::B_SK_REGLEN, k) = 0;
    if (C::NCONN) {  // pending incoming connections die with the socket object: their (tx, rx) drop, FIFO
      u32 al = b(L::B_ACC_LEN, k);
MSIM_NOUNROLL
      for (u32 i = 0; i < al; i++) conn_drop(2 * b(L::B_ACC, k * C::NACC + i) + 1);
      b(L::B_ACC_LEN, k) = 0;
      w(L::W_ACC_WAITER, k) = NONE32;
    }
    sk_state(k) = 0;
    w(L::W_SK_PORTGEN, k) = (w(L::W_SK_PORTGEN, k) & 0xffff0000u) + 0x10000u;  // gen++ : stale refs miss
  }
  /// NetSim::reset_node (network.rs:142-147): every socket of the node leaves the address map.  The socket
  /// OBJECT lives on until the killed futures that hold it are dropped — observable only through the
  /// connections still waiting in its accept queue (they close then, not now): such a socket stays behind
  /// unbound (state 2) until its last reference goes (sock_unref).
  MSIM_HD void reset_node_sockets(int n) {
MSIM_NOUNROLL
    for (int k = 0; k < C::NSOCK; k++)
      if (sk_state(k) == 1 && b(L::B_SK_NODE, k) == n) {
        if (C::NCONN && b(L::B_ACC_LEN, k) != 0) {
          u32 ml = b(L::B_SK_MBLEN, k);
          for (u32 i = 0; i < ml; i++) msg_free(b(L::B_MB, k * C::NMBOX + i));
          b(L::B_SK_MBLEN, k) = 0;
          b(L::B_SK_REGLEN, k) = 0;
          w(L::W_ACC_WAITER, k) = NONE32;
          sk_state(k) = 2;
        } else {
          sock_release_storage(k);
        }
      }
  }
  /// A future takes / drops a clone of the Endpoint (Arc<BindGuard>)
  MSIM_HD void sock_ref(int t, int k) {
    t_sock(t) = (u8)k;
    t_sockgen(t) = (u8)sk_gen(k);
    b(L::B_SK_REFS, k)++;
  }
  MSIM_HD void sock_adopt(int t, int k) {  // the binder: REFS already 1
    t_sock(t) = (u8)k;
    t_sockgen(t) = (u8)sk_gen(k);
  }
  /// BindGuard::drop (net/mod.rs:483-493): close unless the binding incarnation was killed
  MSIM_HD void sock_unref(int t) {
    int k = t_sock(t);
    if (k == NONE8) return;
    t_sock(t) = NONE8;
    if (!sk_state(k) || (u8)sk_gen(k) != t_sockgen(t)) return;
    if (--b(L::B_SK_REFS, k) != 0) return;
    if (C::NCONN && sk_state(k) == 2) {  // the last holder of a reset socket: its accept queue closes now
      sock_release_storage(k);
      return;
    }
    int n = b(L::B_SK_NODE, k);
    bool killed = b(L::B_SK_EPOCH, k) != n_epoch(n) || (n_flags(n) & NF_KILLED);
    if (!killed) sock_release_storage(k);
  }
  MSIM_HD bool link_clogged(int src, int dst) const {
    return ((w(L::W_CLOG_OUT) >> src) & 1) || ((w(L::W_CLOG_IN) >> dst) & 1) || ((w(L::W_CLOG_LINK, src) >> dst) & 1);
  }
  MSIM_HD void clog_node(int n) {
    w(L::W_CLOG_IN) |= 1u << n;
    w(L::W_CLOG_OUT) |= 1u << n;
  }
  MSIM_HD void unclog_node(int n) {
    w(L::W_CLOG_IN) &= ~(1u << n);
    w(L::W_CLOG_OUT) &= ~(1u << n);
  }
  MSIM_HD void clog_link(int sN, int d) { w(L::W_CLOG_LINK, sN) |= 1u << d; }
  MSIM_HD void unclog_link(int sN, int d) { w(L::W_CLOG_LINK, sN) &= ~(1u << d); }

  /// NetSim::rand_delay (net/mod.rs:287-295): returns the Sleep deadline (always >= now + 1 ms)
  MSIM_HD u64 rand_delay_deadline() {
    u64 delay = (u64)gen_range_small(0, 5) * 1000ull;
    if (P->buggify) {
      if (gen_bool_int(P->buggify_p10_int)) delay = (u64)gen_range_small(1, 5) * NS_PER_SEC;
    }
    return sleep_deadline(now + delay);
  }
  /// Network::try_send (network.rs:296-313) with resolve_dest_node (:273-291) and test_link
  /// (:261-270): the destination socket or -1; `latency` is valid on success only.
  MSIM_HD int try_send(int src_node, u32 dst, u64& latency) {
    int dst_node = -1;
    u32 ipk = addr_ipk(dst), dport = addr_port(dst);
    bool own = (ipk != IPK_NODE || (int)addr_node(dst) == src_node) && sock_find(src_node, ipk, dport) >= 0;
    if (ipk == IPK_LOOPBACK || own) {
      dst_node = src_node;
    } else if ((n_flags(src_node) & NF_HAS_IP) && ipk == IPK_NODE && addr_node(dst) < (u32)C::NNODE &&
               (n_flags(addr_node(dst)) & NF_HAS_IP)) {
      dst_node = (int)addr_node(dst);
    }
    // "ip not set" / destination not found: before any draw.  `clogged || gen_bool(loss)`: no
    // draw when clogged
    if (dst_node < 0 || link_clogged(src_node, dst_node) || P->loss_always || gen_bool_int(P->loss_p_int)) return -1;
    ++w(L::W_MSG_COUNT);
    latency = sample_latency();
    // socket lookup AFTER the draws: exact address, then (UNSPECIFIED, port)
    int k = sock_find(dst_node, ipk, dport);
    if (k < 0 && ipk != IPK_UNSPEC) k = sock_find(dst_node, IPK_UNSPEC, dport);
    return k;
  }
  MSIM_HD u32 src_addr_for(u32 dst, int src_node, u32 src_port) const {
    return mk_addr(addr_ipk(dst) == IPK_LOOPBACK ? IPK_LOOPBACK : IPK_NODE, src_node, src_port);
  }
  /// The part of NetSim::send after rand_delay (net/mod.rs:304-332).  `m` is a prepared message
  /// slot (tag, w, v0, v1); consumed.
  MSIM_HD void net_send(int src_node, u32 src_port, u32 dst, int m) {
    u64 latency;
    int k = try_send(src_node, dst, latency);
    if (k < 0) {
      msg_free(m);
      return;
    }
    w(L::W_M_FROM, m) = src_addr_for(dst, src_node, src_port);
    w(L::W_M_DST, m) = ((u32)k << 16) | (sk_gen(k) & 0xffff);
    w(L::W_M_SERIAL, m) = w(L::W_MSG_COUNT) & 0xffffffu;
    timer_push(now + latency, timer_payload(TK_DELIVER, m, 0));
  }

  // ---------------- reliable channels (net/mod.rs:337-430, endpoint.rs:178-274) ----------------
  enum : u8 { CH_TX_ALIVE = 1, CH_RX_ALIVE = 2 };
  static constexpr u64 NONE64 = ~0ull;
  MSIM_HD u8& ch_flags(int d) const { return b(L::B_CH_FLAGS, d); }
  MSIM_HD void wake_waiter(u32& slot_word) {
    u32 wv = slot_word;
    if (wv != NONE32) {  // AtomicWaker::wake takes the waker
      slot_word = NONE32;
      wake(wv >> 20, wv & 0xfffffu);
    }
  }
  /// channel()'s test_link closure: try_send(node, dst).map(|lat| now + lat)
  MSIM_HD u64 ch_test_link(int d) {
    if (X_CH) return cold_g<G_RNG>([d](Engine& e_) { return e_.ch_test_link_inl(d); });
    return ch_test_link_inl(d);
  }
  MSIM_HD u64 ch_test_link_inl(int d) {
    u64 lat;
    int k = try_send(b(L::B_CH_SRC, d), w(L::W_CH_DST, d), lat);
    return k < 0 ? NONE64 : now + lat;
  }
  /// NetSim::connect1 after rand_delay.  0 ok / -1 ConnectionRefused
  MSIM_HD int connect1(int t, u32 dst) {
    if (X_CH) return cold_g<G_RNG | G_QUEUES>([t, dst](Engine& e_) { return e_.connect1_inl(t, dst); });
    return connect1_inl(t, dst);
  }
  MSIM_HD int connect1_inl(int t, u32 dst) {
    int ks = t_sock(t);
    int node = b(L::B_SK_NODE, ks);
    u64 lat;
    int k = try_send(node, dst, lat);  // the latency is drawn but not used (the reference's FIXME)
    if (k < 0) return -1;
    int c = 0;
    while (c < C::NCONN && b(L::B_CONN_USED, c)) c++;
    u32 al = b(L::B_ACC_LEN, k);
    if (c >= C::NCONN || al >= (u32)C::NACC) {
      fail_cap(MSIM_CAP_SOCKETS);
      return -1;
    }
    u32 src = src_addr_for(dst, node, sk_port(ks));
    b(L::B_CONN_USED, c) = 1;
    w(L::W_CONN_SRC, c) = src;
    for (int side = 0; side < 2; side++) {
      int d = 2 * c + side;
      b(L::B_CH_HEAD, d) = 0;
      b(L::B_CH_LEN, d) = 0;
      b(L::B_CH_PHASE, d) = 0;
      ch_flags(d) = CH_TX_ALIVE | CH_RX_ALIVE;
      w(L::W_CH_WAITER, d) = NONE32;
      if (C::USES_CLOSED) w(L::W_CH_CLOSED_WAITER, d) = NONE32;
    }
    b(L::B_CH_SRC, 2 * c) = (u8)node;  // channel(node, dst)
    w(L::W_CH_DST, 2 * c) = dst;
    b(L::B_CH_SRC, 2 * c + 1) = b(L::B_SK_NODE, k);  // channel(dst_node, src)
    w(L::W_CH_DST, 2 * c + 1) = src;
    // socket.new_connection(..): conn_tx.try_send — delivered immediately
    b(L::B_ACC, k * C::NACC + al) = (u8)c;
    b(L::B_ACC_LEN, k) = (u8)(al + 1);
    wake_waiter(w(L::W_ACC_WAITER, k));
    t_conn(t) = (u8)(2 * c);
    if (C::USES_SPLIT) t_connhalf(t) = 3;
    return 0;
  }
  /// Dropping the halves a future owns of one end of a connection (both, unless the Sender was moved away)
  MSIM_HD void conn_drop_end(int t) {
    int e = t_conn(t);
    if (e == NONE8) return;
    t_conn(t) = NONE8;
    conn_drop_half(e, C::USES_SPLIT ? t_connhalf(t) : 3);
  }
  /// A future that holds TWO connection ends (an ObserveStream kept open while another call is made, election.rs:79-86)
  /// works on one at a time: conn_swap exchanges the active end with the parked one.  No await point.
  MSIM_HD void conn_swap(int t) {
    if (!C::USES_CONN2) return;
    u8 a = t_conn(t);
    t_conn(t) = t_conn2(t);
    t_conn2(t) = a;
  }
  /// Sender::closed() inside a hand-written select (a state machine that waits for something the engine does not know,
  /// e.g. a watcher event): true = the peer dropped its Receiver; otherwise this task is the one closed() will wake
  MSIM_HD bool closed_or_register(int t) {
    int d = t_conn(t);
    if (!(ch_flags(d) & CH_RX_ALIVE)) return true;
    if (C::USES_CLOSED) w(L::W_CH_CLOSED_WAITER, d) = ((u32)t << 20) | t_gen(t);
    return false;
  }
  MSIM_HD void closed_unregister(int t) {  // the closed() future is dropped with its select
    if (C::USES_CLOSED && t_conn(t) != NONE8) w(L::W_CH_CLOSED_WAITER, t_conn(t)) = NONE32;
  }
  MSIM_HD u32 waker_of(int t) const { return ((u32)t << 20) | t_gen(t); }
  /// Dropping only the Sender of the end a task holds (a request stream that ended: client.rs:171-196)
  MSIM_HD void conn_drop_tx(int t) {
    if (C::USES_SPLIT) t_connhalf(t) = t_connhalf(t) & ~1;
    conn_drop_half(t_conn(t), 1);
  }
  MSIM_HD void conn_drop(int e) { conn_drop_half(e, 3); }
  /// halves: 1 = the Sender, 2 = the Receiver of end `e` (direction e is the one this end sends on)
  MSIM_HD void conn_drop_half(int e, int halves) {
    int c = e >> 1, d_tx = e, d_rx = e ^ 1;
    // `(tx, rx)` bindings drop in reverse order: the Receiver first, then the Sender
    if (halves & 2) {
      ch_flags(d_rx) &= ~CH_RX_ALIVE;  // UnboundedReceiver dropped: the peer's send fails, its closed() completes
      if (C::USES_CLOSED) wake_waiter(w(L::W_CH_CLOSED_WAITER, d_rx));
    }
    if (halves & 1) {
      ch_flags(d_tx) &= ~CH_TX_ALIVE;  // last UnboundedSender dropped: the peer's recv wakes and drains
      wake_waiter(w(L::W_CH_WAITER, d_tx));
    }
    if (!((ch_flags(d_tx) | ch_flags(d_rx)) & (CH_TX_ALIVE | CH_RX_ALIVE))) b(L::B_CONN_USED, c) = 0;
  }
  /// Sender::is_closed / the ready state of Sender::closed() (endpoint.rs:247-253): the peer dropped its Receiver
  MSIM_HD bool conn_tx_closed(int t) const {
    int e = t_conn(t);
    return !(ch_flags(e) & CH_RX_ALIVE);  // direction 2*conn+side is the one this end sends on
  }
  /// Timer callback of NetSim::send -> EndpointSocket::deliver -> Mailbox::deliver
  /// (endpoint.rs:310-351).  The target socket OBJECT was captured at send time: if it has
  /// been closed/reset since, the message lands in a mailbox nobody can read (dropped here).
  MSIM_HD void deliver(int m) {
    u32 d = w(L::W_M_DST, m);
    int k = d >> 16;
    if (sk_state(k) != 1 || (sk_gen(k) & 0xffff) != (d & 0xffff)) {
      msg_free(m);
      return;
    }
    u64 tag = q(L::Q_M_TAG, m);
    u32 rl = b(L::B_SK_REGLEN, k);
MSIM_NOUNROLL
    for (u32 i = 0; i < rl; i++) {
      if (q(L::Q_RG_TAG, k * C::NREG + i) == tag) {
        u32 waiter = w(L::W_RG_WAITER, k * C::NREG + i);
        rl--;  // swap_remove(i)
        q(L::Q_RG_TAG, k * C::NREG + i) = q(L::Q_RG_TAG, k * C::NREG + rl);
        w(L::W_RG_WAITER, k * C::NREG + i) = w(L::W_RG_WAITER, k * C::NREG + rl);
        b(L::B_SK_REGLEN, k) = (u8)rl;
        // oneshot::Sender::send: registrations of dropped receivers are removed eagerly
        // (mbox_unregister), so the receiver is alive: store the value, wake its task
        int t = waiter >> 20;
        t_rx(t) = (u8)m;
        wake(t, waiter & 0xfffffu);
        return;
      }
    }
    u32 ml = b(L::B_SK_MBLEN, k);
    if (MSIM_UNLIKELY(ml >= (u32)C::NMBOX)) {
      fail_cap(MSIM_CAP_MAILBOX);
      msg_free(m);
      return;
    }
    b(L::B_MB, k * C::NMBOX + ml) = (u8)m;
    b(L::B_SK_MBLEN, k) = (u8)(ml + 1);
  }
  /// Mailbox::recv(tag) (endpoint.rs:353-362) on the endpoint the task holds: either fills the
  /// task's oneshot at once (t_rx) or registers (tag, oneshot sender)
  MSIM_HD void mbox_recv(int t, int k, u64 tag) {
    u32 ml = b(L::B_SK_MBLEN, k);
MSIM_NOUNROLL
    for (u32 i = 0; i < ml; i++) {
      int m = b(L::B_MB, k * C::NMBOX + i);
      if (q(L::Q_M_TAG, m) == tag) {
        ml--;
        b(L::B_MB, k * C::NMBOX + i) = b(L::B_MB, k * C::NMBOX + ml);  // swap_remove
        b(L::B_SK_MBLEN, k) = (u8)ml;
        t_rx(t) = (u8)m;
        return;
      }
    }
    u32 rl = b(L::B_SK_REGLEN, k);
    if (MSIM_UNLIKELY(rl >= (u32)C::NREG)) {
      fail_cap(MSIM_CAP_REGISTERED);
      return;
    }
    q(L::Q_RG_TAG, k * C::NREG + rl) = tag;
    w(L::W_RG_WAITER, k * C::NREG + rl) = ((u32)t << 20) | t_gen(t);
    b(L::B_SK_REGLEN, k) = (u8)(rl + 1);
    t_rx(t) = NONE8;
  }
  /// Dropping a oneshot::Receiver that is still registered.  The reference leaves the dead
  /// sender in `registered` until a message with its tag arrives; removing it eagerly is
  /// unobservable as long as no two LIVE registrations of a mailbox share a tag (the shipped
  /// workloads guarantee it) — DESIGN.md "eager registration GC".
  MSIM_HD void mbox_unregister(int t) {
    int k = t_sock(t);
    if (k == NONE8 || !sk_state(k) || (u8)sk_gen(k) != t_sockgen(t)) return;
    u32 me = ((u32)t << 20) | t_gen(t);
    u32 rl = b(L::B_SK_REGLEN, k);
MSIM_NOUNROLL
    for (u32 i = 0; i < rl; i++)
      if (w(L::W_RG_WAITER, k * C::NREG + i) == me) {
        rl--;
        q(L::Q_RG_TAG, k * C::NREG + i) = q(L::Q_RG_TAG, k * C::NREG + rl);
        w(L::W_RG_WAITER, k * C::NREG + i) = w(L::W_RG_WAITER, k * C::NREG + rl);
        b(L::B_SK_REGLEN, k) = (u8)rl;
        return;
      }
  }
  /// take the message out of the task's oneshot into the result registers
  MSIM_HD void unpack_rx(int t) {
    int m = t_rx(t);
    rv0 = w(L::W_M_V0, m);
    rv1 = w(L::W_M_V1, m);
    rfrom = w(L::W_M_FROM, m);
    rw = q(L::Q_M_W, m);
    msg_free(m);
    t_rx(t) = NONE8;
  }
  MSIM_HD int msg_prepare(u64 tag, u64 wv, u32 v0, u32 v1, u32 dst) {
    int m = msg_alloc();
    if (m < 0) return -1;
    q(L::Q_M_TAG, m) = tag;
    q(L::Q_M_W, m) = wv;
    w(L::W_M_V0, m) = v0;
    w(L::W_M_V1, m) = v1;
    w(L::W_M_FROM, m) = dst;  // the destination rides in the FROM field until the send
    return m;
  }

  // ---------------- lifecycle ----------------
  /// Runtime::with_seed_and_config (runtime/mod.rs:53-69) for this seed + block_on's spawn of main
  MSIM_HD void init_inl(u64 seed) {
    u64* z = reinterpret_cast<u64*>(sb);
MSIM_NOUNROLL
    for (int i = 0; i < L::BYTES / 8; i++) z[i] = 0;
    // seed_from_u64: SplitMix64 x4
    u64 x = seed;
    u64 st[4];
    for (int i = 0; i < 4; i++) {
      x += 0x9e3779b97f4a7c15ull;
      u64 zz = x;
      zz = (zz ^ (zz >> 30)) * 0xbf58476d1ce4e5b9ull;
      zz = (zz ^ (zz >> 27)) * 0x94d049bb133111ebull;
      st[i] = zz ^ (zz >> 31);
    }
    r0 = st[0], r1 = st[1], r2 = st[2], r3 = st[3];
    now = 0;
    hash = MSIM_HASH_INIT;
    draws = 0;
    ev_pop = ev_fire = heap_len = rq_len = peak_timers = 0;
    status = MSIM_ST_RUNNING;
MSIM_NOUNROLL
    for (int t = 0; t < C::NT; t++) t_spawn(t) = NONE32;
    for (int wd = 0; wd < (C::NT + 63) / 64; wd++) {
      int n = C::NT - wd * 64;
      q(L::Q_TFREE, wd) = n >= 64 ? ~0ull : ((1ull << n) - 1);
    }
    for (int wd = 0; wd < L::MW; wd++) {
      int n = C::NMSG - wd * 64;
      q(L::Q_MFREE, wd) = n >= 64 ? ~0ull : ((1ull << n) - 1);
    }
    // TimeRuntime::new (time/mod.rs:26-38): draw #1 of every runtime
    w(L::W_BASE_SECS) = (u32)gen_range_u64(0, 60ull * 60 * 24 * 365);
    // node 0 = "madsim-main" (task/mod.rs:187-198)
    n_flags(0) = NF_EXISTS;
    w(L::W_NEXT_NODE) = 1;
    w(L::W_RESUME) = NONE32;
    for (int k = 0; k < L::NSA; k++) w(L::W_ACC_WAITER, k) = NONE32;
    int m = spawn_inl(WL::MAIN_KIND, 0, TF_JOINABLE);
    w(L::W_MAIN_SLOT) = (u32)m;
    // check_determinism enables the rng log AFTER the Runtime is constructed (runtime/mod.rs:185-187): the base-time
    // draw above is not part of it
    if (TRACE && P->dlog) reinterpret_cast<u32*>(P->dlog)[0] = 0;
  }
  static MSIM_NI void init_cold(unsigned char* sb_, const Params* P_, u64 now_, u64 seed) {
    Engine e_;
    e_.sb = sb_;
    e_.P = P_;
    e_.now = now_;
    e_.init_inl(seed);
    e_.store();
  }
  MSIM_HD void init(u64 seed) {
    init_cold(sb, P, 0, seed);
    load();
  }
  /// registers <- state block
  MSIM_HD void load() {
    load_g<G_ALL>();
    now = q(L::Q_NOW);
  }
  template <int G>
  MSIM_HD void load_g() {
#if !defined(__CUDACC__)
    dbg_valid |= G;
#endif
    if (G & G_RNG) {
      r0 = q(L::Q_RNG, 0), r1 = q(L::Q_RNG, 1), r2 = q(L::Q_RNG, 2), r3 = q(L::Q_RNG, 3);
      draws = q(L::Q_DRAWS);
    }
    if (G & G_CLOCK) {
      hash = q(L::Q_HASH);
      ev_pop = w(L::W_EV_POP), ev_fire = w(L::W_EV_FIRE);
    }
    if (G & G_QUEUES) {
      heap_len = w(L::W_HEAP_LEN), rq_len = w(L::W_RQ_LEN);
      status = w(L::W_STATUS);
      peak_timers = w(L::W_PEAK_TIMERS);
    }
  }
  /// registers -> state block
  MSIM_HD void store() {
    store_g<G_ALL>();
    q(L::Q_NOW) = now;
  }
  template <int G>
  MSIM_HD void store_g() {
#if !defined(__CUDACC__)
    if ((dbg_valid & G) != G) __builtin_trap();  // a light callee reached a heavier cold call
#endif
    if (G & G_RNG) {
      q(L::Q_RNG, 0) = r0, q(L::Q_RNG, 1) = r1, q(L::Q_RNG, 2) = r2, q(L::Q_RNG, 3) = r3;
      q(L::Q_DRAWS) = draws;
    }
    if (G & G_CLOCK) {
      q(L::Q_HASH) = hash;
      w(L::W_EV_POP) = ev_pop, w(L::W_EV_FIRE) = ev_fire;
    }
    if (G & G_QUEUES) {
      w(L::W_HEAP_LEN) = heap_len, w(L::W_RQ_LEN) = rq_len;
      w(L::W_STATUS) = status;
      w(L::W_PEAK_TIMERS) = peak_timers;
    }
  }
  MSIM_HD void result(SeedResult& r) {
    r.trace_hash = hash;
    r.elapsed_ns = now;
    r.draws = draws;
    r.ev_pop = ev_pop;
    r.ev_fire = ev_fire;
    r.status = status;
    r.fail_code = w(L::W_FAIL_CODE);
    r.user[0] = w(L::W_USER0), r.user[1] = w(L::W_USER1), r.user[2] = w(L::W_USER2), r.user[3] = w(L::W_USER3);
    r.peak_timers = peak_timers;
    r.peak_tasks = w(L::W_PEAK_TASKS);
  }

  /// The poll of one task: resume whatever primitive its future is suspended in, then run the
  /// state machine, executing the primitives it requests, until it is Pending / Ready / panics.
  /// Each primitive reproduces the poll-by-poll behaviour of SURVEY.md §3.4 (timers pushed per
  /// pending poll, draws in program order).
  /// poll_iter is ONE round of that loop — resume the awaited primitive, or run the state machine up to its next
  /// request and register it — and returns POLL_AGAIN when the poll goes on.  The single-lane kernels loop over it
  /// (poll_task); the multi-lane kernel runs the rounds of all lanes of a warp in lock step.
  MSIM_HD int poll_iter(int t) {
    {
      if (MSIM_UNLIKELY(!running())) return POLL_PENDING;
      u8 wt = t_wait(t);
      if (wt != W_NONE) {
        u8 st = wt & 0x0f;
        bool pend = false;
        const bool sel_closed = C::USES_CLOSED && (wt & WF_CLOSED);
        if (sel_closed) {  // select_biased!: tx.closed() is polled before the other branch
          int d = t_conn(t);
          if (!(ch_flags(d) & CH_RX_ALIVE)) {
            ret = -2;
            wt = W_NONE;  // the select is over; the other future is dropped where it stands
            st = 15;      // (no case below)
          } else {
            w(L::W_CH_CLOSED_WAITER, d) = ((u32)t << 20) | t_gen(t);
          }
        }
        if (wt & WF_RD) {  // NetSim::rand_delay: draw, then `sleep(delay).await`
          t_dl0(t) = rand_delay_deadline();
          wt = (wt & ~WF_RD) | WF_SLEEP;
        }
        if (wt & WF_SLEEP) {  // Sleep::poll (time/sleep.rs:47-54): EVERY pending poll registers another timer
          u64 dl = t_dl0(t);
          if (now < dl) {
            timer_push(dl, timer_payload(TK_WAKE, t, t_gen(t)));
            pend = true;
          } else {
            wt &= ~WF_SLEEP;
          }
        }
        if (!pend) {
          int k = t_sock(t);
          switch (st) {
            case W_SLEEP:
              wt = W_NONE;
              break;
            case W_BIND:  // BindGuard::bind after rand_delay (net/mod.rs:446-480)
              ret = sock_bind(t_node(t), (u32)t_dl1(t));
              if (ret >= 0) sock_adopt(t, ret);
              wt = W_NONE;
              break;
            case W_SEND:
            case W_CALL_SEND: {  // NetSim::send after rand_delay (shared by send_to_raw and call's send)
              int m = t_msg(t);
              t_msg(t) = NONE8;
              u64 rsp_tag = q(L::Q_M_W, m);
              net_send(b(L::B_SK_NODE, k), sk_port(k), w(L::W_M_FROM, m), m);
              if (st == W_SEND) {
                wt = W_NONE;
                break;
              }
              mbox_recv(t, k, rsp_tag);  // call: recv_from_raw(rsp_tag): mailbox.recv
              wt = st = W_CALL_RX;
            }
            // fallthrough
            case W_CALL_RX:
            case W_RECV_RX:  // oneshot::Receiver await; then rand_delay AFTER the receipt
              if (t_rx(t) == NONE8) {
                pend = true;
              } else {
                t_wait(t) = (st == W_RECV_RX ? W_RECV_DELAY : W_CALL_DELAY) | WF_RD;
                return POLL_AGAIN;
              }
              break;
            case W_RECV_DELAY:
              unpack_rx(t);
              wt = W_NONE;
              break;
            case W_CALL_DELAY:
              unpack_rx(t);
              wt = W_NONE;
              ret = CALL_OK;
              if (rfrom != t_c(t)) return panic(31);  // assert_eq!(from, dst) (rpc.rs:125)
              break;
            case W_CONNECT:
              if (C::NCONN) ret = connect1(t, (u32)t_dl1(t));
              wt = W_NONE;
              break;
            case W_ACCEPT: {  // conn_rx.recv().await (async_channel): FIFO
              if (!C::NCONN) break;  // workloads without connections carry none of this code
              u32 al = b(L::B_ACC_LEN, k);
              if (al == 0) {
                w(L::W_ACC_WAITER, k) = ((u32)t << 20) | t_gen(t);
                pend = true;
              } else {
                int c = b(L::B_ACC, k * C::NACC);
                for (u32 i = 1; i < al; i++) b(L::B_ACC, k * C::NACC + i - 1) = b(L::B_ACC, k * C::NACC + i);
                b(L::B_ACC_LEN, k) = (u8)(al - 1);
                t_conn(t) = (u8)(2 * c + 1);
                if (C::USES_SPLIT) t_connhalf(t) = 3;
                rfrom = w(L::W_CONN_SRC, c);
                ret = 0;
                wt = W_NONE;
              }
              break;
            }
            case W_CH_RECV: {  // poll_next of the channel's stream (net/mod.rs:385-402)
              if (!C::NCONN) break;
              int e = t_conn(t);
              int d = 2 * (e >> 1) + ((e & 1) ^ 1);
              for (;;) {
                u8 ph = b(L::B_CH_PHASE, d);
                if (ph == 0) {  // rx.recv().await
                  u32 len = b(L::B_CH_LEN, d);
                  if (len == 0) {
                    if (!(ch_flags(d) & CH_TX_ALIVE)) {  // None: every sender is gone
                      ret = -1;
                      wt = W_NONE;
                    } else {
                      w(L::W_CH_WAITER, d) = ((u32)t << 20) | t_gen(t);
                      pend = true;
                    }
                    break;
                  }
                  u32 h = b(L::B_CH_HEAD, d);
                  q(L::Q_CH_CUR, d) = q(L::Q_CH_ARRIVE, d * C::NCQ + h);
                  w(L::W_CH_CUR_V0, d) = w(L::W_CH_V0, d * C::NCQ + h);
                  w(L::W_CH_CUR_V1, d) = w(L::W_CH_V1, d * C::NCQ + h);
                  b(L::B_CH_HEAD, d) = (u8)((h + 1) % C::NCQ);
                  b(L::B_CH_LEN, d) = (u8)(len - 1);
                  q(L::Q_CH_BACKOFF, d) = NS_PER_MS;
                } else {  // the stream's Sleep
                  u64 dl = q(L::Q_CH_DL, d);
                  if (now < dl) {
                    timer_push(dl, timer_payload(TK_WAKE, t, t_gen(t)));
                    pend = true;
                    break;
                  }
                  if (ph == 2) {  // sleep_until(arrive) done: yield value
                    b(L::B_CH_PHASE, d) = 0;
                    rv0 = w(L::W_CH_CUR_V0, d);
                    rv1 = w(L::W_CH_CUR_V1, d);
                    ret = 0;
                    wt = W_NONE;
                    break;
                  }
                  u64 bo = q(L::Q_CH_BACKOFF, d) * 2;  // backoff = (backoff * 2).min(10 s); retry
                  q(L::Q_CH_BACKOFF, d) = bo > 10 * NS_PER_SEC ? 10 * NS_PER_SEC : bo;
                  q(L::Q_CH_CUR, d) = ch_test_link(d);
                }
                // wait until the link is ready: Some(arrive) -> sleep_until(arrive); None -> sleep(backoff)
                u64 st8 = q(L::Q_CH_CUR, d);
                if (st8 != NONE64) {
                  q(L::Q_CH_DL, d) = sleep_deadline(st8);
                  b(L::B_CH_PHASE, d) = 2;
                } else {
                  q(L::Q_CH_DL, d) = sleep_deadline(now + q(L::Q_CH_BACKOFF, d));
                  b(L::B_CH_PHASE, d) = 1;
                }
              }
              break;
            }
            case W_CTRLC: {  // rx.changed(): ready once the version moved past the one seen at subscribe()
              if (!C::USES_CTRLC) break;
              int n = t_node(t);
              if ((w(L::W_N_CTRLC, n) & 0x7fffffffu) != (u32)t_dl1(t)) {
                ret = 0;
                wt = W_NONE;
              } else {
                u32 me = ((u32)t << 20) | t_gen(t);
                u32 cur = w(L::W_N_CTRLC_WAITER, n);
                if (cur != NONE32 && cur != me && t_gen(cur >> 20) == (cur & 0xfffffu) && t_wait(cur >> 20) == W_CTRLC)
                  fail_cap(MSIM_CAP_REGISTERED);  // one parked receiver per node in this build
                w(L::W_N_CTRLC_WAITER, n) = me;
                pend = true;
              }
              break;
            }
            case W_JOIN: {
              if (!C::USES_JOIN) break;
              int target = (int)t_c(t);
              int r = join_poll(t, target);
              if (r == 0) {
                pend = true;
              } else {
                ret = r;
                rv0 = t_a(target);
                maybe_free(target);
                wt = W_NONE;
              }
              break;
            }
          }
        }
        if (sel_closed && wt == W_NONE) {  // the select completed: its closed() future (and waker) is gone
          w(L::W_CH_CLOSED_WAITER, t_conn(t)) = NONE32;
          if (st == W_SLEEP) ret = 0;
        }
        t_wait(t) = wt;
        if (pend) {
          if (C::USES_TIMEOUT && t_dl2(t) != NONE64 && (st < W_CALL_SEND || st > W_CALL_DELAY)) {
            u64 dl2 = t_dl2(t);
            if (now < dl2) {
              timer_push(dl2, timer_payload(TK_WAKE, t, t_gen(t)));
              return POLL_PENDING;
            }
            // Elapsed: the future inside the timeout is dropped where it stands
            if (st == W_RECV_RX) mbox_unregister(t);
            if (t_rx(t) != NONE8) {
              msg_free(t_rx(t));
              t_rx(t) = NONE8;
            }
            if (t_msg(t) != NONE8) {
              msg_free(t_msg(t));
              t_msg(t) = NONE8;
            }
            if (sel_closed) w(L::W_CH_CLOSED_WAITER, t_conn(t)) = NONE32;
            if (C::USES_CTRLC && st == W_CTRLC) w(L::W_N_CTRLC_WAITER, t_node(t)) = NONE32;  // Notified::drop leaves the list
            t_wait(t) = W_NONE;
            t_dl2(t) = NONE64;
            ret = -3;
          } else
          if (st < W_CALL_SEND || st > W_CALL_DELAY) return POLL_PENDING;
          else {
          // timeout(): select_biased polls the call first, then the timeout's Sleep (time/mod.rs:128-140)
          u64 dl1 = t_dl1(t);
          if (now < dl1) {
            timer_push(dl1, timer_payload(TK_WAKE, t, t_gen(t)));
            return POLL_PENDING;
          }
          // Elapsed: the call future is dropped where it stands
          if (st == W_CALL_RX) mbox_unregister(t);
          if (t_rx(t) != NONE8) {
            msg_free(t_rx(t));
            t_rx(t) = NONE8;
          }
          if (t_msg(t) != NONE8) {
            msg_free(t_msg(t));
            t_msg(t) = NONE8;
          }
          t_wait(t) = W_NONE;
          ret = CALL_TIMEOUT;
          }
        }
      }
      // ---- run the state machine until its next request ----
      int rq = WL::fsm(*this, t);
      switch (rq) {
        case RQ_PENDING:
          return POLL_PENDING;
        case RQ_READY:
          return POLL_READY;
        case RQ_PANIC:
          return panic(ya);
        case RQ_SLEEP_FOR:
          xa += now;
          // fallthrough
        case RQ_SLEEP_UNTIL:
          t_dl0(t) = sleep_deadline(xa);
          t_wait(t) = W_SLEEP | WF_SLEEP;
          break;
        case RQ_SLEEP_OR_CLOSED:
          if (!C::USES_CLOSED) break;
          t_dl0(t) = sleep_deadline(now + xa);
          t_wait(t) = W_SLEEP | WF_SLEEP | WF_CLOSED;
          break;
        case RQ_CH_RECV_OR_CLOSED:
          if (!C::USES_CLOSED) break;
          t_wait(t) = W_CH_RECV | WF_CLOSED;
          break;
        case RQ_TIMEOUT_BEGIN:
          if (!C::USES_TIMEOUT) break;
          t_dl2(t) = sleep_deadline(now + xa);
          break;
        case RQ_TIMEOUT_END:
          if (!C::USES_TIMEOUT) break;
          t_dl2(t) = NONE64;
          break;
        case RQ_SLEEP_RAW:
          t_dl0(t) = xa;
          t_wait(t) = W_SLEEP | WF_SLEEP;
          break;
        case RQ_CTRL_C: {  // NodeInfo::ctrl_c(): get_or_insert the watch channel, subscribe at the current version
          if (!C::USES_CTRLC) break;
          int n = t_node(t);
          u32 c = w(L::W_N_CTRLC, n) | 0x80000000u;
          w(L::W_N_CTRLC, n) = c;
          t_dl1(t) = c & 0x7fffffffu;
          t_wait(t) = W_CTRLC;
          break;
        }
        case RQ_RAND_DELAY:
          t_wait(t) = W_SLEEP | WF_RD;
          break;
        case RQ_BIND:
          t_dl1(t) = ya;
          t_wait(t) = W_BIND | WF_RD;
          break;
        case RQ_CALL:  // call_timeout: the timeout's Sleep is created first; call: rsp_tag = random::<u64>()
          t_dl1(t) = sleep_deadline(now + xb);
          t_c(t) = ya;
          xb = random_u64();
          yd = W_CALL_SEND | WF_RD;
          goto prepare;
        case RQ_SEND:
          yd = W_SEND | WF_RD;
        prepare : {
          int m = msg_prepare(xa, xb, yb, yc, ya);
          if (m < 0) return POLL_PENDING;
          t_msg(t) = (u8)m;
          t_wait(t) = (u8)yd;
          break;
        }
        case RQ_RECV:
          mbox_recv(t, t_sock(t), xa);
          t_wait(t) = W_RECV_RX;
          break;
        case RQ_CREATE_NODE:
          ret = create_node(ya, yb);
          if (ret < 0) return POLL_PENDING;
          break;
        case RQ_SPAWN: {
          int c = spawn(ya & 0xff, yb, (ya & SPAWN_JOINABLE) ? TF_JOINABLE : 0);
          ret = c;
          if (c < 0) return POLL_PENDING;
          t_a(c) = yc;
          t_b(c) = yd;
          t_c(c) = (u32)xa;
          t_dl1(c) = xb;
          if (ya & SPAWN_EP) sock_ref(c, t_sock(t));
          if (C::NCONN && (ya & SPAWN_CONN)) {
            t_conn(c) = t_conn(t);
            t_conn(t) = NONE8;
          }
          if (C::USES_SPLIT && (ya & SPAWN_CONN_TX)) {  // the child takes the Sender, this future keeps the Receiver
            t_conn(c) = t_conn(t);
            t_connhalf(c) = 1;
            t_connhalf(t) = 2;
          }
          break;
        }
        case RQ_GEN_RANGE:
          ret64 = gen_range_u64(xa, xb);
          break;
        case RQ_JOIN:
          t_c(t) = ya;
          t_wait(t) = W_JOIN;
          break;
        case RQ_CONNECT1:
          t_dl1(t) = ya;
          t_wait(t) = W_CONNECT | WF_RD;
          break;
        case RQ_ACCEPT1:
          t_wait(t) = W_ACCEPT | WF_RD;
          break;
        case RQ_CH_RECV:
          t_wait(t) = W_CH_RECV;
          break;
        case RQ_CH_SEND: {  // PayloadSender::send: state = test_link(); tx.send((value, state)).ok()
          if (!C::NCONN) break;
          int e = t_conn(t);
          int d = 2 * (e >> 1) + (e & 1);
          u64 st8 = ch_test_link(d);
          if (!(ch_flags(d) & CH_RX_ALIVE)) {
            ret = -1;  // ConnectionReset
            break;
          }
          u32 len = b(L::B_CH_LEN, d);
          if (len >= (u32)C::NCQ) {
            fail_cap(MSIM_CAP_MSGS);
            return POLL_PENDING;
          }
          u32 pos = (b(L::B_CH_HEAD, d) + len) % C::NCQ;
          q(L::Q_CH_ARRIVE, d * C::NCQ + pos) = st8;
          w(L::W_CH_V0, d * C::NCQ + pos) = yb;
          w(L::W_CH_V1, d * C::NCQ + pos) = yc;
          b(L::B_CH_LEN, d) = (u8)(len + 1);
          wake_waiter(w(L::W_CH_WAITER, d));
          ret = 0;
          break;
        }
        case RQ_ADVANCE:  // time::advance (time/mod.rs:103-106,195-198): clock += d, then the due timers fire
          now += xa;      // inside this poll: the event loop fires them at its single expire site and
          return POLL_ADVANCE;  // re-enters this poll (W_RESUME) before anything else runs
      }
    }
    return POLL_AGAIN;
  }
  MSIM_HD int poll_task(int t) {
    for (;;) {
      int r = poll_iter(t);
      if (r != POLL_AGAIN) return r;
    }
  }

  /// Runnable::run (async-task) + the panic policy of run_all_ready (task/mod.rs:285-317), in three parts so that the
  /// multi-lane kernel can align the polls of a warp's lanes: run_begin, the poll, run_end.
  /// run_begin: false = the task had been closed while it was queued: its future was dropped instead of polled
  MSIM_HD bool run_begin(int t, bool cont) {
    if (!cont) {
      u8 f = t_flags(t);
      if (C::USES_KILL && MSIM_UNLIKELY(f & TF_CLOSED)) {
        drop_future(t);
        t_flags(t) = t_flags(t) & ~TF_SCHEDULED;
        notify_awaiter(t);
        maybe_free(t);
        return false;
      }
      t_flags(t) = (f & ~TF_SCHEDULED) | TF_RUNNING;
    }
    return true;
  }
  /// the future returned Ready: `future.await; h.exit()` for a node's init task, drop the future, complete the task
  MSIM_HD void finish_ready_inl(int t) {
    if (C::USES_KILL && MSIM_UNLIKELY(t_flags(t) & TF_INIT)) node_exit(t_node(t), t_epoch(t));
    drop_future(t);
    t_flags(t) = (t_flags(t) & ~(TF_RUNNING | TF_SCHEDULED)) | TF_COMPLETED;
    notify_awaiter(t);
    maybe_free(t);
  }
  /// run_end: true = the poll was interrupted by time::advance (still RUNNING: it continues after the timers fired)
  MSIM_HD bool run_end(int t, int r) {
    if (MSIM_UNLIKELY(r == POLL_ADVANCE)) return true;
    if (r == POLL_READY) {
      if (X_FINISH)
        cold_g<(C::USES_KILL ? G_ALL : G_QUEUES)>([t](Engine& e_) { e_.finish_ready_inl(t); });  // (node_exit is a full cold call)
      else
        finish_ready_inl(t);
    } else if (MSIM_UNLIKELY(r == POLL_PANIC)) {
      // async-task's guard: close, drop the future, notify the awaiter
      t_flags(t) = (t_flags(t) & ~(TF_RUNNING | TF_SCHEDULED)) | TF_CLOSED;
      drop_future(t);
      notify_awaiter(t);
      int n = t_node(t);
      maybe_free(t);
      panicked(n);
    } else {
      u8 f = t_flags(t) & ~TF_RUNNING;
      t_flags(t) = f;
      if (C::USES_KILL && MSIM_UNLIKELY(f & TF_CLOSED)) {
        t_flags(t) = f & ~TF_SCHEDULED;
        drop_future(t);
        notify_awaiter(t);
        maybe_free(t);
      } else if (f & TF_SCHEDULED) {
        rq_push(t);  // woken while running
      }
    }
    return false;
  }
  MSIM_HD bool run_task(int t, bool cont) {
    if (!run_begin(t, cont)) return false;
    return run_end(t, poll_task(t));
  }
  /// run_all_ready's panic policy (task/mod.rs:289-317): restart_on_panic nodes are killed and
  /// restarted after gen_range(1s..10s); anything else fails the test
  /// `restart_on_panic_matching.iter().any(|s| error_msg.contains(s))` (task/mod.rs:298-303): device panics carry a code,
  /// not a message, so a pattern is the code of the panic it matches (bit c of the node's mask = Caps::PMATCH_BASE + c)
  MSIM_HD bool panic_matches(int n) const {
    if (!C::USES_PMATCH) return false;
    u32 c = w(L::W_PANIC_CODE) - (u32)C::PMATCH_BASE;
    return c < 32 && ((w(L::W_N_PMATCH, n) >> c) & 1);
  }
  MSIM_HD void panicked_inl(int n) {
    if ((n_flags(n) & NF_RESTART_ON_PANIC) || panic_matches(n)) {
      u64 delay = sample_inclusive_u64(NS_PER_SEC, 10 * NS_PER_SEC - 1);  // UniformDuration Medium
      kill(n);
      timer_push(now + delay, timer_payload(TK_RESTART, n, 0));
    } else {
      fail(MSIM_ST_FAIL_ASSERT, w(L::W_PANIC_CODE));
    }
  }
  MSIM_HD void panicked(int n) {
    cold([n](Engine& e_) { e_.panicked_inl(n); });
  }
  MSIM_HD int panic(u32 code) {
    w(L::W_PANIC_CODE) = code;
    return POLL_PANIC;
  }

  // ---------------- the event step, in the pieces both kernels are made of ----------------
  enum : int { ACT_NONE = 0, ACT_RUN = 1, ACT_DROP = 2 };
  /// try_recv_random (utils/mpsc.rs:73-83) + the dispatch decision of run_all_ready (task/mod.rs:262-283).
  /// ACT_NONE: the pop was a SKIP or a PARK — that step is over (no clock advance, no timers).
  MSIM_HD int pop_ready(int& t) {
    u32 len = rq_len;
    u32 idx = gen_range_small(0, len);
    t = b(L::B_RQ, idx);
    b(L::B_RQ, idx) = b(L::B_RQ, len - 1);  // swap_remove
    rq_len = len - 1;
    ev_pop++;
    u64 aux = ((u64)len << 32) | idx;
    u8 f = 0;
    u32 kind = MSIM_EV_POLL;
    if (C::USES_KILL) {  // programs that never kill / pause / abort / exit a node only ever POLL
      f = t_flags(t);
      if (!(f & TF_ALIVE))
        kind = MSIM_EV_SKIP;  // metadata().upgrade() failed: future already dropped
      else if ((f & TF_CANCELLED) || task_node_killed(t))
        kind = MSIM_EV_DROP;
      else if (n_flags(t_node(t)) & NF_PAUSED)
        kind = MSIM_EV_PARK;
    }
    trace(now, aux, kind, t_spawn(t));
    if (MSIM_UNLIKELY(kind == MSIM_EV_SKIP)) {
      t_flags(t) = f & ~TF_SCHEDULED;
      maybe_free(t);
      return ACT_NONE;
    }
    if (MSIM_UNLIKELY(kind == MSIM_EV_PARK)) {
      u32 pl = w(L::W_PK_LEN);
      if (pl >= (u32)C::NPARK) {
        fail_cap(MSIM_CAP_PARKED);
        return ACT_NONE;
      }
      b(L::B_PK, pl) = (u8)t;
      w(L::W_PK_LEN) = pl + 1;
      return ACT_NONE;
    }
    return kind == MSIM_EV_DROP ? ACT_DROP : ACT_RUN;
  }
  /// The ready Vec is drained: block_on returns if the main task is finished (task/mod.rs:241-243), panics on a
  /// deadlock (:250), else advance_to_next_event (time/mod.rs:45-60): callbacks fire under the OLD clock.
  /// false: the seed has stopped; true: `limit` is what Timer::expire gets.
  MSIM_HD bool jump_limit(u64& limit) {
    int m = (int)w(L::W_MAIN_SLOT);
    u8 mf = t_flags(m);
    if (MSIM_UNLIKELY(mf & (TF_COMPLETED | TF_CLOSED))) {
      if (mf & TF_COMPLETED)
        status = MSIM_ST_PASS;
      else
        fail(MSIM_ST_FAIL_ASSERT, 0xdead);  // main task cancelled
      return false;
    }
    if (MSIM_UNLIKELY(heap_len == 0)) {
      fail(MSIM_ST_FAIL_DEADLOCK, 0);
      return false;
    }
    limit = h_dl(0) + 50;
    return true;
  }
  /// after Timer::expire(limit): the clock moves to `limit`; Runtime::set_time_limit (task/mod.rs:253-258)
  MSIM_HD void clock_to(u64 limit) {
    if (now != limit) {
      now = limit;
      if (MSIM_UNLIKELY(now >= P->time_limit_cmp))  // u64::MAX when there is no limit
        fail(MSIM_ST_FAIL_TIME_LIMIT, 0);
    }
  }
  /// after the poll / drop of a popped runnable: remember an interrupted poll, or advance time 50-100 ns
  /// (task/mod.rs:319-321); the timers due by then fire next
  MSIM_HD u64 after_run(int t, bool adv) {
    if (MSIM_UNLIKELY(adv))
      w(L::W_RESUME) = (u32)t;
    else
      now += gen_range_small(50, 100);
    return now;
  }

  /// ONE EVENT STEP of Executor::block_on (task/mod.rs:239-259): if the ready Vec is not empty,
  /// pop a random runnable, dispatch it and advance the clock 50-100 ns; otherwise finish, or
  /// jump to the next timer.  Returns false when the seed has stopped.
  MSIM_HD bool step() {
    u64 limit;
    u32 rs = w(L::W_RESUME);
    bool cont = rs != NONE32;  // a poll interrupted by time::advance continues before anything else
    if (cont || rq_len > 0) {
      int t;
      int act = ACT_RUN;
      if (cont) {
        t = (int)rs;
        w(L::W_RESUME) = NONE32;
      } else {
        act = pop_ready(t);
        if (MSIM_UNLIKELY(act == ACT_NONE)) return running();
      }
      bool adv = false;
      if (MSIM_UNLIKELY(act == ACT_DROP))
        drop_runnable(t);
      else
        adv = run_task(t, cont);
      if (MSIM_UNLIKELY(!running())) return false;
      limit = after_run(t, adv);
    } else {
      if (!jump_limit(limit)) return false;
    }
    expire(limit);  // the one place timers fire from the event loop
    clock_to(limit);
    return running();
  }
};

}  // namespace msim
