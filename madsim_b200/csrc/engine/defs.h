// msim_b200 engine — common definitions.
//
// The engine is the per-seed event-loop step of the reference
// (madsim/src/sim/task/mod.rs:220-323, time/mod.rs:45-60), re-designed for one GPU lane per
// seed over a columnar, bounded, shared-memory-resident state block.  The engine headers are
// `__host__ __device__` so that tests/enginecheck can compile the SAME state machines with g++
// and diff them against the oracle without a GPU; the shipped library (libmsim_b200.so) only
// ever instantiates them inside CUDA kernels.
#pragma once
#include <stdint.h>

#include "../../../include/msim_b200.h"

#if defined(__CUDACC__)
#define MSIM_HD __host__ __device__ __forceinline__
#define MSIM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MSIM_HD inline
#define MSIM_HD_NOINLINE inline
#endif
// MSIM_NI: rarely executed runtime paths get ONE out-of-line copy on the device, so that the hot
// event loop stays dense in the instruction cache
#if defined(__CUDACC__)
#define MSIM_NI __host__ __device__ __noinline__
#else
#define MSIM_NI inline
#endif
#define MSIM_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define MSIM_LIKELY(x) __builtin_expect(!!(x), 1)
#if defined(__CUDACC__)
#define MSIM_NOUNROLL _Pragma("unroll 1")
#else
#define MSIM_NOUNROLL
#endif
// Heavy helpers that are called from many state-machine sites: one shared copy keeps the
// hot code inside the instruction cache (round-1 ncu: 57 % icc hit rate when all inlined).
#if defined(MSIM_INLINE_ALL)
#define MSIM_FN MSIM_HD
#else
#define MSIM_FN MSIM_HD_NOINLINE
#endif

// Which once-per-request runtime pieces a program runs out of line (Caps::COLD_*): bit 0 spawn, 1 task end, 2 channel link
// test + connect1.  Programs set their own value; this is the default (and the knob of the code-size experiments).
#ifndef MSIM_HEAP16_DEFAULT  // timer heap layout of programs whose Caps do not say: 16-byte records (1) or split arrays (0)
#define MSIM_HEAP16_DEFAULT 1
#endif
#ifndef MSIM_HEAP_ASM_SEL
#define MSIM_HEAP_ASM_SEL 1
#endif
// kernel experiments: -DMSIM_X_<FLAG>=0/1 overrides a program's own code-placement choice (engine.h X_*); -1 = the program's
#ifndef MSIM_X_SPAWN
#define MSIM_X_SPAWN -1
#endif
#ifndef MSIM_X_FINISH
#define MSIM_X_FINISH -1
#endif
#ifndef MSIM_X_CH
#define MSIM_X_CH -1
#endif
#ifndef MSIM_X_NOW
#define MSIM_X_NOW -1
#endif
#ifndef MSIM_X_BIND
#define MSIM_X_BIND -1
#endif
#ifndef MSIM_X_H16
#define MSIM_X_H16 -1
#endif
#ifndef MSIM_X_PORTFIRST
#define MSIM_X_PORTFIRST -1
#endif
#ifndef MSIM_COLD_LIGHT  // 0: every out-of-line call moves the whole register-cached state (kernel experiments)
#define MSIM_COLD_LIGHT 1
#endif
#ifndef MSIM_COLD_DEFAULT
#define MSIM_COLD_DEFAULT 0
#endif

namespace msim {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- portable intrinsics ----
MSIM_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
MSIM_HD int clz64(u64 x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
MSIM_HD int clz32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return __builtin_clz(x);
#endif
}
MSIM_HD int ffs64(u64 x) {  // index of lowest set bit; x != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
MSIM_HD u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }

constexpr u64 NS_PER_MS = 1000000ull;
constexpr u64 NS_PER_SEC = 1000000000ull;

// ---- poll results ----
enum : int { POLL_PENDING = 0, POLL_READY = 1, POLL_PANIC = 2, POLL_ADVANCE = 3, POLL_AGAIN = 4 };  // AGAIN: the poll goes on (poll_iter)

// ---- task flags (async-task's state bits + bookkeeping) ----
enum : u8 {
  TF_SCHEDULED = 1,   // a Runnable exists (in the ready Vec or parked)
  TF_RUNNING = 2,
  TF_COMPLETED = 4,
  TF_CLOSED = 8,
  TF_ALIVE = 16,      // the future (and its TaskInfo) has not been dropped yet
  TF_CANCELLED = 32,  // TaskInfo.cancelled (AbortHandle::abort)
  TF_JOINABLE = 64,   // a JoinHandle still exists: keep the slot after completion
  TF_INIT = 128,      // the node's init task: `future.await; h.exit()` (runtime/mod.rs:363-370)
};

// ---- node flags ----
enum : u8 { NF_EXISTS = 1, NF_KILLED = 2, NF_PAUSED = 4, NF_HAS_IP = 8, NF_RESTART_ON_PANIC = 16 };

// ---- timer payload: kind[31:28] | a[27:20] | b[19:0] ----
enum : u32 { TK_WAKE = 0, TK_DELIVER = 1, TK_RESTART = 2 };
static_assert(MSIM_EV_DELIVER == MSIM_EV_WAKE + 1 && MSIM_EV_RESTART == MSIM_EV_WAKE + 2, "event kinds follow TK_*");
MSIM_HD u32 timer_payload(u32 kind, u32 a, u32 b) { return (kind << 28) | ((a & 0xffu) << 20) | (b & 0xfffffu); }

// ---- socket addresses as the device sees them: ipk[31:24] | node[23:16] | port[15:0] ----
// Every node i with an IP owns exactly one address (10.0.0.i in the shipped workloads), so an
// IP is identified by its owner node.
enum : u32 { IPK_UNSPEC = 0, IPK_NODE = 1, IPK_LOOPBACK = 2 };
MSIM_HD u32 mk_addr(u32 ipk, u32 node, u32 port) { return (ipk << 24) | ((node & 0xff) << 16) | (port & 0xffff); }
MSIM_HD u32 addr_ipk(u32 a) { return a >> 24; }
MSIM_HD u32 addr_node(u32 a) { return (a >> 16) & 0xff; }
MSIM_HD u32 addr_port(u32 a) { return a & 0xffff; }

constexpr u8 NONE8 = 0xff;
constexpr u32 NONE32 = 0xffffffffu;

/// Kernel parameters: one per batch, identical for every seed (a digest of msim_config +
/// msim_workload_desc prepared on the host, see host_params.h).
struct Params {
  u32 workload;
  u32 n_nodes;
  u64 p[8];
  // packet loss: Bernoulli (rand 0.8): p == 1.0 → always true, no draw; else draw < p_int
  u64 loss_p_int;
  u32 loss_always;
  // send latency: UniformDuration over [lo, hi): Small → u32 sampler over sub-second nanos,
  // Medium → u64 sampler over total ns
  u32 lat_mode;  // 0 Small, 1 Medium
  u64 lat_base;  // Small: low_s * 1e9
  u64 lat_lo;    // Small: low_n ; Medium: low ns
  u64 lat_range; // inclusive range size
  u64 lat_zone;  // acceptance zone
  u64 time_limit_ns;  // 0 = none
  u64 time_limit_cmp; // time_limit_ns, or u64::MAX when there is none (one compare per event)
  u32 buggify;
  u32 reserved;
  u64 buggify_p10_int;  // (0.1 * 2^64) as u64, computed on the host in f64 like rand does
  // determinism log of a single-seed replay (rand.rs:64-89): {u32 count, u32 capacity, u8 bytes[]} or 0.  Set per
  // replay launch by msim_batch_determinism_log, never by make_params; read only by the TRACE kernels.
  unsigned char* dlog;
};

/// Per-seed result exactly as the C-ABI returns it.
typedef msim_seed_result SeedResult;

}  // namespace msim
