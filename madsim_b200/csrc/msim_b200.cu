// libmsim_b200.so — CUDA kernels (sm_100a) + the C-ABI of include/msim_b200.h.
//
// Kernel shape (B200-first):
//   * one CTA owns S consecutive seeds; their whole state is S seed-major blocks of
//     Layout::BYTES bytes that live in SHARED MEMORY for the duration of the launch (up to
//     227 KB per CTA → grid sized to one CTA per SM);
//   * the block is staged HBM→smem / smem→HBM with 1-D bulk async copies (TMA, UBLKCP in SASS)
//     completed through an mbarrier — only when a run is split into bursts; a run-to-completion
//     launch initialises the block in place from the seed and writes back 64 B per seed;
//   * lane = seed, but only `lanes_per_warp` lanes of each warp are populated (default 1): seeds
//     follow different branches of the event loop, so co-resident lanes serialise (measured:
//     1.01 active threads per instruction) while separate warps overlap on the 4 schedulers; the
//     seed count per SM is bounded by shared memory, not by threads;
//   * integer-only, no tensor cores: xoshiro256++, widening multiplies, heap sifts.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "engine/host_params.h"
#include "workloads/all.h"

using namespace msim;

namespace {

// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
constexpr int SMEM_HDR = 128;            // mbarrier lives here; the state block follows (128 B aligned)
constexpr unsigned BULK_CHUNK = 32768;   // bytes per cp.async.bulk
#ifndef MSIM_MAX_THREADS
#define MSIM_MAX_THREADS 896
#endif
constexpr int MAX_THREADS = MSIM_MAX_THREADS;  // warps per CTA x 32; bounds registers per thread

struct KernelArgs {
  Params P;
  uint64_t seed_base;
  uint64_t n_seeds;
  uint64_t max_events;      // per seed per launch; 0 = unlimited
  unsigned char* blob;      // [n_blocks][S * BYTES] checkpoint in HBM (may be null if never needed)
  SeedResult* results;      // [n_seeds]
  msim_trace_rec* trace;    // optional (single-seed replay)
  uint32_t* trace_n;
  uint32_t trace_cap;
  uint32_t S;               // seeds per CTA
  uint32_t L;               // populated lanes per warp
  uint32_t resume;          // 0: init from seed, 1: load checkpoint
  uint32_t* unfinished;     // device counter: seeds still RUNNING after this launch
  uint32_t* next_seed;      // device counter of the seed queue (dynamic mode)
  uint32_t dynamic;         // 1: run-to-completion launch, every warp slot pulls the next seed when its own finishes
  uint32_t lockstep;        // 1: dynamic launch with L >= 2 through the lock-step (phase-aligned) kernel
  uint32_t* cap_fail;       // device counter: seeds of this launch that ended with MSIM_ST_FAIL_CAPACITY
  const uint32_t* seed_list;  // general kernel, dynamic: queue entry i is seed offset seed_list[i] (capacity escalation)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
/// TMA 1-D bulk copy global → shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
/// TMA 1-D bulk copy shared → global (bulk async-group)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/// GENERAL = false: the batch kernel — run to completion, dynamic seed queue, no event budget, no
/// checkpoint staging, no trace records.  GENERAL = true: bursts / resume / single-seed replay.
template <class WL, bool GENERAL>
__global__ void __launch_bounds__(MAX_THREADS, 1) msim_step_kernel(const __grid_constant__ KernelArgs a) {
  using E = Engine<WL, GENERAL>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* block = smem + SMEM_HDR;

  const uint32_t S = a.S;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t col = warp * a.L + lane;  // this thread's seed column
  const uint64_t first = (uint64_t)blockIdx.x * S;
  const bool active = lane < a.L && col < S && (a.dynamic || first + col < a.n_seeds);
  const uint32_t block_bytes = S * (uint32_t)E::L::BYTES;
  unsigned char* gblob = a.blob ? a.blob + (size_t)blockIdx.x * block_bytes : nullptr;

  // ---- stage in: TMA bulk copy of the checkpoint (resume) ----
  if (GENERAL && a.resume) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, block_bytes);
      for (uint32_t off = 0; off < block_bytes; off += BULK_CHUNK) {
        uint32_t nb = block_bytes - off < BULK_CHUNK ? block_bytes - off : BULK_CHUNK;
        bulk_g2s(block + off, gblob + off, nb, bar);
      }
    }
    mbar_wait(bar, 0);
  }

  E e;
  e.sb = block + (size_t)col * E::L::BYTES;  // seed-major: one base register per lane
  e.P = &a.P;
  e.tr = nullptr;
  e.tr_cap = 0;
  e.tr_n = 0;
  e.status = MSIM_ST_PASS;
  if (active) {
    if (GENERAL && a.trace) {
      e.tr = a.trace;
      e.tr_cap = a.trace_cap;
    }
    // Static launches (bursts, resume, replay): seed = first + col, once.  Dynamic launches
    // (run to completion): seeds finish at very different event counts (a chaos schedule may last 10
    // or 50 virtual seconds), so CTAs are persistent (one per SM) and a slot whose seed is done takes
    // the next one from the queue (SURVEY §8e "handle load imbalance inside a GPU").  Seeds are
    // independent: the assignment cannot change any result.  One event-loop site for both.
    uint64_t idx = first + col;
    for (;;) {
      if (!GENERAL || a.dynamic) {
        idx = atomicAdd(a.next_seed, 1u);
        if (idx >= a.n_seeds) break;
      }
      if (GENERAL && a.seed_list) idx = a.seed_list[idx];  // escalation: the queue holds offsets into the batch
      if (GENERAL && a.resume) {
        e.load();
        if (a.trace) e.tr_n = *a.trace_n;
      } else {
        e.init(a.seed_base + idx);
      }
      // ---- the event loop: up to max_events steps out of shared memory / registers ----
      if (e.status == MSIM_ST_RUNNING) {
        if (GENERAL) {
          uint64_t budget = a.max_events ? a.max_events : ~0ull;
          while (budget-- && e.step()) {
          }
        } else {
          while (e.step()) {
          }
        }
      }
      e.store();
      SeedResult r;
      e.result(r);
      a.results[idx] = r;
      if (MSIM_UNLIKELY(r.status == MSIM_ST_FAIL_CAPACITY)) atomicAdd(a.cap_fail, 1u);
      if (GENERAL && !a.dynamic) break;
    }
    if (GENERAL && a.trace) *a.trace_n = e.tr_n;
  }
  // ---- stage out: in burst mode (a checkpoint buffer exists) the block always goes back to HBM,
  // finished seeds included, so that a later resume reloads a consistent image ----
  if (GENERAL && active && !a.dynamic && e.status == MSIM_ST_RUNNING) atomicAdd(a.unfinished, 1u);
  if (GENERAL && gblob && !a.dynamic) {
    fence_proxy_async();  // generic-proxy smem writes → visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t off = 0; off < block_bytes; off += BULK_CHUNK) {
        uint32_t nb = block_bytes - off < BULK_CHUNK ? block_bytes - off : BULK_CHUNK;
        bulk_s2g(gblob + off, block + off, nb);
      }
      bulk_commit_wait();
    }
  }
}

/// The multi-lane batch kernel (run to completion, dynamic seed queue, L >= 2 populated lanes per warp).
/// A warp advances its L seeds in LOCK STEP through the phases of the event step, so that the lanes execute the
/// common spine — jump to the next timer, timer pop (heap sift), ready-queue pop draw, clock-advance draw — as one
/// instruction stream and only diverge inside the polls, whose rounds (poll_iter) are aligned too:
///   refill | J: drained queue -> finished / deadlock / jump limit | X: fire due timers, one per round, while any
///   lane has one | P: pop a runnable | R: poll rounds while any lane is still polling | A: clock += 50..100 ns |
///   X again | retire finished seeds.
/// A lane whose queue is drained does J+X and then P+R+A+X in the same iteration (jump steps and pop steps
/// alternate in most programs), i.e. up to two event steps per iteration; every lane still executes exactly the
/// sequence of steps Engine::step() would, so results do not depend on L.  Warp votes (__any_sync) pick the
/// phases that still have work; __syncwarp semantics come with them.
template <class WL>
__global__ void __launch_bounds__(MAX_THREADS, 1) msim_step_kernel_ml(const __grid_constant__ KernelArgs a) {
  using E = Engine<WL, false>;
  using LY = typename E::L;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* block = smem + SMEM_HDR;
  const uint32_t S = a.S, L = a.L;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp * L >= S) return;
  const uint32_t nl = S - warp * L < L ? S - warp * L : L;  // populated lanes of this warp
  if (lane >= nl) return;
  const unsigned mask = nl >= 32 ? 0xffffffffu : ((1u << nl) - 1u);
  const uint32_t col = warp * L + lane;

  E e;
  e.sb = block + (size_t)col * LY::BYTES;
  e.P = &a.P;
  e.tr = nullptr;
  e.tr_cap = 0;
  e.tr_n = 0;
  e.status = MSIM_ST_PASS;
  bool live = false, exhausted = false;
  uint32_t idx = 0;
  for (;;) {
    // ---- refill: a lane without a seed takes the next one from the queue ----
    if (!live && !exhausted) {
      idx = atomicAdd(a.next_seed, 1u);
      if (idx < a.n_seeds) {
        e.init(a.seed_base + idx);
        live = true;
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(mask, live)) break;  // every lane is out of seeds
    // ---- J: the ready Vec is drained ----
    u64 limit = 0;
    bool ex = false;
    bool cont = live && e.w(LY::W_RESUME) != NONE32;
    if (live && !cont && e.rq_len == 0) ex = e.jump_limit(limit);
    // ---- X: timers due by `limit` ----
    for (;;) {
      bool due = ex && e.timer_due(limit);
      if (!__any_sync(mask, due)) break;
      if (due) e.fire_one();
    }
    if (ex) e.clock_to(limit);
    // ---- P: pop a runnable (or continue the poll that time::advance interrupted) ----
    int t = 0, act = E::ACT_NONE;
    if (live && e.running() && (cont || e.rq_len > 0)) {
      if (cont) {
        t = (int)e.w(LY::W_RESUME);
        e.w(LY::W_RESUME) = NONE32;
        act = E::ACT_RUN;
      } else {
        act = e.pop_ready(t);
      }
    }
    // ---- R: dispatch; the polls of all lanes advance round by round ----
    bool polling = false;
    if (act == E::ACT_DROP) e.drop_runnable(t);
    if (act == E::ACT_RUN) polling = e.run_begin(t, cont);
    int r = POLL_AGAIN;
    for (;;) {
      bool more = polling && r == POLL_AGAIN;
      if (!__any_sync(mask, more)) break;
      if (more) r = e.poll_iter(t);
    }
    bool adv = false;
    if (polling) adv = e.run_end(t, r);
    // ---- A: clock advance 50-100 ns, then the timers due by then ----
    ex = act != E::ACT_NONE && e.running();
    if (ex) limit = e.after_run(t, adv);
    for (;;) {
      bool due = ex && e.timer_due(limit);
      if (!__any_sync(mask, due)) break;
      if (due) e.fire_one();
    }
    if (ex) e.clock_to(limit);
    // ---- retire ----
    if (live && !e.running()) {
      e.store();
      SeedResult res;
      e.result(res);
      a.results[idx] = res;
      if (MSIM_UNLIKELY(res.status == MSIM_ST_FAIL_CAPACITY)) atomicAdd(a.cap_fail, 1u);
      live = false;
    }
  }
}

/// Batch summary on the device: counters + pass/fail bitmap (what the N-GPU run all-reduces).
/// out: u64 counters[8] = {ev_pop, ev_fire, draws, n_pass, n_fail, n_unfinished, first_fail_seed, n_seeds}
///      [sharded layout only: u64 first_fail_of_rank[world]]
///      u32 bitmap[...] (bit set = seed failed), this batch's bits starting at bit `bit_off`
/// Local layout (msim_batch_summarize_device): bit_off = 0, slot = nullptr.  Sharded layout
/// (msim_batch_summarize_sharded): the buffer of every rank has the SAME shape and is zero outside the rank's own
/// slot / bit range, so that ONE sum all-reduce of the whole buffer yields the global counters, every rank's first
/// failing seed (a MIN the host takes over `world` slots) and the global bitmap (disjoint bits: sum == or).
__global__ void msim_summary_kernel(const SeedResult* __restrict__ res, uint64_t n, uint64_t seed_base,
                                    unsigned long long* __restrict__ counters, unsigned long long* __restrict__ slot,
                                    uint32_t* __restrict__ bitmap, uint64_t bit_off) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long pop = 0, fire = 0, draws = 0, pass = 0, failn = 0, unf = 0, first = ~0ull;
  bool failed = false;
  if (i < n) {
    SeedResult r = res[i];
    pop = r.ev_pop;
    fire = r.ev_fire;
    draws = r.draws;
    if (r.status == MSIM_ST_PASS)
      pass = 1;
    else if (r.status == MSIM_ST_RUNNING)
      unf = 1;
    else {
      failn = 1;
      failed = true;
      first = seed_base + i;
    }
  }
  uint32_t bits = __ballot_sync(0xffffffffu, failed);
  if ((threadIdx.x & 31) == 0 && i < n && bits) {
    uint64_t g = bit_off + i;  // global index of this warp's first seed
    uint32_t sh = (uint32_t)(g & 31);
    atomicOr(&bitmap[g >> 5], bits << sh);
    if (sh && (bits >> (32 - sh))) atomicOr(&bitmap[(g >> 5) + 1], bits >> (32 - sh));
  }
  for (int o = 16; o > 0; o >>= 1) {
    pop += __shfl_xor_sync(0xffffffffu, pop, o);
    fire += __shfl_xor_sync(0xffffffffu, fire, o);
    draws += __shfl_xor_sync(0xffffffffu, draws, o);
    pass += __shfl_xor_sync(0xffffffffu, pass, o);
    failn += __shfl_xor_sync(0xffffffffu, failn, o);
    unf += __shfl_xor_sync(0xffffffffu, unf, o);
    unsigned long long f2 = __shfl_xor_sync(0xffffffffu, first, o);
    first = f2 < first ? f2 : first;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counters[0], pop);
    atomicAdd(&counters[1], fire);
    atomicAdd(&counters[2], draws);
    atomicAdd(&counters[3], pass);
    atomicAdd(&counters[4], failn);
    atomicAdd(&counters[5], unf);
    if (first != ~0ull) atomicMin(slot ? slot : &counters[6], first);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
thread_local std::string g_err;
int set_err(int code, const std::string& m) {
  g_err = m;
  return code;
}
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t _e = (x);                                                                        \
    if (_e != cudaSuccess)                                                                       \
      return set_err(MSIM_ECUDA, std::string(#x) + ": " + cudaGetErrorString(_e));               \
  } while (0)

template <class WL>
cudaError_t launch_wl(const KernelArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  // two instantiations per program: the batch kernel (dynamic run-to-completion launches), and the general
  // one (bursts through the checkpoint, resume, single-seed replay with trace records)
  auto k = a.dynamic ? (a.lockstep ? msim_step_kernel_ml<WL> : msim_step_kernel<WL, false>) : msim_step_kernel<WL, true>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}

/// the large-capacity variant of a program (engine.h: BigWL) exists as the GENERAL kernel only
template <class WL>
cudaError_t launch_big(const KernelArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  auto k = msim_step_kernel<BigWL<WL>, true>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<grid, block, smem, st>>>(a);
  return cudaGetLastError();
}

int state_bytes_of(uint32_t workload, const uint64_t* p = nullptr, bool big = false) {
  return with_workload(workload, (const u64*)p, -1, [big](auto wl) {
    using W = decltype(wl);
    return big ? (int)Layout<typename BigWL<W>::Caps>::BYTES : (int)Layout<typename W::Caps>::BYTES;
  });
}

}  // namespace

struct msim_batch {
  int device = 0;
  int sm_count = 0;
  int max_smem = 0;
  msim_config cfg{};
  msim_workload_desc wl{};
  bool has_wl = false;
  Params P{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint32_t user_S = 0, user_L = 0;
  bool user_lockstep = false;  // MSIM_GEOMETRY_LOCKSTEP: multi-lane launches go through msim_step_kernel_ml
  // current run
  uint64_t seed_base = 0, n_seeds = 0;
  uint32_t S = 0, L = 0, n_blocks = 0, threads = 0, grid_launched = 0;
  size_t smem = 0;
  bool ran = false, pending = false;
  bool has_checkpoint = false;  // d_blob holds the image of the CURRENT run (a burst run of this geometry)
  SeedResult* d_results = nullptr;
  size_t d_results_cap = 0;
  unsigned char* d_blob = nullptr;
  size_t d_blob_cap = 0;
  uint32_t* d_unfinished = nullptr;  // [3] unfinished seeds, seed queue head, seeds that overflowed a table
  uint32_t* h_unfinished = nullptr;  // pinned copy of the three, refreshed after every run-to-completion launch
  bool escalation = true;            // re-run overflowed seeds on the large-capacity variant (msim_batch_set_capacity_escalation)
  bool always_big = false;           // mode 2 of the same call: every seed of a run to completion on the large variant
  uint32_t escalated = 0;            // seeds of the last run that went through it
  cudaEvent_t ev2 = nullptr, ev3 = nullptr;
  bool esc_pending = false;          // the launch in flight is a run to completion: look at the overflow counter after it
  // a scratch batch of escalate() / replay_one(): the large-capacity variant, optionally over a list of seed offsets of
  // ANOTHER batch whose result array it writes into
  bool big = false;
  const uint32_t* d_seed_list = nullptr;
  SeedResult* results_alias = nullptr;
  SeedResult* h_results = nullptr;   // pinned staging
  size_t h_results_cap = 0;
  unsigned long long* d_summary = nullptr;
  size_t d_summary_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double kernel_ms = 0;
  uint32_t launches = 0;          // step-kernel launches of the last run
  uint32_t summary_launches = 0;  // msim_summary_kernel launches since then (not part of stats.launches)
};

namespace {

int choose_geometry(msim_batch* b, uint64_t n) {
  int bytes = state_bytes_of(b->wl.workload, b->wl.p, b->big);
  if (bytes <= 0) return set_err(MSIM_EINVAL, "unknown workload");
  uint32_t s_max = (uint32_t)((b->max_smem - SMEM_HDR) / bytes);
  if (s_max == 0) return set_err(MSIM_EINVAL, "state block of one seed exceeds shared memory");
  uint32_t S = b->user_S;
  if (S == 0) {
    // one CTA per SM when the batch fits in one wave; otherwise the largest block
    uint64_t per_sm = (n + b->sm_count - 1) / b->sm_count;
    S = (uint32_t)(per_sm < s_max ? per_sm : s_max);
    if (S == 0) S = 1;
    // Lanes of a warp run different seeds, i.e. mostly diverged: populating 2-4 lanes per warp pays
    // only for the small programs (measured, profiles/r1j_geometry_sweeps.txt): STREAM / CHAOS / SIGNAL
    // gain 1.4-1.6x at 4 lanes, RAFT / ETCD / TONIC lose 20-25 % at 2.
    uint32_t lanes = 1;
    switch (b->wl.workload) {
      case MSIM_WL_SLEEP: case MSIM_WL_PINGPONG: case MSIM_WL_STREAM: case MSIM_WL_CHAOS: case MSIM_WL_SIGNAL: lanes = 4; break;
      default: break;
    }
    if (S > lanes * (MAX_THREADS / 32)) S = lanes * (MAX_THREADS / 32);
  }
  if (S > s_max) return set_err(MSIM_EINVAL, "seeds_per_cta exceeds shared memory (" + std::to_string(s_max) + " max)");
  uint32_t L = b->user_L ? b->user_L : 1;  // one seed per warp: lanes never run converged (profiles/), warps do overlap
  if (L > 32) L = 32;
  const uint32_t max_warps = MAX_THREADS / 32;
  if (S > max_warps * 32) return set_err(MSIM_EINVAL, "seeds_per_cta too large for the CTA");
  uint32_t warps = (S + L - 1) / L;
  if (warps > max_warps) {  // populate more lanes per warp instead of more warps
    L = (S + max_warps - 1) / max_warps;
    warps = (S + L - 1) / L;
  }
  b->S = S;
  b->L = L;
  b->threads = warps * 32;
  b->n_blocks = (uint32_t)((n + S - 1) / S);
  b->smem = SMEM_HDR + (size_t)S * bytes;
  return MSIM_OK;
}

int launch(msim_batch* b, uint64_t max_events, bool resume, msim_trace_rec* d_trace, uint32_t* d_trace_n,
           uint32_t trace_cap) {
  KernelArgs a;
  std::memset(&a, 0, sizeof(a));
  a.P = b->P;
  a.seed_base = b->seed_base;
  a.n_seeds = b->n_seeds;
  a.max_events = max_events;
  a.blob = b->has_checkpoint ? b->d_blob : nullptr;  // a stale image of an earlier run is never handed to a kernel
  a.results = b->results_alias ? b->results_alias : b->d_results;
  a.trace = d_trace;
  a.trace_n = d_trace_n;
  a.trace_cap = trace_cap;
  a.S = b->S;
  a.L = b->L;
  a.resume = resume ? 1 : 0;
  a.unfinished = b->d_unfinished;
  a.next_seed = b->d_unfinished + 1;
  a.cap_fail = b->d_unfinished + 2;
  a.seed_list = b->d_seed_list;
  // run-to-completion launches balance the load dynamically; bursts keep seed i in block i / S (the checkpoint layout)
  a.dynamic = (max_events == 0 && !resume && !d_trace && b->n_seeds < 0xffffffffull) ? 1 : 0;
  a.lockstep = (a.dynamic && b->user_lockstep && b->L >= 2) ? 1 : 0;
  CUDA_TRY(cudaMemsetAsync(b->d_unfinished, 0, 3 * sizeof(uint32_t), b->stream));
  uint32_t n_grid = b->n_blocks;
  if (a.dynamic && n_grid > (uint32_t)b->sm_count) n_grid = (uint32_t)b->sm_count;  // 1 CTA per SM (shared memory)
  b->grid_launched = n_grid;
  dim3 grid(n_grid), block(b->threads);
  const cudaError_t unknown = cudaErrorInvalidDeviceFunction;
  cudaError_t e = with_workload(b->wl.workload, (const u64*)a.P.p, unknown, [&](auto wl) {
    using W = decltype(wl);
    return b->big ? launch_big<W>(a, grid, block, b->smem, b->stream) : launch_wl<W>(a, grid, block, b->smem, b->stream);
  });
  if (e == unknown) return set_err(MSIM_EINVAL, "unknown workload");
  if (e != cudaSuccess) return set_err(MSIM_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  b->launches++;
  return MSIM_OK;
}

int ensure_buffers(msim_batch* b, uint64_t n, bool need_blob) {
  if (n > b->d_results_cap) {
    if (b->d_results) cudaFree(b->d_results);
    b->d_results = nullptr;
    CUDA_TRY(cudaMalloc(&b->d_results, n * sizeof(SeedResult)));
    b->d_results_cap = n;
  }
  if (n > b->h_results_cap) {
    if (b->h_results) cudaFreeHost(b->h_results);
    b->h_results = nullptr;
    CUDA_TRY(cudaMallocHost(&b->h_results, n * sizeof(SeedResult)));
    b->h_results_cap = n;
  }
  size_t words = (n + 31) / 32;
  size_t sb = 8 * sizeof(unsigned long long) + words * sizeof(uint32_t);
  if (sb > b->d_summary_cap) {
    if (b->d_summary) cudaFree(b->d_summary);
    b->d_summary = nullptr;
    CUDA_TRY(cudaMalloc(&b->d_summary, sb));
    b->d_summary_cap = sb;
  }
  if (need_blob) {
    size_t bb = (size_t)b->n_blocks * (b->smem - SMEM_HDR);
    if (bb > b->d_blob_cap) {
      if (b->d_blob) cudaFree(b->d_blob);
      b->d_blob = nullptr;
      CUDA_TRY(cudaMalloc(&b->d_blob, bb));
      b->d_blob_cap = bb;
    }
  }
  return MSIM_OK;
}

struct DevBuf {  // cudaMalloc'd scratch, released on every return path
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};
struct BatchGuard {  // a scratch batch, destroyed on every return path
  msim_batch* p;
  ~BatchGuard() { msim_batch_destroy(p); }
};

/// Capacity escalation (engine.h: BigWL).  The seeds of the finished run that ended with MSIM_ST_FAIL_CAPACITY run again
/// on the large-capacity variant of the program and their records replace the failed ones in the batch's result
/// array; a seed that overflows even those tables keeps its MSIM_ST_FAIL_CAPACITY.  Rare by construction — the default
/// tables take the BASELINE configurations without one overflow — so the path is plain: results to the host, the
/// offsets of the failed seeds back, ONE launch of the general kernel (dynamic queue over the list, one seed per warp).
int escalate(msim_batch* b) {
  const uint64_t n = b->n_seeds;
  if (n >= 0xffffffffull) return MSIM_OK;  // (such a run is not a dynamic launch either)
  const int bytes = state_bytes_of(b->wl.workload, b->wl.p, true);
  if (bytes <= 0 || bytes > b->max_smem - SMEM_HDR) return MSIM_OK;  // no larger variant fits an SM: the failures stand
  CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, n * sizeof(SeedResult), cudaMemcpyDeviceToHost, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  std::vector<uint32_t> list;
  for (uint64_t i = 0; i < n; i++)
    if (b->h_results[i].status == MSIM_ST_FAIL_CAPACITY) list.push_back((uint32_t)i);
  if (list.empty()) return MSIM_OK;
  msim_batch* r = nullptr;
  int rc = msim_batch_create(&b->cfg, b->device, &r);
  if (rc) return rc;
  BatchGuard guard{r};
  rc = msim_workload_load(r, &b->wl);
  if (rc) return rc;
  r->escalation = false;
  r->big = true;
  DevBuf d_list;
  CUDA_TRY(cudaMalloc(&d_list.p, list.size() * sizeof(uint32_t)));
  CUDA_TRY(cudaMemcpy(d_list.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  r->d_seed_list = (const uint32_t*)d_list.p;
  r->results_alias = b->d_results;
  r->seed_base = b->seed_base;
  r->n_seeds = list.size();
  uint64_t per_sm = (list.size() + r->sm_count - 1) / r->sm_count;
  uint32_t s_max = (uint32_t)((r->max_smem - SMEM_HDR) / bytes);
  if (s_max > (uint32_t)(MAX_THREADS / 32)) s_max = MAX_THREADS / 32;
  r->user_S = (uint32_t)(per_sm < s_max ? per_sm : s_max);
  r->user_L = 1;
  rc = choose_geometry(r, r->n_seeds);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(r->ev0, r->stream));
  rc = launch(r, 0, false, nullptr, nullptr, 0);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(r->ev1, r->stream));
  CUDA_TRY(cudaStreamSynchronize(r->stream));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  b->kernel_ms += ms;
  b->launches++;
  b->escalated = (uint32_t)list.size();
  return MSIM_OK;
}

int run_impl(msim_batch* b, uint64_t seed_base, uint64_t n, uint64_t max_events, bool sync) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  if (!b->has_wl) return set_err(MSIM_ESTATE, "msim_workload_load must be called before msim_batch_run");
  if (n == 0) return set_err(MSIM_EINVAL, "n_seeds must be > 0");
  CUDA_TRY(cudaSetDevice(b->device));
  b->seed_base = seed_base;
  b->n_seeds = n;
  b->big = false;
  if (b->always_big && max_events == 0) {  // (bursts keep the default variant: the large one has no checkpoint use)
    const int bytes = state_bytes_of(b->wl.workload, b->wl.p, true);
    b->big = bytes > 0 && bytes <= b->max_smem - SMEM_HDR;
  }
  int rc = choose_geometry(b, n);
  if (rc) return rc;
  rc = ensure_buffers(b, n, max_events != 0);
  if (rc) return rc;
  b->has_checkpoint = max_events != 0;  // run to completion keeps no checkpoint; a burst run owns d_blob from here on
  b->launches = 0;
  b->summary_launches = 0;
  CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
  rc = launch(b, max_events, false, nullptr, nullptr, 0);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
  b->escalated = 0;
  b->esc_pending = max_events == 0;
  if (b->esc_pending)
    CUDA_TRY(cudaMemcpyAsync(b->h_unfinished, b->d_unfinished, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
  b->ran = true;
  b->pending = true;
  if (sync) return msim_batch_sync(b);
  return MSIM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int msim_version(void) { return MSIM_VERSION; }
const char* msim_last_error(void) { return g_err.c_str(); }

int msim_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) return set_err(MSIM_ENODEV, "no CUDA device: libmsim_b200 has no CPU fallback");
  return n;
}

int msim_config_default(msim_config* cfg) {
  if (!cfg) return set_err(MSIM_EINVAL, "null config");
  config_defaults(cfg);
  return MSIM_OK;
}

int msim_config_from_toml(const char* text, msim_config* cfg) {
  if (!text || !cfg) return set_err(MSIM_EINVAL, "null argument");
  std::string err;
  int rc = config_from_toml(text, cfg, &err);
  return rc ? set_err(rc, err) : MSIM_OK;
}

int msim_workload_default(uint32_t workload, msim_workload_desc* out) {
  if (!out) return set_err(MSIM_EINVAL, "null descriptor");
  if (workload_defaults(workload, out)) return set_err(MSIM_EINVAL, "unknown workload");
  return MSIM_OK;
}

int msim_batch_create(const msim_config* cfg, int device, msim_batch** out) {
  if (!out) return set_err(MSIM_EINVAL, "null out pointer");
  *out = nullptr;
  int n = msim_device_count();
  if (n < 0) return n;
  if (device < 0 || device >= n) return set_err(MSIM_EINVAL, "bad device ordinal");
  CUDA_TRY(cudaSetDevice(device));
  struct Guard {  // an early return releases whatever has been created so far
    msim_batch* p;
    ~Guard() {
      if (p) msim_batch_destroy(p);
    }
  } guard{new msim_batch()};
  msim_batch* b = guard.p;
  b->device = device;
  if (cfg)
    b->cfg = *cfg;
  else
    config_defaults(&b->cfg);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  b->sm_count = prop.multiProcessorCount;
  b->max_smem = (int)prop.sharedMemPerBlockOptin;
  CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  b->own_stream = true;
  CUDA_TRY(cudaEventCreate(&b->ev0));
  CUDA_TRY(cudaEventCreate(&b->ev1));
  CUDA_TRY(cudaMalloc(&b->d_unfinished, 3 * sizeof(uint32_t)));
  CUDA_TRY(cudaMallocHost(&b->h_unfinished, 3 * sizeof(uint32_t)));
  CUDA_TRY(cudaEventCreate(&b->ev2));
  CUDA_TRY(cudaEventCreate(&b->ev3));
  guard.p = nullptr;
  *out = b;
  return MSIM_OK;
}

void msim_batch_destroy(msim_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  if (b->d_results) cudaFree(b->d_results);
  if (b->d_blob) cudaFree(b->d_blob);
  if (b->d_unfinished) cudaFree(b->d_unfinished);
  if (b->d_summary) cudaFree(b->d_summary);
  if (b->h_unfinished) cudaFreeHost(b->h_unfinished);
  if (b->h_results) cudaFreeHost(b->h_results);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev2) cudaEventDestroy(b->ev2);
  if (b->ev3) cudaEventDestroy(b->ev3);
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int msim_workload_load(msim_batch* b, const msim_workload_desc* w) {
  if (!b || !w) return set_err(MSIM_EINVAL, "null argument");
  Params P;
  std::string err;
  int rc = make_params(b->cfg, *w, &P, &err);
  if (rc) return set_err(rc, err);
  if (state_bytes_of(w->workload, w->p) <= 0) return set_err(MSIM_EINVAL, "workload not available on the device");
  b->wl = *w;
  b->wl.n_nodes = P.n_nodes;
  b->P = P;
  b->has_wl = true;
  b->ran = false;
  b->has_checkpoint = false;
  return MSIM_OK;
}

int msim_batch_set_geometry(msim_batch* b, uint32_t seeds_per_cta, uint32_t lanes_per_warp) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  const bool lockstep = lanes_per_warp & MSIM_GEOMETRY_LOCKSTEP;
  lanes_per_warp &= ~(uint32_t)MSIM_GEOMETRY_LOCKSTEP;
  if (lanes_per_warp > 32) return set_err(MSIM_EINVAL, "lanes_per_warp must be 0..32");
  b->user_S = seeds_per_cta;
  b->user_L = lanes_per_warp;
  b->user_lockstep = lockstep;
  return MSIM_OK;
}

int msim_batch_set_capacity_escalation(msim_batch* b, int on) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  if (on < 0 || on > 2) return set_err(MSIM_EINVAL, "capacity escalation mode must be 0, 1 or 2");
  b->escalation = on != 0;
  b->always_big = on == 2;
  return MSIM_OK;
}

int msim_batch_set_stream(msim_batch* b, void* cuda_stream) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  if (b->own_stream && b->stream) {
    cudaStreamSynchronize(b->stream);
    cudaStreamDestroy(b->stream);
  }
  b->stream = (cudaStream_t)cuda_stream;
  b->own_stream = false;
  return MSIM_OK;
}

int msim_batch_sync(msim_batch* b) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  if (!b->pending) return MSIM_OK;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->kernel_ms = ms;
  b->pending = false;
  if (b->esc_pending) {
    b->esc_pending = false;
    if (b->escalation && !b->big && b->h_unfinished[2] > 0) return escalate(b);
  }
  return MSIM_OK;
}

int msim_batch_run(msim_batch* b, uint64_t seed_base, uint64_t n_seeds, uint64_t max_events_per_seed) {
  return run_impl(b, seed_base, n_seeds, max_events_per_seed, true);
}
int msim_batch_run_async(msim_batch* b, uint64_t seed_base, uint64_t n_seeds, uint64_t max_events_per_seed) {
  return run_impl(b, seed_base, n_seeds, max_events_per_seed, false);
}

int msim_batch_resume(msim_batch* b, uint64_t max_events_per_seed) {
  if (!b || !b->ran) return set_err(MSIM_ESTATE, "nothing to resume");
  if (!b->has_checkpoint || !b->d_blob)
    return set_err(MSIM_ESTATE, "the previous run kept no checkpoint (max_events_per_seed was 0)");
  CUDA_TRY(cudaSetDevice(b->device));
  int rc = msim_batch_sync(b);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
  rc = launch(b, max_events_per_seed, true, nullptr, nullptr, 0);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
  b->pending = true;
  double before = b->kernel_ms;
  rc = msim_batch_sync(b);
  b->kernel_ms += before;
  return rc;
}

int msim_batch_results_device(msim_batch* b, void** dptr, size_t* bytes) {
  if (!b || !b->ran) return set_err(MSIM_ESTATE, "no run yet");
  int rc = msim_batch_sync(b);  // the records are final only after the run (and a capacity escalation, if any) completed
  if (rc) return rc;
  if (dptr) *dptr = b->d_results;
  if (bytes) *bytes = b->n_seeds * sizeof(SeedResult);
  return MSIM_OK;
}

/// Device-side summary into caller-provided DEVICE memory: u64 counters[8] + u32 bitmap[ceil(n/32)].
int msim_batch_summarize_device(msim_batch* b, void* d_out, size_t bytes) {
  if (!b || !b->ran) return set_err(MSIM_ESTATE, "no run yet");
  size_t words = (b->n_seeds + 31) / 32;
  size_t need = 8 * sizeof(unsigned long long) + words * sizeof(uint32_t);
  if (!d_out || bytes < need) return set_err(MSIM_ERANGE, "summary buffer too small");
  CUDA_TRY(cudaSetDevice(b->device));
  int rc_sync = msim_batch_sync(b);  // the run is complete on the host's side (capacity escalation may follow it)
  if (rc_sync) return rc_sync;
  unsigned long long init[8] = {0, 0, 0, 0, 0, 0, ~0ull, (unsigned long long)b->n_seeds};
  CUDA_TRY(cudaMemsetAsync(d_out, 0, need, b->stream));
  CUDA_TRY(cudaMemcpyAsync(d_out, init, sizeof(init), cudaMemcpyHostToDevice, b->stream));
  auto* counters = (unsigned long long*)d_out;
  auto* bitmap = (uint32_t*)((unsigned char*)d_out + sizeof(init));
  unsigned threads = 256, blocks = (unsigned)((b->n_seeds + threads - 1) / threads);
  msim_summary_kernel<<<blocks, threads, 0, b->stream>>>(b->d_results, b->n_seeds, b->seed_base, counters, nullptr, bitmap, 0);
  CUDA_TRY(cudaGetLastError());
  b->summary_launches++;
  return MSIM_OK;
}

/// The summary of one shard of an N-GPU sweep, shaped for ONE sum all-reduce (the reference's
/// sweep loop is runtime/builder.rs:129-150): u64 counters[8] (slot 6 unused, 0), u64 first_fail_of_rank[world]
/// (own slot: first failing seed or all-ones; other slots 0), u32 bitmap[ceil(n_total/32)] with this shard's bits at
/// (first seed of the shard - seed_base) and zeros elsewhere.  A rank whose shard is empty calls this after no run
/// at all: it contributes zeros (and all-ones in its slot).
int msim_batch_summarize_sharded(msim_batch* b, void* d_out, size_t bytes, uint32_t rank, uint32_t world,
                                 uint64_t seed_base, uint64_t n_total) {
  if (!b) return set_err(MSIM_EINVAL, "null batch");
  if (world == 0 || rank >= world) return set_err(MSIM_EINVAL, "bad rank / world");
  const bool empty = !b->ran;  // an empty shard never ran
  uint64_t n = empty ? 0 : b->n_seeds;
  if (!empty && (b->seed_base < seed_base || b->seed_base - seed_base + n > n_total))
    return set_err(MSIM_ERANGE, "the batch's seeds are not inside [seed_base, seed_base + n_total)");
  size_t words = (n_total + 31) / 32;
  size_t need = (8 + (size_t)world) * sizeof(unsigned long long) + words * sizeof(uint32_t);
  if (!d_out || bytes < need) return set_err(MSIM_ERANGE, "summary buffer too small");
  CUDA_TRY(cudaSetDevice(b->device));
  int rc_sync = msim_batch_sync(b);  // the run is complete on the host's side (capacity escalation may follow it)
  if (rc_sync) return rc_sync;
  CUDA_TRY(cudaMemsetAsync(d_out, 0, need, b->stream));
  auto* counters = (unsigned long long*)d_out;
  auto* slot = counters + 8 + rank;
  auto* bitmap = (uint32_t*)(counters + 8 + world);
  CUDA_TRY(cudaMemsetAsync(slot, 0xff, sizeof(unsigned long long), b->stream));
  if (n) {
    unsigned long long nn = n;
    CUDA_TRY(cudaMemcpyAsync(counters + 7, &nn, sizeof(nn), cudaMemcpyHostToDevice, b->stream));
    unsigned threads = 256, blocks = (unsigned)((n + threads - 1) / threads);
    msim_summary_kernel<<<blocks, threads, 0, b->stream>>>(b->d_results, n, b->seed_base, counters, slot, bitmap,
                                                           b->seed_base - seed_base);
    CUDA_TRY(cudaGetLastError());
    b->summary_launches++;
  }
  return MSIM_OK;
}

int msim_batch_results(msim_batch* b, msim_seed_result* out, uint32_t* fail_bitmap, msim_stats* stats) {
  if (!b || !b->ran) return set_err(MSIM_ESTATE, "no run yet");
  CUDA_TRY(cudaSetDevice(b->device));
  int rc = msim_batch_sync(b);
  if (rc) return rc;
  uint64_t n = b->n_seeds;
  size_t words = (n + 31) / 32;
  std::vector<unsigned long long> h(8 + (words + 1) / 2 + 1);
  if (fail_bitmap || stats) {
    rc = msim_batch_summarize_device(b, b->d_summary, b->d_summary_cap);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h.data(), b->d_summary, 8 * sizeof(unsigned long long) + words * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, b->stream));
  }
  if (out) {
    CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, n * sizeof(SeedResult), cudaMemcpyDeviceToHost, b->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  if (out) std::memcpy(out, b->h_results, n * sizeof(SeedResult));
  if (fail_bitmap) std::memcpy(fail_bitmap, h.data() + 8, words * sizeof(uint32_t));
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_seeds = n;
    stats->ev_pop = h[0];
    stats->ev_fire = h[1];
    stats->draws = h[2];
    stats->n_pass = h[3];
    stats->n_fail = h[4];
    stats->n_unfinished = h[5];
    stats->first_fail_seed = h[6];
    stats->kernel_ms = b->kernel_ms;
    stats->launches = b->launches;  // step-kernel launches (the summary kernel is not counted)
    stats->seeds_per_cta = b->S;
    stats->lanes_per_warp = b->L;
    stats->ctas = b->grid_launched ? b->grid_launched : b->n_blocks;
    stats->smem_bytes = (uint32_t)b->smem;
    stats->escalated = b->escalated;
  }
  return MSIM_OK;
}

int msim_batch_trace_hash(msim_batch* b, uint64_t* hash) {
  if (!b || !b->ran) return set_err(MSIM_ESTATE, "no run yet");
  if (!hash) return set_err(MSIM_EINVAL, "null output");
  int rc = msim_batch_sync(b);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, b->n_seeds * sizeof(SeedResult), cudaMemcpyDeviceToHost,
                           b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  for (uint64_t i = 0; i < b->n_seeds; i++) hash[i] = b->h_results[i].trace_hash;
  return MSIM_OK;
}

namespace {
/// Replays ONE seed in a scratch batch that shares config + workload, with the canonical trace and / or the
/// determinism log switched on (both only exist in the GENERAL kernels).
int replay_pass(msim_batch* b, bool big, uint64_t seed, msim_trace_rec* tr_out, size_t tr_cap, size_t* tr_n, uint8_t* log_out,
                size_t log_cap, size_t* log_n, bool want_log, msim_seed_result* result) {
  msim_batch* r = nullptr;
  int rc = msim_batch_create(&b->cfg, b->device, &r);
  if (rc) return rc;
  BatchGuard guard{r};
  rc = msim_workload_load(r, &b->wl);
  if (rc) return rc;
  r->big = big;
  r->user_S = 1;
  r->user_L = 1;
  r->seed_base = seed;
  r->n_seeds = 1;
  rc = choose_geometry(r, 1);
  if (rc) return rc;
  rc = ensure_buffers(r, 1, false);
  if (rc) return rc;
  DevBuf d_tr, d_n, d_log;
  CUDA_TRY(cudaMalloc(&d_tr.p, (tr_cap ? tr_cap : 1) * sizeof(msim_trace_rec)));
  CUDA_TRY(cudaMalloc(&d_n.p, sizeof(uint32_t)));
  CUDA_TRY(cudaMemsetAsync(d_n.p, 0, sizeof(uint32_t), r->stream));
  if (want_log) {  // {u32 count, u32 capacity, bytes}
    CUDA_TRY(cudaMalloc(&d_log.p, 8 + (log_cap ? log_cap : 1)));
    uint32_t hdr[2] = {0, (uint32_t)log_cap};
    CUDA_TRY(cudaMemcpyAsync(d_log.p, hdr, sizeof(hdr), cudaMemcpyHostToDevice, r->stream));
    r->P.dlog = (unsigned char*)d_log.p;
  }
  rc = launch(r, 0, false, (msim_trace_rec*)d_tr.p, (uint32_t*)d_n.p, (uint32_t)tr_cap);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(r->stream));
  uint32_t hn = 0, hdr[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(&hn, d_n.p, sizeof(hn), cudaMemcpyDeviceToHost));
  if (tr_out && tr_cap)
    CUDA_TRY(cudaMemcpy(tr_out, d_tr.p, (hn < tr_cap ? hn : tr_cap) * sizeof(msim_trace_rec), cudaMemcpyDeviceToHost));
  if (want_log) {
    CUDA_TRY(cudaMemcpy(hdr, d_log.p, sizeof(hdr), cudaMemcpyDeviceToHost));
    if (log_out && log_cap)
      CUDA_TRY(cudaMemcpy(log_out, (unsigned char*)d_log.p + 8, hdr[0] < log_cap ? hdr[0] : log_cap, cudaMemcpyDeviceToHost));
  }
  CUDA_TRY(cudaMemcpy(result, r->d_results, sizeof(SeedResult), cudaMemcpyDeviceToHost));
  if (tr_n) *tr_n = hn;
  if (log_n) *log_n = hdr[0];
  return MSIM_OK;
}
int replay_one(msim_batch* b, uint64_t seed, msim_trace_rec* tr_out, size_t tr_cap, size_t* tr_n, uint8_t* log_out,
               size_t log_cap, size_t* log_n, bool want_log, msim_seed_result* result) {
  if (!b || !b->has_wl) return set_err(MSIM_ESTATE, "msim_workload_load must be called first");
  if (tr_cap > 0xffffffffull || log_cap > 0xfffffff0ull) return set_err(MSIM_ERANGE, "capacity too large");
  CUDA_TRY(cudaSetDevice(b->device));
  int rc = msim_batch_sync(b);
  if (rc) return rc;
  msim_seed_result res;
  rc = replay_pass(b, false, seed, tr_out, tr_cap, tr_n, log_out, log_cap, log_n, want_log, &res);
  if (rc) return rc;
  const int big_bytes = state_bytes_of(b->wl.workload, b->wl.p, true);
  if (res.status == MSIM_ST_FAIL_CAPACITY && b->escalation && big_bytes > 0 && big_bytes <= b->max_smem - SMEM_HDR) {
    // the same escalation a batch run applies: the replay of an overflowed seed comes from the large-capacity variant
    rc = replay_pass(b, true, seed, tr_out, tr_cap, tr_n, log_out, log_cap, log_n, want_log, &res);
    if (rc) return rc;
  }
  if (result) *result = res;
  return MSIM_OK;
}
}  // namespace

int msim_batch_trace(msim_batch* b, uint64_t seed, msim_trace_rec* out, size_t cap, size_t* n, msim_seed_result* result) {
  return replay_one(b, seed, out, cap, n, nullptr, 0, nullptr, false, result);
}

/// The log Runtime::check_determinism records and checks (runtime/mod.rs:178-202, rand.rs:64-89): one byte per
/// GlobalRng::with call, (next u8 of a clone of the rng) ^ xor-fold(elapsed nanos).  Device replay of one seed.
int msim_batch_determinism_log(msim_batch* b, uint64_t seed, uint8_t* out, size_t cap, size_t* n, msim_seed_result* result) {
  return replay_one(b, seed, nullptr, 0, nullptr, out, cap, n, true, result);
}

/// Bytes of per-seed state for a workload's default variant (DESIGN.md capacity table; bench roofline accounting)
int msim_workload_state_bytes(uint32_t workload) {
  int b = state_bytes_of(workload);
  return b > 0 ? b : set_err(MSIM_EINVAL, "unknown workload");
}

}  // extern "C"
