// TEST HARNESS ONLY — never part of libmsim_b200.so.
// Compiles the device engine + workload state machines (madsim_b200/csrc/engine, workloads)
// with g++ and runs seeds one after another on the host, so that `pytest -m "not gpu"` can
// diff the state machines against the oracle here, where there is no GPU.  The product path
// is CUDA-only (kernel.cu / capi.cu) and fails with MSIM_ENODEV without a device.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../madsim_b200/csrc/engine/host_params.h"
#include "../../madsim_b200/csrc/workloads/all.h"

using namespace msim;

template <class WL>
static int run_wl(const Params& P, uint64_t seed_base, uint64_t n, uint32_t S, uint64_t max_events,
                  msim_seed_result* out, uint64_t trace_seed, msim_trace_rec* tr, size_t cap, size_t* n_tr) {
  using E = Engine<WL>;
  if (S == 0) S = 1;
  std::vector<unsigned char> block((size_t)S * E::L::BYTES + 64);
  for (uint64_t blk = 0; blk * S < n; blk++) {
    std::memset(block.data(), 0xA5, block.size());  // catch reads of uninitialised columns
    for (uint32_t s = 0; s < S && blk * S + s < n; s++) {
      uint64_t i = blk * S + s;
      E e;
      std::memset(&e, 0, sizeof(e));
      e.sb = block.data() + (size_t)s * E::L::BYTES;
      e.P = &P;
      bool want_trace = tr && (seed_base + i) == trace_seed;
      e.tr = want_trace ? tr : nullptr;
      e.tr_cap = want_trace ? (u32)cap : 0;
      e.tr_n = 0;
      e.init(seed_base + i);
      uint64_t budget = max_events ? max_events : ~0ull;
      // burst / checkpoint emulation: spill registers to the block every `budget` events
      for (;;) {
        uint64_t ev = 0;
        while (ev < budget && e.step()) ev++;
        e.store();
        if (e.status != MSIM_ST_RUNNING) break;
        e.load();
      }
      e.result(out[i]);
      if (want_trace && n_tr) *n_tr = e.tr_n;
    }
  }
  return 0;
}

extern "C" int enginecheck_run(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed_base, uint64_t n,
                               uint32_t S, uint64_t max_events, msim_seed_result* out, uint64_t trace_seed,
                               msim_trace_rec* tr, size_t cap, size_t* n_tr) {
  Params P;
  std::string err;
  int rc = make_params(*cfg, *wl, &P, &err);
  if (rc) return rc;
  switch (wl->workload) {
    case MSIM_WL_SLEEP: return run_wl<WL_Sleep>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_PINGPONG: return run_wl<WL_PingPong>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_RAFT: return run_wl<WL_Raft>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_CHAOS: return run_wl<WL_Chaos>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_STREAM: return run_wl<WL_Stream>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_ETCD: return run_wl<WL_Etcd>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_TONIC:
      return (P.p[2] & 8) ? run_wl<WL_TonicBig>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr)
                          : run_wl<WL_Tonic>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
    case MSIM_WL_SIGNAL: return run_wl<WL_Signal>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
  }
  return MSIM_EINVAL;
}

extern "C" int enginecheck_state_bytes(uint32_t workload) {
  switch (workload) {
    case MSIM_WL_SLEEP: return Layout<WL_Sleep::Caps>::BYTES;
    case MSIM_WL_PINGPONG: return Layout<WL_PingPong::Caps>::BYTES;
    case MSIM_WL_RAFT: return Layout<WL_Raft::Caps>::BYTES;
    case MSIM_WL_CHAOS: return Layout<WL_Chaos::Caps>::BYTES;
    case MSIM_WL_STREAM: return Layout<WL_Stream::Caps>::BYTES;
    case MSIM_WL_ETCD: return Layout<WL_Etcd::Caps>::BYTES;
    case MSIM_WL_TONIC: return Layout<WL_Tonic::Caps>::BYTES;
    case MSIM_WL_SIGNAL: return Layout<WL_Signal::Caps>::BYTES;
  }
  return -1;
}
