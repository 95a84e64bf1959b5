// Every registered program, and the ONE place that maps (workload id, parameters) to its state-machine type.
// The kernels, the launch code and the host test harness all go through with_workload(); registering a program
// means: its header here, its defaults / validation in engine/host_params.h, its id in include/msim_b200.h.
#pragma once
#include "chaos.h"
#include "election.h"
#include "etcd.h"
#include "pingpong.h"
#include "raft.h"
#include "signal.h"
#include "sleep.h"
#include "stream.h"
#include "tonic.h"

namespace msim {

/// f(WL{}) for the program (workload, p[]) selects — p2 picks the tonic variant, p7 the etcd one; p may be null (the
/// default variant); `otherwise` for an unknown id
template <class F, class R>
inline R with_workload(uint32_t workload, const u64* p, R otherwise, F&& f) {
  const u64 p2 = p ? p[2] : 0, p7 = p ? p[7] : 0;
  (void)p2, (void)p7;
#if defined(MSIM_ONLY_WL)  // kernel experiments: build one program only (seconds instead of minutes)
  if (workload == MSIM_ONLY_WL) return f(MSIM_ONLY_WL_TYPE{});
  return otherwise;
#else
  switch (workload) {
    case MSIM_WL_SLEEP: return f(WL_Sleep{});
    case MSIM_WL_PINGPONG: return f(WL_PingPong{});
    case MSIM_WL_RAFT: return f(WL_Raft{});
    case MSIM_WL_CHAOS: return f(WL_Chaos{});
    case MSIM_WL_STREAM: return f(WL_Stream{});
    case MSIM_WL_ETCD: return f(WL_Etcd{});  // (the KV-only instantiation WL_EtcdKv measured 5 % SLOWER: profiles/r2e_etcd_variants.txt)
    case MSIM_WL_TONIC: return tonic_variant(p2, f);
    case MSIM_WL_SIGNAL: return f(WL_Signal{});
    case MSIM_WL_ELECTION: return f(WL_Election{});
  }
  return otherwise;
#endif
}

}  // namespace msim
