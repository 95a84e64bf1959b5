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

// Host emulation of msim_step_kernel_ml (the multi-lane batch kernel): `lanes` engines advance in lock step through
// the same phases, a vote being a loop over the lanes.  Checks that the phase decomposition executes, per seed,
// exactly the steps Engine::step() would.
template <class WL>
static int run_wl_ml(const Params& P, uint64_t seed_base, uint64_t n, uint32_t lanes, msim_seed_result* out) {
  using E = Engine<WL, false>;
  using LY = typename E::L;
  struct Lane {
    E e;
    bool live = false, exhausted = false, cont = false, ex = false, polling = false, adv = false;
    uint64_t idx = 0, limit = 0;
    int t = 0, act = 0, r = 0;
  };
  std::vector<unsigned char> block((size_t)lanes * LY::BYTES + 64, 0xA5);
  std::vector<Lane> ln(lanes);
  for (uint32_t i = 0; i < lanes; i++) {
    std::memset(&ln[i].e, 0, sizeof(E));
    ln[i].e.sb = block.data() + (size_t)i * LY::BYTES;
    ln[i].e.P = &P;
    ln[i].e.status = MSIM_ST_PASS;
  }
  uint64_t next = 0;
  auto any = [&](auto pred) {
    for (auto& l : ln)
      if (pred(l)) return true;
    return false;
  };
  auto expire_phase = [&]() {
    for (;;) {
      bool some = false;
      for (auto& l : ln) {
        bool due = l.ex && l.e.timer_due(l.limit);
        if (due) l.e.fire_one(), some = true;
      }
      if (!some) break;
    }
    for (auto& l : ln)
      if (l.ex) l.e.clock_to(l.limit);
  };
  for (;;) {
    for (auto& l : ln)
      if (!l.live && !l.exhausted) {
        l.idx = next++;
        if (l.idx < n) {
          l.e.init(seed_base + l.idx);
          l.live = true;
        } else {
          l.exhausted = true;
        }
      }
    if (!any([](Lane& l) { return l.live; })) break;
    for (auto& l : ln) {
      l.limit = 0;
      l.ex = false;
      l.cont = l.live && l.e.w(LY::W_RESUME) != NONE32;
      if (l.live && !l.cont && l.e.rq_len == 0) l.ex = l.e.jump_limit(l.limit);
    }
    expire_phase();
    for (auto& l : ln) {
      l.t = 0;
      l.act = E::ACT_NONE;
      if (l.live && l.e.running() && (l.cont || l.e.rq_len > 0)) {
        if (l.cont) {
          l.t = (int)l.e.w(LY::W_RESUME);
          l.e.w(LY::W_RESUME) = NONE32;
          l.act = E::ACT_RUN;
        } else {
          l.act = l.e.pop_ready(l.t);
        }
      }
    }
    for (auto& l : ln) {
      l.polling = false;
      if (l.act == E::ACT_DROP) l.e.drop_runnable(l.t);
      if (l.act == E::ACT_RUN) l.polling = l.e.run_begin(l.t, l.cont);
      l.r = POLL_AGAIN;
    }
    while (any([](Lane& l) { return l.polling && l.r == POLL_AGAIN; }))
      for (auto& l : ln)
        if (l.polling && l.r == POLL_AGAIN) l.r = l.e.poll_iter(l.t);
    for (auto& l : ln) {
      l.adv = false;
      if (l.polling) l.adv = l.e.run_end(l.t, l.r);
      l.ex = l.act != E::ACT_NONE && l.e.running();
      if (l.ex) l.limit = l.e.after_run(l.t, l.adv);
    }
    expire_phase();
    for (auto& l : ln)
      if (l.live && !l.e.running()) {
        l.e.store();
        l.e.result(out[l.idx]);
        l.live = false;
      }
  }
  return 0;
}

extern "C" int enginecheck_run_ml(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed_base, uint64_t n,
                                  uint32_t lanes, msim_seed_result* out) {
  Params P;
  std::string err;
  int rc = make_params(*cfg, *wl, &P, &err);
  if (rc) return rc;
  if (lanes == 0) lanes = 1;
  return with_workload(wl->workload, P.p, (int)MSIM_EINVAL, [&](auto w) { return run_wl_ml<decltype(w)>(P, seed_base, n, lanes, out); });
}

extern "C" int enginecheck_run(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed_base, uint64_t n,
                               uint32_t S, uint64_t max_events, msim_seed_result* out, uint64_t trace_seed,
                               msim_trace_rec* tr, size_t cap, size_t* n_tr) {
  Params P;
  std::string err;
  int rc = make_params(*cfg, *wl, &P, &err);
  if (rc) return rc;
  return with_workload(wl->workload, P.p, (int)MSIM_EINVAL, [&](auto w) {
    return run_wl<decltype(w)>(P, seed_base, n, S, max_events, out, trace_seed, tr, cap, n_tr);
  });
}

/// The same seeds on the large-capacity variant of the program (engine.h: BigWL — what capacity escalation re-runs
/// overflowed seeds on).
extern "C" int enginecheck_run_big(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed_base, uint64_t n,
                                   msim_seed_result* out, uint64_t trace_seed, msim_trace_rec* tr, size_t cap, size_t* n_tr) {
  Params P;
  std::string err;
  int rc = make_params(*cfg, *wl, &P, &err);
  if (rc) return rc;
  return with_workload(wl->workload, P.p, (int)MSIM_EINVAL, [&](auto w) {
    return run_wl<BigWL<decltype(w)>>(P, seed_base, n, 1, 0, out, trace_seed, tr, cap, n_tr);
  });
}
extern "C" int enginecheck_state_bytes_big(uint32_t workload) {
  return with_workload(workload, (const u64*)nullptr, -1,
                       [](auto w) { return (int)Layout<typename BigWL<decltype(w)>::Caps>::BYTES; });
}

/// The determinism log (rand.rs:64-89) as the replay kernel produces it, from the host build of the same engine.
extern "C" int enginecheck_determinism_log(const msim_config* cfg, const msim_workload_desc* wl, uint64_t seed, uint8_t* out,
                                           size_t cap, size_t* n, msim_seed_result* result) {
  Params P;
  std::string err;
  int rc = make_params(*cfg, *wl, &P, &err);
  if (rc) return rc;
  std::vector<unsigned char> buf(8 + cap + 8);
  uint32_t hdr[2] = {0, (uint32_t)cap};
  std::memcpy(buf.data(), hdr, 8);
  P.dlog = buf.data();
  size_t ntr = 0;
  rc = with_workload(wl->workload, P.p, (int)MSIM_EINVAL,
                     [&](auto w) { return run_wl<decltype(w)>(P, seed, 1, 1, 0, result, ~0ull, nullptr, 0, &ntr); });
  std::memcpy(hdr, buf.data(), 8);
  if (n) *n = hdr[0];
  if (out) std::memcpy(out, buf.data() + 8, hdr[0] < cap ? hdr[0] : cap);
  return rc;
}

extern "C" int enginecheck_state_bytes(uint32_t workload) {
  static const u64 streaming[8] = {0, 0, 2, 0, 0, 0, 0, 1};
  return with_workload(workload, streaming, -1, [](auto w) { return (int)Layout<typename decltype(w)::Caps>::BYTES; });
}
