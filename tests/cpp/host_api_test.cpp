// C++ host-API test: the reference's own tests for the hot path, written the way they read in
// Rust, against include/msim_b200.hpp (→ C-ABI → CUDA).  Run by tests/test_gpu_parity.py on the GPU
// box; prints one line per case and exits non-zero on failure.
//   connect_send_recv        net/endpoint.rs:550-584 (BASELINE config 1)
//   kill / restart / pause   task/mod.rs:858-1016 (folded into the CHAOS schedule)
//   restart_on_panic         task/mod.rs:939-963
//   random interleavings     task/mod.rs:1018-1041 (10 seeds ⇒ 10 distinct traces)
//   deadlock / time limit    task/mod.rs:245-258
//   seed sweep + repro note  runtime/builder.rs:129-159, runtime/mod.rs:205-210
#include <cstdio>
#include <cstring>
#include <set>

#include "../../include/msim_b200.hpp"

using namespace madsim;
using madsim::runtime::Builder;
using madsim::runtime::Runtime;

static int failed = 0;
#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      std::printf("  EXPECT FAILED line %d: %s\n", __LINE__, #c); \
      failed++;                                               \
    }                                                         \
  } while (0)

int main() {
  if (msim_device_count() < 1) {
    std::printf("no CUDA device: %s\n", msim_last_error());
    return 2;
  }
  {  // connect_send_recv: Runtime::new().block_on(..)
    Runtime rt;
    msim_seed_result r = rt.block_on(Workload::pingpong(1));
    EXPECT(r.status == MSIM_ST_PASS && r.user[0] == 1);
    std::printf("ok   connect_send_recv (1 round): %u events, elapsed %.6f s\n", r.ev_pop + r.ev_fire, r.elapsed_ns / 1e9);
  }
  {  // BASELINE config 1: R = 1000 rounds, seed 0
    msim_seed_result r = Runtime::with_seed_and_config(0, Config{}).block_on(Workload::pingpong(1000));
    EXPECT(r.user[0] == 1000);
    EXPECT((r.ev_pop + r.ev_fire) / 1000 == 14);  // ≈ 14 events per round
    std::printf("ok   pingpong R=1000: %u events\n", r.ev_pop + r.ev_fire);
  }
  {  // different seed ⇒ different interleaving; same seed ⇒ same
    std::set<uint64_t> hashes;
    for (uint64_t seed = 0; seed < 10; seed++)
      hashes.insert(Runtime::with_seed_and_config(seed, Config{}).block_on(Workload::pingpong(20)).trace_hash);
    EXPECT(hashes.size() == 10);
    Runtime::check_determinism(7, Config{}, Workload::raft(5, 2000, 7));
    std::printf("ok   10 seeds => 10 traces; check_determinism\n");
  }
  {  // "no events, all tasks will block forever" when every packet is lost
    Config cfg;
    cfg.net.packet_loss_rate = 1.0;
    bool deadlock = false;
    try {
      Runtime::with_seed_and_config(3, cfg).block_on(Workload::pingpong(3));
    } catch (SimulationFailure& e) {
      deadlock = e.status == MSIM_ST_FAIL_DEADLOCK && std::strstr(e.what(), "MADSIM_TEST_SEED=3");
    }
    EXPECT(deadlock);
    Runtime rt = Runtime::with_seed_and_config(0, Config{});
    rt.set_time_limit(2000000000ull);
    bool limit = false;
    try {
      rt.block_on(Workload::pingpong(1000));
    } catch (SimulationFailure& e) {
      limit = e.status == MSIM_ST_FAIL_TIME_LIMIT;
    }
    EXPECT(limit);
    std::printf("ok   deadlock panic, time limit\n");
  }
  {  // kill / restart / pause / resume / restart_on_panic schedule
    Builder b;
    b.count = 256;
    auto out = b.run(Workload::chaos());
    EXPECT(out.stats.n_pass == 256);
    // under `check` only `seed` runs, twice, through Runtime::check_determinism (builder.rs:126-128)
    auto out2 = Builder{0, 64, 1, Config{}, 0, true}.run(Workload::chaos(4, 3000, 50));
    EXPECT(out2.results.size() == 1 && out2.results[0].status == MSIM_ST_PASS);
    // MADSIM_TEST_JOBS: the sweep sharded over (up to) that many GPUs gives the same results
    Builder jb;
    jb.seed = 9, jb.count = 100, jb.jobs = 4;
    auto out3 = jb.run(Workload::chaos());
    Builder one;
    one.seed = 9, one.count = 100;
    auto out4 = one.run(Workload::chaos());
    bool same = out3.results.size() == 100 && out4.results.size() == 100;
    for (size_t i = 0; same && i < 100; i++) same = out3.results[i].trace_hash == out4.results[i].trace_hash;
    EXPECT(same);
    // MADSIM_TEST_CONFIG: TOML → Config
    Config tc = Config::from_toml("[net]\npacket_loss_rate = 0.5\n");
    EXPECT(tc.net.packet_loss_rate == 0.5 && tc.net.send_latency_lo_ns == 1000000);
    std::printf("ok   chaos: %llu seeds pass, worker restarted %u times in seed 0\n",
                (unsigned long long)out.stats.n_pass, out.results[0].user[1]);
  }
  {  // BASELINE config 2 through Builder: 4096 seeds, all pass, a leader everywhere
    Builder b;
    b.count = 4096;
    auto out = b.run(Workload::raft());
    EXPECT(out.stats.n_pass == 4096 && out.stats.n_fail == 0);
    bool leaders = true;
    for (auto& r : out.results) leaders = leaders && r.user[0] >= 1 && r.user[0] <= 5;
    EXPECT(leaders);
    std::printf("ok   raft x4096: %llu events in %.2f ms\n", (unsigned long long)(out.stats.ev_pop + out.stats.ev_fire),
                out.stats.kernel_ms);
  }
  {  // seeds-to-first-bug with the injected double-vote bug: the FIRST failing seed is reported
    Builder b;
    b.count = 512;
    uint64_t first = ~0ull;
    try {
      b.run(Workload::raft(5, 10000, 0, 1));
    } catch (SimulationFailure& e) {
      first = e.seed;
      EXPECT(e.status == MSIM_ST_FAIL_ASSERT && e.fail_code == 1);
      bool again = false;
      try {  // MADSIM_TEST_SEED=<first> reproduces it
        Builder r;
        r.seed = first;
        r.run(Workload::raft(5, 10000, 0, 1));
      } catch (SimulationFailure& e2) {
        again = e2.seed == first;
      }
      EXPECT(again);
    }
    EXPECT(first == 88);  // pinned by the oracle (tests/golden: raft_cfg2_bug)
    std::printf("ok   seeds-to-first-bug = %llu\n", (unsigned long long)first);
  }
  std::printf("%s (%d failed)\n", failed ? "HOST API TEST FAILED" : "HOST API TEST PASSED", failed);
  return failed ? 1 : 0;
}
