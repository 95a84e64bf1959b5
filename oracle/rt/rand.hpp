// ORACLE — TEST INFRASTRUCTURE ONLY (see future.hpp).
//
// GlobalRng: restatement of madsim/src/sim/rand.rs:28-159 plus the arithmetic of the
// un-vendored crates it calls (madsim/Cargo.toml:28,44): rand_xoshiro 0.6
// (Xoshiro256PlusPlus, seed_from_u64 via SplitMix64) and rand 0.8 (UniformInt::sample_single,
// Uniform::new_inclusive/sample, UniformDuration, Bernoulli, Standard).  Published
// algorithms restated from the crates' documentation; pinned by the KATs in
// tests/test_oracle_kats.py.
#pragma once
#include <vector>

#include "future.hpp"

namespace msim_oracle {

/// SplitMix64 (rand_core::SeedableRng::seed_from_u64 as overridden by rand_xoshiro).
struct SplitMix64 {
  u64 x;
  u64 next() {
    x += 0x9e3779b97f4a7c15ull;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
};

struct Xoshiro256PlusPlus {
  u64 s[4];
  static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
  static Xoshiro256PlusPlus seed_from_u64(u64 seed) {
    SplitMix64 sm{seed};
    Xoshiro256PlusPlus r;
    for (auto& w : r.s) w = sm.next();
    return r;
  }
  u64 next_u64() {
    u64 result = rotl(s[0] + s[3], 23) + s[0];
    u64 t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
  }
  u32 next_u32() { return u32(next_u64() >> 32); }  // rand_xoshiro: upper bits
};

inline void wmul64(u64 a, u64 b, u64& hi, u64& lo) {
  unsigned __int128 p = (unsigned __int128)a * b;
  hi = u64(p >> 64);
  lo = u64(p);
}

/// rand.rs:28-38 `GlobalRng { inner: Arc<Mutex<Inner>> }` — one per Runtime.
class GlobalRng {
 public:
  explicit GlobalRng(u64 seed) : seed_(seed), rng_(Xoshiro256PlusPlus::seed_from_u64(seed)) {}

  u64 seed() const { return seed_; }
  u64 draws() const { return draws_; }
  const Xoshiro256PlusPlus& raw() const { return rng_; }

  // --- rand.rs:64-89 `with`: every draw goes through here; the determinism log/check
  // hashes (next byte of a CLONE of the rng) ^ xor-fold(elapsed nanos) per call.
  std::function<u64()> elapsed_ns_hook;  // TimeHandle::try_current().map(|t| t.elapsed())
  template <class Fn>
  auto with(Fn&& f) {
    auto ret = f(*this);
    if (log_ || check_) {
      unsigned __int128 t = elapsed_ns_hook ? (unsigned __int128)elapsed_ns_hook() : 0;
      u8 th = 0;
      for (int i = 0; i < 16; i++) th ^= u8(t >> (8 * i));
      Xoshiro256PlusPlus c = rng_;
      u8 v = u8(c.next_u32()) ^ th;  // Standard for u8: next_u32() as u8
      if (log_) log_->push_back(v);
      if (check_) {
        if (check_pos_ >= check_->size() || (*check_)[check_pos_] != v) panic("non-determinism detected");
        check_pos_++;
      }
    }
    return ret;
  }
  void enable_log() { log_.emplace(); }
  void enable_check(std::vector<u8> log) {
    check_ = std::move(log);
    check_pos_ = 0;
  }
  std::optional<std::vector<u8>> take_log() {
    if (log_) {
      auto l = std::move(log_);
      log_.reset();
      return l;
    }
    if (check_) {
      auto l = std::move(check_);
      check_.reset();
      return l;
    }
    return std::nullopt;
  }

  // --- buggify (rand.rs:109-134)
  void enable_buggify() { buggify_ = true; }
  void disable_buggify() { buggify_ = false; }
  bool is_buggify_enabled() const { return buggify_; }
  bool buggify() { return buggify_ && gen_bool(0.25); }
  bool buggify_with_prob(double p) { return buggify_ && gen_bool(p); }

  // --- raw generator (RngCore impl, rand.rs:142-159); each call is one `with`.
  u64 next_u64() {
    return with([](GlobalRng& g) { return g.step64(); });
  }
  u32 next_u32() {
    return with([](GlobalRng& g) { return u32(g.step64() >> 32); });
  }

  // --- rand 0.8 `Rng::gen_range(lo..hi)` for u64 / usize: UniformInt::sample_single
  u64 gen_range_u64(u64 lo, u64 hi) {
    return with([&](GlobalRng& g) { return g.sample_single_u64(lo, hi); });
  }
  u32 gen_range_u32(u32 lo, u32 hi) {
    return with([&](GlobalRng& g) { return g.sample_single_u32(lo, hi); });
  }
  /// gen_range(Duration..Duration): UniformDuration::new(lo, hi).sample (durations in ns).
  u64 gen_range_duration(u64 lo_ns, u64 hi_ns) {
    return with([&](GlobalRng& g) { return g.sample_duration(lo_ns, hi_ns); });
  }
  /// gen_bool(p): Bernoulli::new(p).sample
  bool gen_bool(double p) {
    return with([&](GlobalRng& g) { return g.sample_bool(p); });
  }
  /// Standard for u64 (`random::<u64>()`, rpc.rs:121)
  u64 gen_u64() { return next_u64(); }

  // ---- un-hooked arithmetic (exposed for KAT tests) ----
  u64 step64() {
    draws_++;
    return rng_.next_u64();
  }
  u64 sample_single_u64(u64 lo, u64 hi) {
    if (!(lo < hi)) panic("cannot sample empty range");
    // sample_single_inclusive(lo, hi - 1)
    u64 range = (hi - 1) - lo + 1;  // wrapping
    if (range == 0) return step64();
    u64 zone = (range << __builtin_clzll(range)) - 1;  // wrapping_sub
    for (;;) {
      u64 v = step64(), h, l;
      wmul64(v, range, h, l);
      if (l <= zone) return lo + h;
    }
  }
  u32 sample_single_u32(u32 lo, u32 hi) {
    if (!(lo < hi)) panic("cannot sample empty range");
    u32 range = (hi - 1) - lo + 1;
    if (range == 0) return u32(step64() >> 32);
    u32 zone = (range << __builtin_clz(range)) - 1;
    for (;;) {
      u32 v = u32(step64() >> 32);
      u64 m = u64(v) * range;
      if (u32(m) <= zone) return lo + u32(m >> 32);
    }
  }
  /// Uniform::<u64>::new_inclusive(lo, hi).sample
  u64 sample_inclusive_u64(u64 lo, u64 hi) {
    u64 range = hi - lo + 1;  // wrapping
    if (range == 0) return step64();
    u64 ints_to_reject = (UINT64_MAX - range + 1) % range;
    u64 zone = UINT64_MAX - ints_to_reject;
    for (;;) {
      u64 v = step64(), h, l;
      wmul64(v, range, h, l);
      if (l <= zone) return lo + h;
    }
  }
  u32 sample_inclusive_u32(u32 lo, u32 hi) {
    u32 range = hi - lo + 1;
    if (range == 0) return u32(step64() >> 32);
    u32 ints_to_reject = (UINT32_MAX - range + 1) % range;
    u32 zone = UINT32_MAX - ints_to_reject;
    for (;;) {
      u32 v = u32(step64() >> 32);
      u64 m = u64(v) * range;
      if (u32(m) <= zone) return lo + u32(m >> 32);
    }
  }
  /// UniformDuration (rand 0.8 distributions/uniform.rs): new(lo, hi) = new_inclusive(lo, hi - 1ns)
  u64 sample_duration(u64 lo_ns, u64 hi_ns) {
    if (!(lo_ns < hi_ns)) panic("Uniform::new called with `low >= high`");
    const u64 NS = 1000000000ull;
    u64 hi = hi_ns - 1;
    u64 low_s = lo_ns / NS, high_s = hi / NS;
    u32 low_n = u32(lo_ns % NS), high_n = u32(hi % NS);
    if (high_n < low_n) {
      high_s -= 1;
      high_n += u32(NS);
    }
    if (low_s == high_s) {  // Small: secs fixed, Uniform<u32> over nanos (may exceed 1e9; Duration::new carries)
      u32 n = sample_inclusive_u32(low_n, high_n);
      return low_s * NS + n;
    }
    // Medium: total nanoseconds fit u64 (always true for a u64-ns Duration model)
    return sample_inclusive_u64(low_s * NS + low_n, high_s * NS + high_n);
  }
  bool sample_bool(double p) {
    if (!(p >= 0.0 && p < 1.0)) {
      if (p == 1.0) return true;  // ALWAYS_TRUE: no draw
      panic("p is outside range [0.0, 1.0]");
    }
    u64 p_int = u64(p * 18446744073709551616.0);  // SCALE = 2^64
    return step64() < p_int;
  }

 private:
  u64 seed_;
  Xoshiro256PlusPlus rng_;
  u64 draws_ = 0;
  bool buggify_ = false;
  std::optional<std::vector<u8>> log_, check_;
  size_t check_pos_ = 0;
};

}  // namespace msim_oracle
