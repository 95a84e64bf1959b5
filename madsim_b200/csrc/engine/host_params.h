// Host-side digest of (msim_config, msim_workload_desc) → engine Params.
// Mirrors how rand 0.8 builds its samplers so the device only does integer work:
//   Bernoulli::new(p)            p_int = (p * 2^64) as u64, p == 1.0 → always true
//   UniformDuration::new(lo,hi)  Small / Medium mode, Uniform::new_inclusive zone
// (reference call sites network.rs:262,267).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "defs.h"

namespace msim {

inline int workload_defaults(uint32_t workload, msim_workload_desc* d) {
  std::memset(d, 0, sizeof(*d));
  d->workload = workload;
  switch (workload) {
    case MSIM_WL_SLEEP:
      d->p[0] = NS_PER_SEC;
      return 0;
    case MSIM_WL_PINGPONG:
      d->n_nodes = 2;
      d->p[0] = 1000;
      d->p[1] = 10 * NS_PER_MS;
      return 0;
    case MSIM_WL_RAFT:
      d->n_nodes = 5;
      d->p[0] = 10000;
      d->p[3] = 40;
      d->p[4] = 50;
      d->p[5] = 150;
      d->p[6] = 300;
      return 0;
    case MSIM_WL_STREAM:
      d->n_nodes = 2;
      d->p[0] = 200;
      d->p[1] = 5 * NS_PER_MS;
      d->p[2] = 1;
      return 0;
    case MSIM_WL_ETCD:
      d->n_nodes = 3;
      d->p[0] = 32;
      d->p[1] = 50 * NS_PER_MS;
      d->p[4] = 1;
      return 0;
    case MSIM_WL_TONIC:
      d->n_nodes = 7;
      d->p[0] = 10;
      d->p[1] = 100 * NS_PER_MS;
      return 0;
    case MSIM_WL_SIGNAL:
      d->n_nodes = 1;
      d->p[0] = 8;
      d->p[1] = 100;
      d->p[3] = 1;
      return 0;
    case MSIM_WL_CHAOS:
      d->n_nodes = 1;
      d->p[0] = 8;
      d->p[1] = 2000;
      return 0;
    case MSIM_WL_ELECTION:
      d->n_nodes = 4;
      d->p[0] = 30;
      return 0;
  }
  return MSIM_EINVAL;
}

inline void config_defaults(msim_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->packet_loss_rate = 0.0;
  c->send_latency_lo_ns = 1 * NS_PER_MS;
  c->send_latency_hi_ns = 10 * NS_PER_MS;
}

/// `impl FromStr for Config` (config.rs:28-35): the TOML a MADSIM_TEST_CONFIG file holds (builder.rs:82-88) —
///   [net]  packet_loss_rate = 0.1
///          send_latency = { start = { secs = 0, nanos = 1000000 }, end = { secs = 0, nanos = 10000000 } }
///   [tcp]
/// in the inline form of the reference's own test (config.rs:52-61) or the dotted-table form its Display impl
/// prints ([net.send_latency.start] ...).  A small TOML subset: tables, dotted keys, inline tables, integers,
/// floats, booleans and strings; unknown keys are ignored (serde's default), missing ones take the serde defaults
/// (packet_loss_rate 0.0, send_latency 1 ms..10 ms).  Runtime knobs (time limit, buggify) are not part of the file.
namespace toml_detail {
struct Parser {
  const char* s;
  std::string* err;
  std::map<std::string, double> num;  // dotted path -> numeric value
  std::map<std::string, bool> seen;   // dotted path -> present (any type)
  bool fail(const std::string& m) {
    if (err && err->empty()) *err = "config: " + m;
    return false;
  }
  void mark(const std::string& path) {  // the path and every table above it exist
    for (size_t i = 0; i <= path.size(); i++)
      if (i == path.size() || path[i] == '.') seen[path.substr(0, i)] = true;
  }
  void ws(bool newlines) {
    for (;;) {
      while (*s == ' ' || *s == '\t' || *s == '\r' || (newlines && *s == '\n')) s++;
      if (*s == '#') {
        while (*s && *s != '\n') s++;
        if (newlines) continue;
      }
      break;
    }
  }
  bool key(std::string* out) {  // bare or quoted, dotted
    out->clear();
    for (;;) {
      ws(false);
      std::string part;
      if (*s == '"' || *s == '\'') {
        char q = *s++;
        while (*s && *s != q) part += *s++;
        if (*s != q) return fail("unterminated key");
        s++;
      } else {
        while ((*s >= 'a' && *s <= 'z') || (*s >= 'A' && *s <= 'Z') || (*s >= '0' && *s <= '9') || *s == '_' || *s == '-') part += *s++;
      }
      if (part.empty()) return fail("expected a key");
      if (!out->empty()) *out += '.';
      *out += part;
      ws(false);
      if (*s != '.') return true;
      s++;
    }
  }
  bool value(const std::string& path) {
    ws(false);
    mark(path);
    if (*s == '{') {  // inline table
      s++;
      ws(true);
      if (*s == '}') return s++, true;
      for (;;) {
        std::string k;
        if (!key(&k)) return false;
        if (*s != '=') return fail("expected '=' in inline table");
        s++;
        if (!value(path + "." + k)) return false;
        ws(true);
        if (*s == ',') {
          s++;
          ws(true);
          continue;
        }
        if (*s == '}') return s++, true;
        return fail("expected ',' or '}' in inline table");
      }
    }
    if (*s == '"' || *s == '\'') {
      char q = *s++;
      while (*s && *s != q) s += (*s == '\\' && q == '"' && s[1]) ? 2 : 1;
      if (*s != q) return fail("unterminated string");
      s++;
      return true;
    }
    if (!std::strncmp(s, "true", 4) || !std::strncmp(s, "false", 5)) {
      num[path] = *s == 't';
      s += *s == 't' ? 4 : 5;
      return true;
    }
    if (*s == '[') return fail("arrays are not part of madsim::Config");
    std::string lit;
    while (*s && !std::strchr(" \t\r\n,}#", *s)) {
      if (*s != '_') lit += *s;
      s++;
    }
    if (lit == "inf" || lit == "+inf") return num[path] = INFINITY, true;
    if (lit == "nan" || lit == "+nan" || lit == "-nan") return num[path] = NAN, true;
    char* end = nullptr;
    double v = std::strtod(lit.c_str(), &end);
    if (lit.empty() || *end) return fail("bad value for " + path);
    num[path] = v;
    return true;
  }
  bool document() {
    std::string table;
    for (;;) {
      ws(true);
      if (!*s) return true;
      if (*s == '[') {
        if (s[1] == '[') return fail("arrays of tables are not part of madsim::Config");
        s++;
        if (!key(&table)) return false;
        if (*s != ']') return fail("expected ']'");
        s++;
        mark(table);
      } else {
        std::string k;
        if (!key(&k)) return false;
        if (*s != '=') return fail("expected '='");
        s++;
        if (!value(table.empty() ? k : table + "." + k)) return false;
      }
      ws(false);
      if (*s && *s != '\n') return fail("expected end of line");
    }
  }
};
}  // namespace toml_detail

inline int config_from_toml(const char* text, msim_config* out, std::string* err) {
  toml_detail::Parser p{text, err, {}, {}};
  if (err) err->clear();
  if (!p.document()) return MSIM_EINVAL;
  config_defaults(out);
  if (p.num.count("net.packet_loss_rate")) out->packet_loss_rate = p.num["net.packet_loss_rate"];
  if (p.seen.count("net.send_latency")) {
    u64 v[2];
    const char* ends[2] = {"start", "end"};
    for (int i = 0; i < 2; i++) {
      std::string base = std::string("net.send_latency.") + ends[i];
      if (!p.num.count(base + ".secs") || !p.num.count(base + ".nanos")) {
        if (err) *err = "config: send_latency needs " + base + " = { secs, nanos }";  // serde: missing field
        return MSIM_EINVAL;
      }
      double secs = p.num[base + ".secs"], nanos = p.num[base + ".nanos"];
      if (!(secs >= 0 && secs < 1.8e10) || !(nanos >= 0 && nanos < 4294967296.0) || secs != std::floor(secs) || nanos != std::floor(nanos)) {
        if (err) *err = "config: " + base + " out of range";
        return MSIM_EINVAL;
      }
      v[i] = (u64)secs * NS_PER_SEC + (u64)nanos;  // Duration::new carries nanos >= 1e9 into secs
    }
    out->send_latency_lo_ns = v[0];
    out->send_latency_hi_ns = v[1];
  }
  return 0;
}

/// Returns 0 or MSIM_EINVAL (+ message).
inline int make_params(const msim_config& c, const msim_workload_desc& wl, Params* out, std::string* err) {
  Params P;
  std::memset(&P, 0, sizeof(P));
  P.workload = wl.workload;
  P.n_nodes = wl.n_nodes;
  for (int i = 0; i < 8; i++) P.p[i] = wl.p[i];
  double p = c.packet_loss_rate;
  if (!(p >= 0.0 && p <= 1.0)) {
    if (err) *err = "packet_loss_rate outside [0, 1]";
    return MSIM_EINVAL;
  }
  if (p == 1.0) {
    P.loss_always = 1;
  } else {
    P.loss_p_int = (u64)(p * 18446744073709551616.0);
  }
  u64 lo = c.send_latency_lo_ns, hi = c.send_latency_hi_ns;
  if (!(lo < hi)) {
    if (err) *err = "send_latency: low >= high";
    return MSIM_EINVAL;
  }
  u64 hi_incl = hi - 1;
  u64 low_s = lo / NS_PER_SEC, high_s = hi_incl / NS_PER_SEC;
  u64 low_n = lo % NS_PER_SEC, high_n = hi_incl % NS_PER_SEC;
  if (high_n < low_n) {
    high_s -= 1;
    high_n += NS_PER_SEC;
  }
  if (low_s == high_s) {
    P.lat_mode = 0;
    P.lat_base = low_s * NS_PER_SEC;
    P.lat_lo = low_n;
    u32 range = (u32)(high_n - low_n + 1);
    u32 reject = range ? (u32)((0xffffffffu - range + 1) % range) : 0;
    P.lat_range = range;
    P.lat_zone = 0xffffffffu - reject;
    if (range == 0) {
      if (err) *err = "send_latency: full u32 range unsupported";
      return MSIM_EINVAL;
    }
  } else {
    P.lat_mode = 1;
    u64 l = low_s * NS_PER_SEC + low_n, h = high_s * NS_PER_SEC + high_n;
    u64 range = h - l + 1;
    P.lat_lo = l;
    P.lat_range = range;
    P.lat_zone = ~0ull - ((~0ull - range + 1) % range);
  }
  P.time_limit_ns = c.time_limit_ns;
  P.time_limit_cmp = c.time_limit_ns ? c.time_limit_ns : ~0ull;
  P.buggify = c.buggify ? 1 : 0;
  P.buggify_p10_int = (u64)(0.1 * 18446744073709551616.0);
  // workload parameter validation
  switch (wl.workload) {
    case MSIM_WL_SLEEP:
      break;
    case MSIM_WL_PINGPONG:
      break;
    case MSIM_WL_RAFT:
      if (P.n_nodes == 0) P.n_nodes = 5;
      if (P.n_nodes > 5 || ((wl.p[1] & MSIM_RAFT_FAULT_CLOG_LINK) && P.n_nodes < 2)) {
        if (err) *err = "raft: n_nodes must be 1..5 (>= 2 with link faults)";
        return MSIM_EINVAL;
      }
      if (!(wl.p[5] < wl.p[6]) || wl.p[3] == 0 || wl.p[4] == 0) {
        if (err) *err = "raft: need el_lo < el_hi, rpc timeout > 0, heartbeat > 0";
        return MSIM_EINVAL;
      }
      if (wl.p[1] > 7) {
        if (err) *err = "raft: unknown fault mask bits";
        return MSIM_EINVAL;
      }
      break;
    case MSIM_WL_STREAM:
      if (wl.p[2] > 1) {
        if (err) *err = "stream: unknown fault bits";
        return MSIM_EINVAL;
      }
      break;
    case MSIM_WL_ETCD: {
      if (P.n_nodes == 0) P.n_nodes = 3;
      float rate;
      u32 bits = (u32)wl.p[2];
      std::memcpy(&rate, &bits, 4);
      if (P.n_nodes < 2 || P.n_nodes > 4 || wl.p[0] * (P.n_nodes - 1) > 128 || wl.p[3] > 3 || wl.p[4] > 0xffff || wl.p[7] > 1 ||
          !(rate >= 0.0f && rate <= 1.0f)) {
        if (err) *err = "etcd: need 1..3 clients, clients x ops <= 128, bug bits <= 3, 0 <= timeout_rate <= 1";
        return MSIM_EINVAL;
      }
      // Bernoulli::new(timeout_rate as f64) (service.rs:166)
      P.p[5] = rate == 1.0f ? 1 : 0;  // (p[7], the lease-admin switch, passes through)
      P.p[6] = rate == 1.0f ? 0 : (u64)((double)rate * 18446744073709551616.0);
      break;
    }
    case MSIM_WL_TONIC:
      if (P.n_nodes == 0) P.n_nodes = 7;
      if (P.n_nodes < 2 || P.n_nodes > 7 || wl.p[2] > 31 || ((wl.p[2] & 16) && (wl.p[2] & 10))) {
        if (err) *err = "tonic: need 1..6 client nodes, flags <= 31, flag 16 (interceptors) without 2 and 8";
        return MSIM_EINVAL;
      }
      break;
    case MSIM_WL_SIGNAL:
      if (wl.p[1] == 0 || wl.p[2] > 2 || wl.p[3] > 1) {
        if (err) *err = "signal: period must be > 0, behavior 0..2, flags <= 1";
        return MSIM_EINVAL;
      }
      break;
    case MSIM_WL_CHAOS:
      if (wl.p[1] == 0 || wl.p[3] > 10 || (wl.p[3] && wl.p[2] > 10)) {
        if (err) *err = "chaos: period must be > 0; matching mode (p3) takes at most 10 patterns and 10 panics";
        return MSIM_EINVAL;
      }
      break;
    case MSIM_WL_ELECTION: {
      if (P.n_nodes == 0) P.n_nodes = 4;
      float rate;
      u32 bits = (u32)wl.p[1];
      std::memcpy(&rate, &bits, 4);
      // the observer counts two keys about 20 s in, leases last 60 s (tests/test.rs:154-232)
      if (P.n_nodes != 4 || wl.p[0] < 20 || wl.p[0] > 50 || !(rate >= 0.0f && rate <= 1.0f)) {
        if (err) *err = "election: 4 nodes, 20 <= hold seconds <= 50, 0 <= timeout_rate <= 1";
        return MSIM_EINVAL;
      }
      P.p[5] = rate == 1.0f ? 1 : 0;  // Bernoulli::new(timeout_rate as f64) (service.rs:166)
      P.p[6] = rate == 1.0f ? 0 : (u64)((double)rate * 18446744073709551616.0);
      break;
    }
    default:
      if (err) *err = "unknown workload";
      return MSIM_EINVAL;
  }
  *out = P;
  return 0;
}

}  // namespace msim
