//! madsim-b200 — madsim's many-seed surface over libmsim_b200.so.
//!
//! The seam is `madsim::runtime::Builder::run` (reference `madsim/src/sim/runtime/builder.rs:121-162`): it runs the
//! test body once per seed, each on a fresh `Runtime`, on `MADSIM_TEST_JOBS` threads.  Here the same sweep is ONE
//! batched launch on a B200: every seed is a lane-private `Runtime` (own xoshiro256++ stream, clock, timer heap,
//! ready queue, network) advanced by the CUDA step kernel, bit-identical to the reference executor per seed.
//!
//! An arbitrary `async` test body cannot run on a GPU; a test selects one of the REGISTERED programs
//! ([`Workload`]) — state machines compiled into the library — instead of passing a closure:
//!
//! ```ignore
//! use madsim_b200::{runtime::Builder, Workload};
//! #[test]
//! fn raft_election() {
//!     // MADSIM_TEST_SEED / _NUM / _JOBS / _CONFIG / _TIME_LIMIT / _CHECK_DETERMINISM keep their meaning
//!     Builder::from_env().run(Workload::raft());
//! }
//! ```
//!
//! Names, defaults, environment variables and panic texts follow the reference so that a test reads the same.
//! SOURCE ONLY here (no cargo in the build image); `tests/test_rust_shim.py` checks `ffi.rs` against the header.

pub mod ffi;

use std::ffi::{CStr, CString};
use std::fmt;
use std::ops::Range;
use std::time::{Duration, SystemTime};

/// Error of a library call (not of a simulated seed): negative `MSIM_E*` code + `msim_last_error()`.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "msim error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

fn check(rc: i32) -> Result<(), Error> {
    if rc == ffi::MSIM_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(ffi::msim_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

/// `madsim::net::Config` (reference `net/network.rs:66-89`).
#[derive(Debug, Clone, PartialEq)]
pub struct NetConfig {
    /// Possibility of packet loss.
    pub packet_loss_rate: f64,
    /// The latency range of sending packets.
    pub send_latency: Range<Duration>,
}
impl Default for NetConfig {
    fn default() -> Self {
        NetConfig { packet_loss_rate: 0.0, send_latency: Duration::from_millis(1)..Duration::from_millis(10) }
    }
}

/// `madsim::Config` (reference `config.rs:11-21`).
#[derive(Debug, Clone, Default, PartialEq)]
pub struct Config {
    pub net: NetConfig,
}
impl Config {
    fn to_c(&self, time_limit: Option<Duration>, buggify: bool) -> ffi::msim_config {
        ffi::msim_config {
            packet_loss_rate: self.net.packet_loss_rate,
            send_latency_lo_ns: self.net.send_latency.start.as_nanos() as u64,
            send_latency_hi_ns: self.net.send_latency.end.as_nanos() as u64,
            time_limit_ns: time_limit.map(|d| d.as_nanos() as u64).unwrap_or(0),
            buggify: buggify as u32,
            reserved0: 0,
        }
    }
}
/// Parse a config from TOML (reference `config.rs:28-35`; the `[net]` table with `packet_loss_rate` and
/// `send_latency = { start = { secs, nanos }, end = { secs, nanos } }`; `[tcp]` is accepted and empty).
impl std::str::FromStr for Config {
    type Err = Error;
    fn from_str(s: &str) -> Result<Self, Self::Err> {
        let text = CString::new(s).map_err(|_| Error { code: ffi::MSIM_EINVAL, message: "NUL in config".into() })?;
        let mut c = ffi::msim_config {
            packet_loss_rate: 0.0,
            send_latency_lo_ns: 0,
            send_latency_hi_ns: 0,
            time_limit_ns: 0,
            buggify: 0,
            reserved0: 0,
        };
        check(unsafe { ffi::msim_config_from_toml(text.as_ptr(), &mut c) })?;
        Ok(Config {
            net: NetConfig {
                packet_loss_rate: c.packet_loss_rate,
                send_latency: Duration::from_nanos(c.send_latency_lo_ns)..Duration::from_nanos(c.send_latency_hi_ns),
            },
        })
    }
}

/// A registered device program: what `f` is in `Builder::run(f)`.  Parameter slots as in `include/msim_b200.h`.
#[derive(Debug, Clone, Copy, PartialEq)]
pub struct Workload(pub ffi::msim_workload_desc);
impl Workload {
    fn defaults(id: u32) -> ffi::msim_workload_desc {
        let mut d = ffi::msim_workload_desc::default();
        check(unsafe { ffi::msim_workload_default(id, &mut d) }).expect("unknown workload id");
        d
    }
    /// `block_on(sleep(d))`
    pub fn sleep(d: Duration) -> Self {
        let mut w = Self::defaults(ffi::MSIM_WL_SLEEP);
        w.p[0] = d.as_nanos() as u64;
        Workload(w)
    }
    /// 2-node ping-pong over `Endpoint::send_to / recv_from` + `time::sleep` (BASELINE config 1)
    pub fn pingpong(rounds: u64) -> Self {
        let mut w = Self::defaults(ffi::MSIM_WL_PINGPONG);
        w.p[0] = rounds;
        Workload(w)
    }
    /// 5-node Raft leader election over `call_timeout` / `add_rpc_handler`, 10 virtual seconds (BASELINE config 2)
    pub fn raft() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_RAFT))
    }
    /// the same with the supervisor's partition / clock-jump fault schedule (BASELINE config 3); `mask` = MSIM_RAFT_FAULT_*
    pub fn raft_with_faults(mask: u64) -> Self {
        let mut w = Self::defaults(ffi::MSIM_WL_RAFT);
        w.p[1] = mask;
        Workload(w)
    }
    /// madsim-etcd-client `SimServer` + KV / lease / election clients with a linearizability check (BASELINE config 4)
    pub fn etcd() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_ETCD))
    }
    /// tonic-example Greeter under madsim-tonic with kill / restart chaos of the client nodes (BASELINE config 5)
    pub fn tonic() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_TONIC))
    }
    /// madsim-etcd-client's election API (campaign / proclaim / leader / observe / resign) as its `election()` test drives it
    pub fn election() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_ELECTION))
    }
    pub fn chaos() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_CHAOS))
    }
    pub fn stream() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_STREAM))
    }
    pub fn signal() -> Self {
        Workload(Self::defaults(ffi::MSIM_WL_SIGNAL))
    }
    /// set parameter slot `i`
    pub fn param(mut self, i: usize, v: u64) -> Self {
        self.0.p[i] = v;
        self
    }
    pub fn nodes(mut self, n: u32) -> Self {
        self.0.n_nodes = n;
        self
    }
}

/// Outcome of one seed — what a normal return / a panic of `Runtime::block_on` is in the reference.
pub type SeedResult = ffi::msim_seed_result;

/// Many `Runtime`s that share a config, advanced together on one GPU (RAII over `msim_batch`).
pub struct Batch {
    raw: *mut ffi::msim_batch,
    n_seeds: u64,
}
// one host thread drives a batch at a time; the handle itself may move between threads
unsafe impl Send for Batch {}
impl Batch {
    pub fn new(config: &ffi::msim_config, device: i32) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::msim_batch_create(config, device, &mut raw) })?;
        Ok(Batch { raw, n_seeds: 0 })
    }
    pub fn load(&mut self, w: &Workload) -> Result<&mut Self, Error> {
        check(unsafe { ffi::msim_workload_load(self.raw, &w.0) })?;
        Ok(self)
    }
    /// seeds that overflow a bounded table re-run on the program's large-capacity variant (default on)
    pub fn set_capacity_escalation(&mut self, on: bool) -> Result<&mut Self, Error> {
        check(unsafe { ffi::msim_batch_set_capacity_escalation(self.raw, on as std::os::raw::c_int) })?;
        Ok(self)
    }
    /// seeds `seed_base .. seed_base + n`, each to completion
    pub fn run(&mut self, seed_base: u64, n: u64) -> Result<&mut Self, Error> {
        check(unsafe { ffi::msim_batch_run(self.raw, seed_base, n, 0) })?;
        self.n_seeds = n;
        Ok(self)
    }
    pub fn results(&mut self) -> Result<(Vec<SeedResult>, Vec<u32>, ffi::msim_stats), Error> {
        let mut res = vec![SeedResult::default(); self.n_seeds as usize];
        let mut bitmap = vec![0u32; ((self.n_seeds + 31) / 32) as usize];
        let mut stats = ffi::msim_stats::default();
        check(unsafe { ffi::msim_batch_results(self.raw, res.as_mut_ptr(), bitmap.as_mut_ptr(), &mut stats) })?;
        Ok((res, bitmap, stats))
    }
    /// full canonical trace of one seed (device replay) — the reference's debugging story is replay by seed
    pub fn trace(&mut self, seed: u64, cap: usize) -> Result<(Vec<ffi::msim_trace_rec>, SeedResult), Error> {
        let mut buf = vec![ffi::msim_trace_rec::default(); cap];
        let (mut n, mut r) = (0usize, SeedResult::default());
        check(unsafe { ffi::msim_batch_trace(self.raw, seed, buf.as_mut_ptr(), cap, &mut n, &mut r) })?;
        buf.truncate(n.min(cap));
        Ok((buf, r))
    }
    /// the per-draw determinism log of one seed (reference `rand.rs:64-89`): one byte per `GlobalRng::with` call
    pub fn determinism_log(&mut self, seed: u64) -> Result<(Vec<u8>, SeedResult), Error> {
        let (mut n, mut r) = (0usize, SeedResult::default());
        check(unsafe { ffi::msim_batch_determinism_log(self.raw, seed, std::ptr::null_mut(), 0, &mut n, &mut r) })?;
        let mut buf = vec![0u8; n];
        check(unsafe { ffi::msim_batch_determinism_log(self.raw, seed, buf.as_mut_ptr(), n, &mut n, &mut r) })?;
        Ok((buf, r))
    }
}
impl Drop for Batch {
    fn drop(&mut self) {
        unsafe { ffi::msim_batch_destroy(self.raw) }
    }
}

/// The reference's `panic_with_info` (runtime/mod.rs:205-210): the note, then the panic of the failing seed.
fn panic_with_info(seed: u64, r: &SeedResult) -> ! {
    eprintln!("note: run with `MADSIM_TEST_SEED={seed}` environment variable to reproduce this error");
    match r.status {
        ffi::MSIM_ST_FAIL_DEADLOCK => panic!("no events, all tasks will block forever"),
        ffi::MSIM_ST_FAIL_TIME_LIMIT => panic!("time limit exceeded"),
        ffi::MSIM_ST_FAIL_CAPACITY => panic!("msim: per-seed table {} overflowed (seed {seed})", r.fail_code),
        _ => panic!("seed {seed}: workload assertion {} failed", r.fail_code),
    }
}

pub mod runtime {
    //! `madsim::runtime::{Runtime, Builder}` (reference `runtime/mod.rs:34-210`, `runtime/builder.rs`).
    use super::*;

    /// One simulated world with a seed.
    pub struct Runtime {
        seed: u64,
        config: Config,
        time_limit: Option<Duration>,
        buggify: bool,
    }
    impl Runtime {
        /// `Runtime::new()`: seed 0, default config (runtime/mod.rs:42-45)
        pub fn new() -> Self {
            Self::with_seed_and_config(0, Config::default())
        }
        pub fn with_seed_and_config(seed: u64, config: Config) -> Self {
            Runtime { seed, config, time_limit: None, buggify: false }
        }
        /// `Runtime::set_time_limit` (runtime/mod.rs:133-135)
        pub fn set_time_limit(&mut self, limit: Duration) {
            self.time_limit = Some(limit);
        }
        /// `madsim::buggify::enable` before the run (rand.rs:119-127)
        pub fn enable_buggify(&mut self) {
            self.buggify = true;
        }
        /// `Runtime::block_on` for a registered program: returns the seed's outcome, panics as the reference would
        pub fn block_on(&self, w: Workload) -> SeedResult {
            let c = self.config.to_c(self.time_limit, self.buggify);
            let mut b = Batch::new(&c, 0).expect("msim_batch_create");
            b.load(&w).and_then(|b| b.run(self.seed, 1)).expect("msim_batch_run");
            let (res, _, _) = b.results().expect("msim_batch_results");
            if res[0].status != ffi::MSIM_ST_PASS {
                panic_with_info(self.seed, &res[0]);
            }
            res[0]
        }
        /// `Runtime::check_determinism` (runtime/mod.rs:178-202): run the seed twice, the second run checks every
        /// `GlobalRng::with` call against the first run's log of `(next rng byte) ^ fold(elapsed)`
        pub fn check_determinism(seed: u64, config: Config, w: Workload) -> SeedResult {
            let c = config.to_c(None, false);
            let mut b = Batch::new(&c, 0).expect("msim_batch_create");
            b.load(&w).expect("msim_workload_load");
            let (log, first) = b.determinism_log(seed).expect("msim_batch_determinism_log");
            if first.status != ffi::MSIM_ST_PASS {
                panic_with_info(seed, &first);
            }
            let (check, second) = b.determinism_log(seed).expect("msim_batch_determinism_log");
            if let Some(i) = (0..log.len().max(check.len())).find(|&i| log.get(i) != check.get(i)) {
                panic!("non-determinism detected at draw {i}");
            }
            second
        }
    }
    impl Default for Runtime {
        fn default() -> Self {
            Self::new()
        }
    }

    /// Builds the many-seed sweep from the environment, as `#[madsim::test]` does.
    pub struct Builder {
        /// The random seed for test.
        pub seed: u64,
        /// The number of tests.
        pub count: u64,
        /// The number of jobs to run simultaneously: here, the number of GPUs the sweep is sharded over.
        pub jobs: u16,
        /// The configuration.
        pub config: Config,
        /// The time limit for the test.
        pub time_limit: Option<Duration>,
        /// Enable determinism check.
        pub check: bool,
    }
    impl Builder {
        /// The same environment variables, defaults and panics as reference `builder.rs:64-118`.
        pub fn from_env() -> Self {
            let seed: u64 = if let Ok(s) = std::env::var("MADSIM_TEST_SEED") {
                s.parse().expect("MADSIM_TEST_SEED should be an integer")
            } else {
                SystemTime::now().duration_since(SystemTime::UNIX_EPOCH).unwrap().as_nanos() as _
            };
            let jobs: u16 = if let Ok(s) = std::env::var("MADSIM_TEST_JOBS") {
                s.parse().expect("MADSIM_TEST_JOBS should be an integer")
            } else {
                1
            };
            let config = if let Ok(path) = std::env::var("MADSIM_TEST_CONFIG") {
                let content = std::fs::read_to_string(path).expect("failed to read config file");
                content.parse::<Config>().expect("failed to parse config file")
            } else {
                Config::default()
            };
            let mut count: u64 = if let Ok(s) = std::env::var("MADSIM_TEST_NUM") {
                s.parse().expect("MADSIM_TEST_NUM should be an integer")
            } else {
                1
            };
            let time_limit = std::env::var("MADSIM_TEST_TIME_LIMIT").ok().map(|s| {
                Duration::from_secs_f64(s.parse::<f64>().expect("MADSIM_TEST_TIME_LIMIT should be an number"))
            });
            let check = std::env::var("MADSIM_TEST_CHECK_DETERMINISM").is_ok();
            if check {
                count = count.max(2);
            }
            Builder { seed, count, jobs, config, time_limit, check }
        }

        /// `Builder::run` (builder.rs:121-162): seeds `seed .. seed + count`; the first failing seed panics with the
        /// `MADSIM_TEST_SEED=` note; with `check`, `Runtime::check_determinism(seed)` instead.  `jobs` GPUs take
        /// contiguous seed ranges (one batch per device; the reference's jobs are threads).
        pub fn run(self, w: Workload) -> SeedResult {
            if self.check {
                return Runtime::check_determinism(self.seed, self.config, w);
            }
            let c = self.config.to_c(self.time_limit, false);
            let devices = (unsafe { ffi::msim_device_count() }).max(1) as u64;
            let jobs = (self.jobs as u64).clamp(1, devices).min(self.count.max(1));
            let (per, rem) = (self.count / jobs, self.count % jobs);
            let mut batches = Vec::new();
            let mut first = self.seed;
            for j in 0..jobs {
                let n = per + (j < rem) as u64;
                let mut b = Batch::new(&c, j as i32).expect("msim_batch_create");
                b.load(&w).expect("msim_workload_load");
                // enqueue on every device first, collect afterwards
                check(unsafe { ffi::msim_batch_run_async(b.raw, first, n, 0) }).expect("msim_batch_run_async");
                b.n_seeds = n;
                batches.push((first, b));
                first += n;
            }
            let mut last = None;
            for (first, mut b) in batches {
                let (res, _, stats) = b.results().expect("msim_batch_results");
                if stats.n_fail > 0 {
                    let s = stats.first_fail_seed;
                    panic_with_info(s, &res[(s - first) as usize]);
                }
                last = res.last().copied();
            }
            last.expect("count must be > 0")
        }
    }
}

#[cfg(feature = "madsim-compat")]
pub use runtime::{Builder, Runtime};
