//! Raw bindings of `include/msim_b200.h` (libmsim_b200.so).  One `extern "C"` item per exported function and
//! one `#[repr(C)]` struct per C struct, field for field — `tests/test_rust_shim.py` parses this file against the
//! header on every test run (the build image has no Rust toolchain to compile it).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const MSIM_VERSION: c_int = 0x0001_0000;

// ---- return codes ----
pub const MSIM_OK: c_int = 0;
pub const MSIM_EINVAL: c_int = -1;
pub const MSIM_ENODEV: c_int = -2;
pub const MSIM_ECUDA: c_int = -3;
pub const MSIM_ENOMEM: c_int = -4;
pub const MSIM_ESTATE: c_int = -5;
pub const MSIM_ERANGE: c_int = -6;

// ---- per-seed outcome ----
pub const MSIM_ST_RUNNING: u32 = 0;
pub const MSIM_ST_PASS: u32 = 1;
pub const MSIM_ST_FAIL_ASSERT: u32 = 2;
pub const MSIM_ST_FAIL_DEADLOCK: u32 = 3;
pub const MSIM_ST_FAIL_TIME_LIMIT: u32 = 4;
pub const MSIM_ST_FAIL_CAPACITY: u32 = 5;

// ---- registered programs ----
pub const MSIM_WL_SLEEP: u32 = 0;
pub const MSIM_WL_PINGPONG: u32 = 1;
pub const MSIM_WL_RAFT: u32 = 2;
pub const MSIM_WL_CHAOS: u32 = 3;
pub const MSIM_WL_STREAM: u32 = 4;
pub const MSIM_WL_ETCD: u32 = 5;
pub const MSIM_WL_TONIC: u32 = 6;
pub const MSIM_WL_SIGNAL: u32 = 7;
pub const MSIM_WL_ELECTION: u32 = 8;

/// `struct msim_config` — madsim::Config{ net{ packet_loss_rate, send_latency } } + Runtime knobs
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct msim_config {
    pub packet_loss_rate: f64,
    pub send_latency_lo_ns: u64,
    pub send_latency_hi_ns: u64,
    pub time_limit_ns: u64,
    pub buggify: u32,
    pub reserved0: u32,
}

/// `struct msim_workload_desc` — what `f` is in `Builder::run(f)`
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct msim_workload_desc {
    pub workload: u32,
    pub n_nodes: u32,
    pub p: [u64; 8],
}

/// `struct msim_seed_result` (64 bytes)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct msim_seed_result {
    pub trace_hash: u64,
    pub elapsed_ns: u64,
    pub draws: u64,
    pub ev_pop: u32,
    pub ev_fire: u32,
    pub status: u32,
    pub fail_code: u32,
    pub user: [u32; 4],
    pub peak_timers: u32,
    pub peak_tasks: u32,
}

/// `struct msim_trace_rec` (24 bytes)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct msim_trace_rec {
    pub now_ns: u64,
    pub aux: u64,
    pub kind_id: u32,
    pub draws: u32,
}

/// `struct msim_stats`
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct msim_stats {
    pub n_seeds: u64,
    pub ev_pop: u64,
    pub ev_fire: u64,
    pub draws: u64,
    pub n_pass: u64,
    pub n_fail: u64,
    pub n_unfinished: u64,
    pub first_fail_seed: u64,
    pub kernel_ms: f64,
    pub launches: u32,
    pub seeds_per_cta: u32,
    pub lanes_per_warp: u32,
    pub ctas: u32,
    pub smem_bytes: u32,
    pub escalated: u32,
}

/// opaque `struct msim_batch`
#[repr(C)]
pub struct msim_batch {
    _private: [u8; 0],
}

extern "C" {
    pub fn msim_version() -> c_int;
    pub fn msim_last_error() -> *const c_char;
    pub fn msim_device_count() -> c_int;
    pub fn msim_config_default(cfg: *mut msim_config) -> c_int;
    pub fn msim_config_from_toml(text: *const c_char, cfg: *mut msim_config) -> c_int;
    pub fn msim_batch_create(cfg: *const msim_config, device: c_int, out: *mut *mut msim_batch) -> c_int;
    pub fn msim_batch_destroy(b: *mut msim_batch);
    pub fn msim_workload_default(workload: u32, out: *mut msim_workload_desc) -> c_int;
    pub fn msim_workload_load(b: *mut msim_batch, w: *const msim_workload_desc) -> c_int;
    pub fn msim_batch_set_geometry(b: *mut msim_batch, seeds_per_cta: u32, lanes_per_warp: u32) -> c_int;
    pub fn msim_batch_set_stream(b: *mut msim_batch, cuda_stream: *mut c_void) -> c_int;
    pub fn msim_batch_set_capacity_escalation(b: *mut msim_batch, on: c_int) -> c_int;
    pub fn msim_batch_run(b: *mut msim_batch, seed_base: u64, n_seeds: u64, max_events_per_seed: u64) -> c_int;
    pub fn msim_batch_resume(b: *mut msim_batch, max_events_per_seed: u64) -> c_int;
    pub fn msim_batch_run_async(b: *mut msim_batch, seed_base: u64, n_seeds: u64, max_events_per_seed: u64) -> c_int;
    pub fn msim_batch_sync(b: *mut msim_batch) -> c_int;
    pub fn msim_batch_results(b: *mut msim_batch, out: *mut msim_seed_result, fail_bitmap: *mut u32, stats: *mut msim_stats) -> c_int;
    pub fn msim_batch_trace_hash(b: *mut msim_batch, hash: *mut u64) -> c_int;
    pub fn msim_batch_summarize_device(b: *mut msim_batch, d_out: *mut c_void, bytes: usize) -> c_int;
    pub fn msim_batch_summarize_sharded(b: *mut msim_batch, d_out: *mut c_void, bytes: usize, rank: u32, world: u32, seed_base: u64, n_total: u64) -> c_int;
    pub fn msim_workload_state_bytes(workload: u32) -> c_int;
    pub fn msim_batch_results_device(b: *mut msim_batch, dptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
    pub fn msim_batch_trace(b: *mut msim_batch, seed: u64, out: *mut msim_trace_rec, cap: usize, n: *mut usize, result: *mut msim_seed_result) -> c_int;
    pub fn msim_batch_determinism_log(b: *mut msim_batch, seed: u64, out: *mut u8, cap: usize, n: *mut usize, result: *mut msim_seed_result) -> c_int;
}
