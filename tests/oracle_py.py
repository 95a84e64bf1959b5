"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the CPU restatement of the
reference executor) and of tests/enginecheck (the engine's state machines compiled for the host).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
ENGINECHECK_LIB = os.path.join(ROOT, "tests", "enginecheck", "libenginecheck.so")

import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from madsim_b200 import _native as N  # struct layouts only (no CUDA is loaded by this import)
from madsim_b200.api import SEED_RESULT_DTYPE, TRACE_DTYPE, Config, Workload

_ora = None
_chk = None
ORACLE_BUILD = "x86-64-v2 (shipped: built in the build container for any x86-64 host)"


def use_native_build():
    """Rebuild the oracle with -march=native for THIS host (SURVEY.md §8d asks for it; the shipped library is
    -march=x86-64-v2 because it is built in another container) into oracle/_native/ and make it the library the
    next oracle() call loads.  Returns the description of the build in use; falls back to the shipped one."""
    global _ora, ORACLE_LIB, ORACLE_BUILD
    import shutil

    out = os.path.join(ORACLE_DIR, "_native", "liboracle.so")
    cxx = os.environ.get("CXX") or shutil.which("g++")
    try:
        if cxx is None:
            raise RuntimeError("no g++")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
        srcs += [os.path.join(ORACLE_DIR, "rt", f) for f in os.listdir(os.path.join(ORACLE_DIR, "rt"))]
        if not os.path.exists(out) or any(os.path.getmtime(x) > os.path.getmtime(out) for x in srcs):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run([cxx, "-std=c++20", "-O3", "-march=native", "-fPIC", "-pthread", "-shared", "-o", out,
                            os.path.join(ORACLE_DIR, "oracle_capi.cpp")], check=True, capture_output=True, timeout=300)
        lib = C.CDLL(out)
        lib.oracle_run_seed.restype = C.c_int
        lib.oracle_run_many.restype = C.c_int
        lib.oracle_check_determinism.restype = C.c_int
        _ora, ORACLE_LIB, ORACLE_BUILD = lib, out, "-O3 -march=native (rebuilt on this host)"
    except Exception as e:  # keep the shipped build
        ORACLE_BUILD = f"x86-64-v2 (shipped; native rebuild failed: {type(e).__name__})"
    return ORACLE_BUILD


def oracle():
    global _ora
    if _ora is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True)
        _ora = C.CDLL(ORACLE_LIB)
        _ora.oracle_run_seed.restype = C.c_int
        _ora.oracle_run_many.restype = C.c_int
        _ora.oracle_check_determinism.restype = C.c_int
    return _ora


def enginecheck():
    global _chk
    if _chk is None:
        if not os.path.exists(ENGINECHECK_LIB):
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", ENGINECHECK_LIB,
                            os.path.join(ROOT, "tests", "enginecheck", "enginecheck.cpp")], check=True)
        _chk = C.CDLL(ENGINECHECK_LIB)
        _chk.enginecheck_run.restype = C.c_int
        _chk.enginecheck_run_ml.restype = C.c_int
    return _chk


def run_seed(config: Config, wl: Workload, seed: int, trace: bool = False, cap: int = 1 << 20):
    """→ (result record, trace ndarray or None, message, extra[4])"""
    cfg, d = config.to_c(), wl.to_c()
    r = N.MsimSeedResult()
    n = C.c_size_t(0)
    buf = np.empty(cap, dtype=TRACE_DTYPE) if trace else None
    msg = C.create_string_buffer(256)
    extra = (C.c_uint32 * 6)()
    rc = oracle().oracle_run_seed(C.byref(cfg), C.byref(d), C.c_uint64(seed), C.byref(r),
                                  buf.ctypes.data_as(C.c_void_p) if trace else None, C.c_size_t(cap if trace else 0),
                                  C.byref(n), msg, C.c_size_t(256), extra)
    if rc != 0:
        raise RuntimeError(f"oracle_run_seed failed: {rc}")
    rec = np.frombuffer(bytes(r), dtype=SEED_RESULT_DTYPE)[0]
    return rec, (buf[: min(n.value, cap)].copy() if trace else None), msg.value.decode(), list(extra)


def run_many(config: Config, wl: Workload, seed_base: int, n: int, jobs: int = 0) -> np.ndarray:
    cfg, d = config.to_c(), wl.to_c()
    out = np.empty(n, dtype=SEED_RESULT_DTYPE)
    rc = oracle().oracle_run_many(C.byref(cfg), C.byref(d), C.c_uint64(seed_base), C.c_uint64(n), C.c_uint32(jobs),
                                  out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle_run_many failed: {rc}")
    return out


def check_determinism(config: Config, wl: Workload, seed: int) -> bool:
    cfg, d = config.to_c(), wl.to_c()
    return bool(oracle().oracle_check_determinism(C.byref(cfg), C.byref(d), C.c_uint64(seed)))


def engine_host(config: Config, wl: Workload, seed_base: int, n: int, seeds_per_block: int = 1, max_events: int = 0,
                trace_seed: int | None = None, cap: int = 1 << 20):
    """The device engine's state machines compiled with g++ (tests/enginecheck) — CPU-side
    development check only.  → (results ndarray, trace ndarray or None)"""
    cfg, d = config.to_c(), wl.to_c()
    out = np.empty(n, dtype=SEED_RESULT_DTYPE)
    buf = np.empty(cap, dtype=TRACE_DTYPE) if trace_seed is not None else None
    nt = C.c_size_t(0)
    rc = enginecheck().enginecheck_run(C.byref(cfg), C.byref(d), C.c_uint64(seed_base), C.c_uint64(n),
                                       C.c_uint32(seeds_per_block), C.c_uint64(max_events),
                                       out.ctypes.data_as(C.c_void_p),
                                       C.c_uint64(trace_seed if trace_seed is not None else 2**64 - 1),
                                       buf.ctypes.data_as(C.c_void_p) if buf is not None else None,
                                       C.c_size_t(cap), C.byref(nt))
    if rc != 0:
        raise RuntimeError(f"enginecheck_run failed: {rc}")
    return out, (buf[: min(nt.value, cap)].copy() if buf is not None else None)


def engine_host_big(config: Config, wl: Workload, seed_base: int, n: int, trace_seed: int | None = None, cap: int = 1 << 20):
    """Host build of the program's large-capacity variant (engine.h BigWL: what capacity escalation re-runs on).
    → results ndarray, or (results, trace of `trace_seed`) when a trace is asked for"""
    cfg, d = config.to_c(), wl.to_c()
    out = np.empty(n, dtype=SEED_RESULT_DTYPE)
    buf = np.empty(cap, dtype=TRACE_DTYPE) if trace_seed is not None else None
    nt = C.c_size_t(0)
    fn = enginecheck().enginecheck_run_big
    fn.restype = C.c_int
    rc = fn(C.byref(cfg), C.byref(d), C.c_uint64(seed_base), C.c_uint64(n), out.ctypes.data_as(C.c_void_p),
            C.c_uint64(trace_seed if trace_seed is not None else 2**64 - 1),
            buf.ctypes.data_as(C.c_void_p) if buf is not None else None, C.c_size_t(cap), C.byref(nt))
    if rc != 0:
        raise RuntimeError(f"enginecheck_run_big failed: {rc}")
    return out if trace_seed is None else (out, buf[: min(nt.value, cap)].copy())


def engine_host_ml(config: Config, wl: Workload, seed_base: int, n: int, lanes: int = 4) -> np.ndarray:
    """Host emulation of the multi-lane batch kernel's phase loop (tests/enginecheck run_wl_ml): `lanes`
    engines in lock step — CPU-side check that the phases execute exactly Engine::step()'s sequence."""
    cfg, d = config.to_c(), wl.to_c()
    out = np.empty(n, dtype=SEED_RESULT_DTYPE)
    rc = enginecheck().enginecheck_run_ml(C.byref(cfg), C.byref(d), C.c_uint64(seed_base), C.c_uint64(n),
                                          C.c_uint32(lanes), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"enginecheck_run_ml failed: {rc}")
    return out


def determinism_log(config: Config, wl: Workload, seed: int, impl: str = "oracle", cap: int = 1 << 22):
    """The per-`GlobalRng::with` byte log of Runtime::check_determinism (rand.rs:64-89) for one seed, from the oracle
    ("oracle") or from the host build of the device engine ("engine").  → (bytes ndarray, result record)"""
    cfg, d = config.to_c(), wl.to_c()
    buf = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    r = N.MsimSeedResult()
    fn = oracle().oracle_determinism_log if impl == "oracle" else enginecheck().enginecheck_determinism_log
    fn.restype = C.c_int
    rc = fn(C.byref(cfg), C.byref(d), C.c_uint64(seed), buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"determinism_log({impl}) failed: {rc}")
    assert n.value <= cap, "log capacity too small"
    return buf[: n.value].copy(), np.frombuffer(bytes(r), dtype=SEED_RESULT_DTYPE)[0]


def check_against_log(config: Config, wl: Workload, seed: int, log: np.ndarray) -> bool:
    """Second run of Runtime::check_determinism on the oracle against `log`: False = "non-determinism detected"."""
    cfg, d = config.to_c(), wl.to_c()
    fn = oracle().oracle_check_against_log
    fn.restype = C.c_int
    log = np.ascontiguousarray(log, dtype=np.uint8)
    return fn(C.byref(cfg), C.byref(d), C.c_uint64(seed), log.ctypes.data_as(C.c_void_p), C.c_size_t(len(log))) == 1


def first_trace_diff(a: np.ndarray, b: np.ndarray):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return None if len(a) == len(b) else n
