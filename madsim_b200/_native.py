"""ctypes binding of libmsim_b200.so (the C-ABI of include/msim_b200.h).

The library is CUDA-only.  If it has not been built, or no CUDA device is present, the
errors below are raised — there is deliberately no CPU fallback (the CPU restatement under
oracle/ is test infrastructure and is never imported from this package).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSIM_B200_LIB") or os.path.join(_HERE, "libmsim_b200.so")  # env override: kernel experiments

OK, EINVAL, ENODEV, ECUDA, ENOMEM, ESTATE, ERANGE = 0, -1, -2, -3, -4, -5, -6
ST_RUNNING, ST_PASS, ST_FAIL_ASSERT, ST_FAIL_DEADLOCK, ST_FAIL_TIME_LIMIT, ST_FAIL_CAPACITY = range(6)
WL_SLEEP, WL_PINGPONG, WL_RAFT, WL_CHAOS, WL_STREAM, WL_ETCD, WL_TONIC, WL_SIGNAL, WL_ELECTION = range(9)
EV_NAMES = {1: "POLL", 2: "DROP", 3: "PARK", 4: "SKIP", 5: "WAKE", 6: "DELIVER", 7: "RESTART"}
STATUS_NAMES = {0: "RUNNING", 1: "PASS", 2: "FAIL_ASSERT", 3: "FAIL_DEADLOCK", 4: "FAIL_TIME_LIMIT", 5: "FAIL_CAPACITY"}
RAFT_FAULT_CLOG_NODE, RAFT_FAULT_CLOG_LINK, RAFT_FAULT_CLOCK_JUMP = 1, 2, 4
GEOMETRY_LOCKSTEP = 0x100


class MsimConfig(C.Structure):
    _fields_ = [
        ("packet_loss_rate", C.c_double),
        ("send_latency_lo_ns", C.c_uint64),
        ("send_latency_hi_ns", C.c_uint64),
        ("time_limit_ns", C.c_uint64),
        ("buggify", C.c_uint32),
        ("reserved0", C.c_uint32),
    ]


class MsimWorkloadDesc(C.Structure):
    _fields_ = [("workload", C.c_uint32), ("n_nodes", C.c_uint32), ("p", C.c_uint64 * 8)]


class MsimSeedResult(C.Structure):
    _fields_ = [
        ("trace_hash", C.c_uint64),
        ("elapsed_ns", C.c_uint64),
        ("draws", C.c_uint64),
        ("ev_pop", C.c_uint32),
        ("ev_fire", C.c_uint32),
        ("status", C.c_uint32),
        ("fail_code", C.c_uint32),
        ("user", C.c_uint32 * 4),
        ("peak_timers", C.c_uint32),
        ("peak_tasks", C.c_uint32),
    ]


class MsimTraceRec(C.Structure):
    _fields_ = [("now_ns", C.c_uint64), ("aux", C.c_uint64), ("kind_id", C.c_uint32), ("draws", C.c_uint32)]


class MsimStats(C.Structure):
    _fields_ = [
        ("n_seeds", C.c_uint64),
        ("ev_pop", C.c_uint64),
        ("ev_fire", C.c_uint64),
        ("draws", C.c_uint64),
        ("n_pass", C.c_uint64),
        ("n_fail", C.c_uint64),
        ("n_unfinished", C.c_uint64),
        ("first_fail_seed", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("launches", C.c_uint32),
        ("seeds_per_cta", C.c_uint32),
        ("lanes_per_warp", C.c_uint32),
        ("ctas", C.c_uint32),
        ("smem_bytes", C.c_uint32),
        ("escalated", C.c_uint32),
    ]


assert C.sizeof(MsimSeedResult) == 64 and C.sizeof(MsimTraceRec) == 24

# name -> (restype, argtypes): exactly the symbols include/msim_b200.h declares
PROTOTYPES = {
    "msim_version": (C.c_int, []),
    "msim_last_error": (C.c_char_p, []),
    "msim_device_count": (C.c_int, []),
    "msim_config_default": (C.c_int, [C.POINTER(MsimConfig)]),
    "msim_config_from_toml": (C.c_int, [C.c_char_p, C.POINTER(MsimConfig)]),
    "msim_batch_create": (C.c_int, [C.POINTER(MsimConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "msim_batch_destroy": (None, [C.c_void_p]),
    "msim_workload_default": (C.c_int, [C.c_uint32, C.POINTER(MsimWorkloadDesc)]),
    "msim_workload_load": (C.c_int, [C.c_void_p, C.POINTER(MsimWorkloadDesc)]),
    "msim_batch_set_geometry": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "msim_batch_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msim_batch_set_capacity_escalation": (C.c_int, [C.c_void_p, C.c_int]),
    "msim_batch_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "msim_batch_resume": (C.c_int, [C.c_void_p, C.c_uint64]),
    "msim_batch_run_async": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "msim_batch_sync": (C.c_int, [C.c_void_p]),
    "msim_batch_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MsimStats)]),
    "msim_batch_trace_hash": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msim_batch_summarize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "msim_batch_summarize_sharded": (
        C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]),
    "msim_workload_state_bytes": (C.c_int, [C.c_uint32]),
    "msim_batch_results_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "msim_batch_trace": (
        C.c_int,
        [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(MsimSeedResult)],
    ),
    "msim_batch_determinism_log": (
        C.c_int,
        [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(MsimSeedResult)],
    ),
}


class MsimError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"msim error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load libmsim_b200.so (built by `__graft_entry__.build()` / `make -C madsim_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "madsim_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise MsimError(rc, load().msim_last_error().decode("utf-8", "replace"))
    return rc
