"""madsim_b200 — B200-native many-seed executor for madsim-style deterministic simulation.

Only the hot path of the reference (the per-seed executor step) lives here: CUDA kernels behind
a C-ABI (csrc/, include/msim_b200.h) plus this thin host mirror of the reference's runtime
interface.  Importing the package does not load CUDA; creating a Batch does, and fails loudly
when the extension or the device is missing.
"""
from .api import (  # noqa: F401
    MS,
    SEC,
    SEED_RESULT_DTYPE,
    TRACE_DTYPE,
    Batch,
    Builder,
    Config,
    NonDeterminism,
    SimulationFailure,
    check_determinism,
    Workload,
    device_count,
    state_bytes,
)
from ._native import (  # noqa: F401
    EV_NAMES,
    GEOMETRY_LOCKSTEP,
    RAFT_FAULT_CLOCK_JUMP,
    RAFT_FAULT_CLOG_LINK,
    RAFT_FAULT_CLOG_NODE,
    ST_FAIL_ASSERT,
    ST_FAIL_CAPACITY,
    ST_FAIL_DEADLOCK,
    ST_FAIL_TIME_LIMIT,
    ST_PASS,
    ST_RUNNING,
    STATUS_NAMES,
    WL_CHAOS,
    WL_ELECTION,
    WL_PINGPONG,
    WL_RAFT,
    WL_SLEEP,
    WL_STREAM,
    WL_ETCD,
    WL_TONIC,
    WL_SIGNAL,
    MsimError,
)

__version__ = "0.1.0"
