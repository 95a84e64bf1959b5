"""world_size-2 (and 3) gloo test of the N>1 path on CPU: contiguous seed shards, per-rank
summary, the all-reduce of counters / first failing seed / disjoint fail bitmap (madsim_b200/dist.py
— exactly the code the NCCL run uses; here the per-rank results come from the oracle because
there is no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed_base, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py as o

    import madsim_b200 as m
    from madsim_b200 import dist as D

    wl = m.Workload.raft(duration_ms=1500, bug=1, faults=7)
    first, count = D.shard(seed_base, n_total, rank, world)
    res = o.run_many(m.Config(), wl, first, count, jobs=2) if count else np.zeros(0, dtype=m.SEED_RESULT_DTYPE)
    c, bm = D.summary_from_results(res, first)
    counters = torch.from_numpy(c.view(np.int64).copy())
    bitmap = torch.from_numpy(bm.view(np.int32).copy())
    out, gbm = D.reduce_summary(counters, bitmap, first, seed_base, n_total)
    if rank == 0:
        q.put((out, gbm.numpy().view(np.uint32).tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 150), (3, 100)])
def test_sharded_reduction_equals_single_process(oracle, world, n_total):
    import madsim_b200 as m
    from madsim_b200 import dist as D

    seed_base = 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed_base, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, gbm = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wl = m.Workload.raft(duration_ms=1500, bug=1, faults=7)
    ref = oracle.run_many(m.Config(), wl, seed_base, n_total)
    c, bm = D.summary_from_results(ref, seed_base)
    assert out["ev_pop"] == int(c[0]) and out["ev_fire"] == int(c[1]) and out["draws"] == int(c[2])
    assert out["n_pass"] == int(c[3]) and out["n_fail"] == int(c[4]) and out["n_seeds"] == n_total
    assert out["n_fail"] > 0, "the double-vote bug must be caught in this range"
    assert out["first_fail_seed"] == int(c[6])
    assert gbm == bm.tolist()
