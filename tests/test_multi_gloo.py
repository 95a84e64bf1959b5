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
    # what msim_batch_summarize_sharded writes on the device, restated on the host; then the ONE all-reduce
    buf = torch.from_numpy(D.sharded_summary_from_results(res, first, rank, world, seed_base, n_total))
    D.reduce_summary(buf, n_total, world)
    out, gbm = D.finish_summary(buf.numpy(), n_total, world)
    if rank == 0:
        q.put((out, gbm.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,seed_base", [(2, 150, 40), (3, 100, 40), (3, 2, 41), (2, 70, 2**64 - 200)],
                         ids=["ws2", "ws3", "ws3-empty-shard", "ws2-seeds-above-2^63"])
def test_sharded_reduction_equals_single_process(oracle, world, n_total, seed_base):
    import madsim_b200 as m
    from madsim_b200 import dist as D

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
    if n_total > 50:
        assert out["n_fail"] > 0, "the double-vote bug must be caught in this range"
    assert out["first_fail_seed"] == (int(c[6]) if int(c[4]) else None)
    assert gbm == bm.tolist()
