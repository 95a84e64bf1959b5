"""Multi-GPU: seeds shard by contiguous ranges, one process per GPU (torch.distributed).

The reference's only many-seed parallelism is `MADSIM_TEST_JOBS` threads over `seed..seed+count`
(runtime/builder.rs:129-150); runtimes share nothing, so there is NO data-path collective.  The
single exchange is the end-of-run reduction of the batch summary:
    counters  {ev_pop, ev_fire, draws, n_pass, n_fail, n_unfinished}   SUM
    first failing seed                                                  MIN
    pass/fail bitmap (each rank owns a disjoint bit range)              SUM  (== OR on disjoint bits)
over NCCL on the GPU box (gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

U64_MAX = (1 << 64) - 1


def shard(seed_base: int, n_total: int, rank: int, world: int):
    """Contiguous seed range of `rank`: (first_seed, count).  Earlier ranks take the remainder."""
    per, rem = divmod(n_total, world)
    count = per + (1 if rank < rem else 0)
    first = seed_base + rank * per + min(rank, rem)
    return first, count


def summary_layout(n_local: int):
    """Element counts of the device summary written by msim_batch_summarize_device: 8 u64 + bitmap u32."""
    return 8, (n_local + 31) // 32


def reduce_summary(counters: torch.Tensor, bitmap_local: torch.Tensor, first_seed_local: int, seed_base: int,
                   n_total: int, group=None):
    """All-reduce one rank's summary into the global one.

    counters: int64[8] on the collective's device = {ev_pop, ev_fire, draws, n_pass, n_fail,
              n_unfinished, first_fail_seed (u64 bit pattern, all-ones = none), n_seeds}
    bitmap_local: int32[ceil(n_local/32)] — bit i = local seed i failed
    Returns (dict of global counters, global bitmap int32[ceil(n_total/32)] on the same device).
    """
    dev = counters.device
    sums = counters[[0, 1, 2, 3, 4, 5, 7]].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    # MIN over u64 seeds done in int64 space: map all-ones (none) to INT64_MAX
    ff = counters[6:7].clone()
    ff = torch.where(ff < 0, torch.full_like(ff, torch.iinfo(torch.int64).max), ff)
    dist.all_reduce(ff, op=dist.ReduceOp.MIN, group=group)
    # scatter the local bits into the global bitmap at this rank's offset, then SUM (disjoint → OR)
    words = (n_total + 31) // 32
    g = torch.zeros(words, dtype=torch.int64, device=dev)
    off = first_seed_local - seed_base
    n_local = int(counters[7].item())
    if n_local:
        lb = bitmap_local.to(torch.int64) & 0xFFFFFFFF
        idx = torch.arange(n_local, device=dev)
        bits = (lb[idx // 32] >> (idx % 32)) & 1
        gi = idx + off
        g.index_add_(0, gi // 32, bits << (gi % 32))
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    first = int(ff.item())
    out = {
        "ev_pop": int(sums[0]), "ev_fire": int(sums[1]), "draws": int(sums[2]), "n_pass": int(sums[3]),
        "n_fail": int(sums[4]), "n_unfinished": int(sums[5]), "n_seeds": int(sums[6]),
        "first_fail_seed": None if first == torch.iinfo(torch.int64).max else first,
    }
    return out, g.to(torch.int32) if words else g.to(torch.int32)


def summary_from_results(res: np.ndarray, seed_first: int):
    """Host restatement of msim_summary_kernel for a results array (used by the CPU tests)."""
    n = len(res)
    failed = (res["status"] != 1) & (res["status"] != 0)
    c = np.zeros(8, dtype=np.uint64)
    c[0], c[1], c[2] = res["ev_pop"].sum(), res["ev_fire"].sum(), res["draws"].sum()
    c[3], c[4], c[5] = (res["status"] == 1).sum(), failed.sum(), (res["status"] == 0).sum()
    c[6] = np.uint64(seed_first + int(np.nonzero(failed)[0][0])) if failed.any() else np.uint64(U64_MAX)
    c[7] = n
    bm = np.zeros((n + 31) // 32, dtype=np.uint32)
    for i in np.nonzero(failed)[0]:
        bm[i // 32] |= np.uint32(1 << (i % 32))
    return c, bm


class ShardedRun:
    """Run `n_total` seeds of a workload across the ranks of the default process group."""

    def __init__(self, config, workload, device_index: int):
        from .api import Batch

        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = torch.device("cuda", device_index)
        self.batch = Batch(config, device_index).load(workload)
        self.stream = torch.cuda.Stream(device=self.device)
        self.batch.set_stream(self.stream.cuda_stream)
        self._summary = None

    def run(self, seed_base: int, n_total: int, sync: bool = True):
        self.seed_base, self.n_total = seed_base, n_total
        self.first, self.count = shard(seed_base, n_total, self.rank, self.world)
        self.batch.run(self.first, self.count, sync=sync)
        return self

    def reduce(self):
        """Device summary → NCCL all-reduce → global counters + bitmap."""
        from .api import Batch

        nbytes = Batch.summary_nbytes(self.count)
        buf = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device=self.device)
        with torch.cuda.stream(self.stream):
            self.batch.summarize_into(buf.data_ptr(), nbytes)
            counters = buf[:8]
            bitmap = buf[8:].view(torch.int32)[: (self.count + 31) // 32]
            if self.world > 1:
                out = reduce_summary(counters, bitmap, self.first, self.seed_base, self.n_total)
            else:
                self.stream.synchronize()
                c = counters.cpu().numpy().view(np.uint64)
                out = ({"ev_pop": int(c[0]), "ev_fire": int(c[1]), "draws": int(c[2]), "n_pass": int(c[3]),
                        "n_fail": int(c[4]), "n_unfinished": int(c[5]), "n_seeds": int(c[7]),
                        "first_fail_seed": None if int(c[6]) == U64_MAX else int(c[6])}, bitmap.clone())
        self.stream.synchronize()
        return out

    def close(self):
        self.batch.close()
