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


def sharded_layout(n_total: int, world: int):
    """int64 element counts of the one-all-reduce summary buffer (msim_batch_summarize_sharded): 8 counters,
    `world` first-fail slots, then the global fail bitmap (u32 words, padded to whole int64s)."""
    words = (n_total + 31) // 32
    return 8, world, (words + 1) // 2


def finish_summary(buf: np.ndarray, n_total: int, world: int):
    """Host side of the reduction: `buf` = the all-reduced (summed) int64 buffer as a numpy array.
    → (dict of global counters, global bitmap uint32[ceil(n_total/32)])."""
    u = buf.view(np.uint64)
    slots = u[8:8 + world]
    first = int(slots.min()) if world else U64_MAX  # unsigned MIN over the ranks' slots (all-ones = no failure)
    out = {
        "ev_pop": int(u[0]), "ev_fire": int(u[1]), "draws": int(u[2]), "n_pass": int(u[3]), "n_fail": int(u[4]),
        "n_unfinished": int(u[5]), "n_seeds": int(u[7]), "first_fail_seed": None if first == U64_MAX else first,
    }
    bitmap = buf[8 + world:].view(np.uint32)[: (n_total + 31) // 32].copy()
    return out, bitmap


def reduce_summary(buf: torch.Tensor, n_total: int, world: int, group=None):
    """ONE collective: sum all-reduce of the whole sharded summary buffer (int64 view).  Every rank's buffer is
    zero outside its own first-fail slot and bit range, so the sum is the global counters, every rank's first
    failing seed in its slot (the MIN is taken on the host, unsigned — seeds near 2^64 are fine) and the OR of the
    disjoint bitmaps.  No host synchronisation here; the caller copies the result back once."""
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


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


def sharded_summary_from_results(res: np.ndarray, seed_first: int, rank: int, world: int, seed_base: int, n_total: int):
    """Host restatement of msim_batch_summarize_sharded for a results array (used by the CPU tests): the int64
    buffer one rank contributes to the all-reduce."""
    nc, ns, nb = sharded_layout(n_total, world)
    buf = np.zeros(nc + ns + nb, dtype=np.int64)
    u = buf.view(np.uint64)
    c, bm = summary_from_results(res, seed_first)
    u[:6] = c[:6]
    u[7] = c[7]
    u[8 + rank] = c[6]  # own slot: first failing seed or all-ones
    bits = buf[nc + ns:].view(np.uint32)
    off = seed_first - seed_base
    for i in range(len(res)):
        if (bm[i // 32] >> np.uint32(i % 32)) & np.uint32(1):
            g = off + i
            bits[g // 32] |= np.uint32(1 << (g % 32))
    return buf


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
        self._buf = None
        self._host = None

    def run(self, seed_base: int, n_total: int, sync: bool = True):
        self.seed_base, self.n_total = seed_base, n_total
        self.first, self.count = shard(seed_base, n_total, self.rank, self.world)
        self.ran = self.count > 0
        if self.ran:  # a rank whose shard is empty (world > n_total) launches nothing but still joins the reduction
            self.batch.run(self.first, self.count, sync=sync)
        return self

    def reduce(self):
        """Device summary (one kernel) → ONE sum all-reduce (NCCL) → one D2H copy → global counters + bitmap."""
        from .api import Batch

        n64 = sum(sharded_layout(self.n_total, self.world))
        if self._buf is None or self._buf.numel() != n64:
            self._buf = torch.empty(n64, dtype=torch.int64, device=self.device)
            self._host = torch.empty(n64, dtype=torch.int64, pin_memory=True)
        with torch.cuda.stream(self.stream):
            if self.ran:
                self.batch.summarize_sharded_into(self._buf.data_ptr(), n64 * 8, self.rank, self.world, self.seed_base,
                                                  self.n_total)
            else:  # nothing ran here: zeros, and "no failure" in the own slot
                self._buf.zero_()
                self._buf[8 + self.rank] = -1
            reduce_summary(self._buf, self.n_total, self.world)
            self._host.copy_(self._buf, non_blocking=True)
        self.stream.synchronize()
        return finish_summary(self._host.numpy(), self.n_total, self.world)

    def close(self):
        self.batch.close()
