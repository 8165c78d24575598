"""Multi-GPU plumbing, caller-driven form: one process per GPU, histories sharded by id range, batch grid all-reduced.

Since round 2 the PRODUCT path for several GPUs lives inside the C library (omc_gpu_comm_init / omc_gpu_multi_*,
include/ompmc_b200.h: the same sharding rule -- omc_gpu_shard_range -- and the same all-reduce-before-accumEndep, on a side
stream); bench.py uses that.  This module keeps the form in which the CALLER owns the collective (any torch.distributed
backend on the pointers of omc_gpu_device_ptrs()), which is also what the world-size-2 gloo tests run on CPU.

The reference's only parallelism is the OpenMP ``parallel for`` over the histories of one batch with
a shared dose grid (omc_dosxyz.c:1252-1259, :690-691).  Here every rank transports a contiguous slice
of the batch's history ids into its private batch grid; the grids are summed over ranks BEFORE
accumEndep() squares them, so the batch statistics (accum, accum2) are exactly those of a
single-GPU run of the same batch.  History id -> RNG stream, hence the result does
not depend on the number of ranks beyond fp64 summation order.
"""
from __future__ import annotations

import numpy as np


def shard_range(first: int, n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of [first, first+n) for ``rank``; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = first + rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


class DevicePtr:
    """Wrap a raw device pointer for torch.as_tensor() via __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def allreduce_gpu_grid(tr, group=None) -> None:
    """In-place NCCL sum of the batch grid over ranks, ordered after the transport kernels on the
    context's stream (the kernels and the collective share it through torch's ExternalStream)."""
    import torch
    import torch.distributed as dist
    endep, _, _, nreg = tr.device_ptrs()
    t = torch.as_tensor(DevicePtr(endep, nreg), device=f"cuda:{tr.device}")
    with torch.cuda.stream(torch.cuda.ExternalStream(tr.stream_ptr(), device=f"cuda:{tr.device}")):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def allreduce_cpu_grid(tr, group=None) -> None:
    """Same exchange for a CPU checker transport (gloo), used by the world_size-2 host-logic tests."""
    import torch
    import torch.distributed as dist
    g = torch.from_numpy(np.ascontiguousarray(tr.get_endep()))
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    tr.set_endep(g.numpy())


def run_batch_sharded(tr, first: int, nperbatch: int, rank: int, world: int, allreduce=None) -> None:
    """One iteration of the reference batch loop (omc_dosxyz.c:1237-1263) spread over ``world`` ranks."""
    lo, n = shard_range(first, nperbatch, rank, world)
    if n > 0:
        tr.run_histories(lo, n)
    if world > 1:
        allreduce(tr)
    tr.accum_batch()


def settle_completed(tr, world: int, allreduce=None) -> int:
    """Sum over ranks and accumulate (accumEndep) every completed batch grid that is waiting, oldest first."""
    n = 0
    while tr.completed_batches() > 0:
        if world > 1:
            allreduce(tr)
        tr.accum_batch()
        n += 1
    return n


def start_batch_sharded(tr, first: int, nperbatch: int, rank: int, world: int, allreduce=None) -> None:
    """Pipelined form of run_batch_sharded: this rank's slice of the batch is started while the tail of the previous
    batch is still in flight; whatever batch completed meanwhile is reduced over ranks and accumulated.  Every rank
    completes batch k-1 inside its start of batch k, so the collective calls line up across ranks."""
    lo, n = shard_range(first, nperbatch, rank, world)
    tr.start_batch(lo, n)            # (also when this rank's slice is empty: it still owes the batch its all-reduce)
    settle_completed(tr, world, allreduce)


def finish_batches_sharded(tr, rank: int, world: int, allreduce=None) -> None:
    tr.finish_batches()
    settle_completed(tr, world, allreduce)
