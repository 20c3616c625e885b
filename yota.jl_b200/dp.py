"""Batch-sharded data parallelism: one process per GPU, the tape replicated, each rank runs
its column shard of the batch, parameter gradients are summed by NCCL inside
`yota_program_run` (include/yota_cuda.h: yota_comm_init / yota_program_set_allreduce).

The reference has no multi-device code; the loss is a sum over the batch
axis (the last dim, contiguous in column-major), so sharding is exact up to FP32 summation
order.  `torch.distributed` is used only as the rendezvous that ships the ncclUniqueId."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ffi
from .device import context


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_columns(a, rank, nranks):
    """Contiguous column (last-dim) shard of a column-major array; ragged tails go to the
    first ranks."""
    n = a.shape[-1]
    base, extra = divmod(n, nranks)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return np.asfortranarray(a[..., lo:hi])


def init_process_group():
    """gloo rendezvous from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group(backend="gloo")
    return dist


def init_comm(ctx=None):
    """Create the NCCL communicator of this process' context; returns (rank, nranks)."""
    import torch

    rank, nranks, _ = world()
    if nranks == 1:
        return 0, 1
    dist = init_process_group()
    ctx = ctx or context()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        ffi.check(ffi.lib().yota_comm_unique_id(buf))
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    dist.broadcast(t, src=0)
    idb = (C.c_ubyte * 128)(*t.tolist())
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set; keep stdout for results
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        ffi.check(ffi.lib().yota_comm_init(ctx.h, rank, nranks, idb))
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    return rank, nranks


def param_grad_slots(compiled, n_params):
    """Out slots of the gradients of the first `n_params` arguments (the replicated
    parameters); data gradients stay sharded."""
    return [g[1] for g in compiled.low.grads[:n_params] if isinstance(g, tuple) and g[0] == "out"]


def symmetric_outputs(compiled, slots, ctx=None):
    """Output arrays for the all-reduced slots carved from ONE symmetric allocation
    (yota_comm_symm_alloc: collective), so that the library's NVSwitch all-reduce kernel is
    used instead of ncclAllReduce.  Returns ({slot: B200Array view}, base pointer) or
    ({}, None) when symmetric memory is unavailable (single rank, NCCL < 2.28)."""
    from .device import B200Array

    ctx = ctx or context()
    rank, nranks, _ = world()
    if nranks == 1 or not ffi.lib().yota_comm_symm_available(ctx.h):
        return {}, None
    offs, total = {}, 0
    for s in slots:
        dims = compiled._out_dims[s]
        nbytes = 4 * int(np.prod(dims, dtype=np.int64)) if dims else 4
        offs[s] = total
        total += (nbytes + 255) & ~255
    base = C.c_void_p()
    try:
        ffi.check(ffi.lib().yota_comm_symm_alloc(ctx.h, total, C.byref(base)))
    except ffi.YotaError as e:
        # no symmetric windows on this communicator (the call is collective, so every rank lands
        # here): the outputs stay ordinary allocations and are reduced by ncclAllReduce
        import sys
        if rank == 0:
            print(f"yota_b200.dp: symmetric memory unavailable ({e}); using ncclAllReduce", file=sys.stderr)
        return {}, None
    ffi.check(ffi.lib().yota_memset(ctx.h, base, 0, total))
    return {s: B200Array(compiled._out_dims[s], "f32", ctx, ptr=base.value + offs[s], owner=False) for s in slots}, base


def allreduce_host_max(x):
    """max over ranks of a Python float (timing: max over ranks, never the wall clock)."""
    import torch

    rank, nranks, _ = world()
    if nranks == 1:
        return x
    dist = init_process_group()
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier():
    rank, nranks, _ = world()
    if nranks > 1:
        init_process_group().barrier()


def dest_shard(n_dst, rank, nranks):
    """Contiguous 0-based [lo, hi) slice of the n_dst destinations owned by `rank` (ragged tails
    to the first ranks, like shard_columns)."""
    base, extra = divmod(n_dst, nranks)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_add_cols_sharded(dy, idx, n_dst, rank=None, nranks=None, ctx=None):
    """This rank's slice of  dx = zero(rows, n_dst); dx[:, idx[k]] += dy[:, k]  (the embedding
    pullback, src/helpers.jl:7-11) sharded by DESTINATION: every rank scans all (dy, idx), folds only
    the table columns it owns, exchanges nothing -- the concatenation over ranks equals the one-GPU
    result bit for bit.  Returns (B200Array rows x (hi - lo), (lo, hi))."""
    from .device import B200Array

    ctx = ctx or context()
    r, n, _ = world()
    rank, nranks = (r if rank is None else rank), (n if nranks is None else nranks)
    lo, hi = dest_shard(n_dst, rank, nranks)
    rows, n_src = dy.dims
    out = B200Array((rows, hi - lo), "f32", ctx)
    ffi.check(ffi.lib().yota_scatter_add_cols_sharded(ctx.h, out.ptr, rows, n_dst, lo, hi, dy.ptr, idx.ptr, n_src))
    return out, (lo, hi)
