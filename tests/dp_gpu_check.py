"""Run under torchrun on N GPUs (gpurun --gpus N): batch-sharded grads + NCCL all-reduce must
equal the single-GPU full-batch grads (FP32 summation order aside)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import dp, ffi  # noqa: E402

rank, nranks, local = dp.world()
ctx = Y.context(local)
dp.init_comm(ctx)
for mode, flags, tol in (("fp32", 0, 2e-5), ("tf32", ffi.FLAG_GEMM_TF32, 5e-3)):
    L, W, B = 3, 256, 256 * nranks
    args = cases.c3_args(L, W, B)
    f = cases.make_c3(L, B)
    full = [Y.cu(a) for a in args]
    fv, fg = Y.grad(f, *full, flags=flags)  # replicated full-batch reference on every rank
    shard = list(args)
    shard[-2], shard[-1] = dp.shard_columns(args[-2], rank, nranks), dp.shard_columns(args[-1], rank, nranks)
    dsh = [Y.cu(a) for a in shard]
    gf = Y.CompiledGrad(Y.gradtape(f, *dsh), ctx, flags)

    def check(variant, val, g):
        v = float(val.to_host())
        assert abs(v - fv) <= tol * abs(fv), (mode, variant, v, fv)
        for i in range(2 * L):
            a, b = g[i].to_host(), fg[1 + i].to_host()
            err = np.abs(a - b).max() / np.abs(b).max()
            assert err <= tol, (mode, variant, i, err)
        for i in (2 * L, 2 * L + 1):  # data grads stay sharded
            a, b = g[i].to_host(), dp.shard_columns(fg[1 + i].to_host(), rank, nranks)
            assert np.abs(a - b).max() <= tol * np.abs(b).max(), (mode, variant, i)

    slots = dp.param_grad_slots(gf, 2 * L) + [gf.low.val[1]]
    gf.set_allreduce(slots)
    for variant in ("nccl", "symmetric"):
        outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
        base = None
        if variant == "symmetric":  # the library's own NVSwitch all-reduce kernel
            symm, base = dp.symmetric_outputs(gf, slots, ctx)
            if not symm:
                break
            for s_, arr in symm.items():
                outs[s_] = arr
        for it in range(3):  # eager + repeated runs
            val, g = gf.run(dsh, out=outs)
            ctx.sync()
        check(variant, val, g)
        first = [x.to_host() for x in g[:2 * L]]
        # back to back without a sync in between: the last reductions of one eval overlap the
        # forward sweep of the next (the run no longer joins the communication stream); the
        # results read afterwards must be those of a fully reduced eval
        for it in range(4):
            val, g = gf.run(dsh, out=outs)
        check(variant + "/back-to-back", val, g)
        for a, b in zip(first, g[:2 * L]):  # run to run, and rank to rank below: bit-identical sums
            assert variant == "nccl" or np.array_equal(a.view(np.uint32), b.to_host().view(np.uint32)), (mode, "run-to-run")
        if variant == "symmetric":
            import torch
            import torch.distributed as dist
            h = torch.from_numpy(first[0].copy())
            ref = h.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(h.view(torch.int32), ref.view(torch.int32)), (mode, "ranks hold different sums")
        del outs, g, val
        if base is not None:
            ffi.check(ffi.lib().yota_comm_symm_free(ctx.h, base))
        dp.barrier()
        if rank == 0:
            print(f"DP_GPU_OK mode={mode} allreduce={variant} nranks={nranks} symm={ffi.lib().yota_comm_symm_available(ctx.h)}", flush=True)


# destination-sharded, bit-exact scatter-add over the ranks: each rank folds its
# slice of the table from the replicated (dy, idx); gathered over ranks it equals the one-GPU result
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

r_ = np.random.default_rng(99)
rows, n_dst, n_src = 64, 4096, 50000
V = r_.standard_normal((rows, n_src)).astype(np.float32)
idx = r_.integers(1, n_dst + 1, size=n_src).astype(np.int64)
dV, di = Y.cu(V), Y.cu(idx)
full = Y.B200Array((rows, n_dst))
ffi.check(ffi.lib().yota_scatter_add_cols(ctx.h, full.ptr, rows, n_dst, dV.ptr, di.ptr, n_src))
shard, (lo, hi) = dp.scatter_add_cols_sharded(dV, di, n_dst)
mine = torch.from_numpy(np.ascontiguousarray(shard.to_host().T))  # (hi - lo, rows): rows of the gather are destinations
parts = [torch.empty((dp.dest_shard(n_dst, q, nranks)[1] - dp.dest_shard(n_dst, q, nranks)[0], rows)) for q in range(nranks)]
dist.all_gather(parts, mine)
got = torch.cat(parts, dim=0).numpy().T
assert np.array_equal(np.ascontiguousarray(got).view(np.uint32), np.ascontiguousarray(full.to_host()).view(np.uint32)), "sharded scatter-add differs"
if rank == 0:
    print(f"DP_GPU_OK dest_sharded_scatter nranks={nranks}", flush=True)
