#!/usr/bin/env python
"""Under torchrun on N GPUs: both-stream timeline of one C3 eval (how far the NCCL all-reduces
overlap the pullback sweep) and the cost of the same all-reduces run alone.
   torchrun --nproc-per-node N tools/dp_timeline.py [tf32|bf16] > gpurun_out/timeline.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import yota_b200 as Y  # noqa: E402
from yota_b200 import dp, ffi  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
rank, nranks, local = dp.world()
ctx = Y.context(local)
lib = ffi.lib()
L, W, B = 8, 4096, int(os.environ.get("BATCH", 8192))
args = list(cases.c3_args(L, W, B))
args[-2], args[-1] = dp.shard_columns(args[-2], rank, nranks), dp.shard_columns(args[-1], rank, nranks)
dargs = [Y.cu(a) for a in args]
gf = Y.CompiledGrad(Y.gradtape(cases.make_c3(L, B), *dargs), ctx, {"tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16}[mode])
outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
if nranks > 1:
    dp.init_comm(ctx)
    slots = dp.param_grad_slots(gf, 2 * L) + [gf.low.val[1]]
    gf.set_allreduce(slots)
    if not os.environ.get("YOTA_AR_NCCL"):
        symm, _base = dp.symmetric_outputs(gf, slots, ctx)
        for s_, arr in symm.items():
            outs[s_] = arr
for _ in range(3):
    gf.run(dargs, out=outs)
ctx.sync()
dp.barrier()
tl = gf.timeline(dargs, out=outs, reps=4)
ctx.sync()
dp.barrier()

# the same reductions alone, back to back on the compute stream
alone = None
if nranks > 1:
    e0, e1 = C.c_void_p(), C.c_void_p()
    ffi.check(lib.yota_event_create(ctx.h, C.byref(e0)))
    ffi.check(lib.yota_event_create(ctx.h, C.byref(e1)))
    bufs = [outs[s] for s in dp.param_grad_slots(gf, 2 * L)]
    for rep in range(3):
        if rep == 2:
            ffi.check(lib.yota_event_record(ctx.h, e0))
        for b in bufs:
            n = 1
            for d in b.dims:
                n *= d
            ffi.check(lib.yota_allreduce_f32(ctx.h, b.ptr, n))
    ffi.check(lib.yota_event_record(ctx.h, e1))
    ctx.sync()
    ms = C.c_float()
    ffi.check(lib.yota_event_elapsed_ms(ctx.h, e0, e1, C.byref(ms)))
    alone = ms.value
    dp.barrier()

if rank == 0:
    print(f"# C3 {mode} nranks={nranks} per-rank batch {args[-1].shape[1]}; NCCL max CTAs {os.environ.get('YOTA_NCCL_MAX_CTAS', '16 (default)')}, "
          f"reserve {os.environ.get('YOTA_COMM_SM_RESERVE', '= max CTAs')}")
    print(f"# {'label':60s} {'dev start':>9s} {'dev end':>9s} {'dur':>8s} {'host enq':>9s}  (ms)")
    for label, a, b, h in tl:
        print(f"{label[:60]:60s} {a:9.3f} {b:9.3f} {b - a:8.3f} {h:9.3f}")
    if alone is not None:
        nbytes = sum(4 * int(__import__('numpy').prod(b.dims)) for b in bufs)
        print(f"# the {len(bufs)} parameter-gradient all-reduces alone, back to back: {alone:.3f} ms for {nbytes} B "
              f"-> algBW {nbytes / alone / 1e6:.1f} GB/s, busBW {nbytes * 2 * (nranks - 1) / nranks / alone / 1e6:.1f} GB/s")
