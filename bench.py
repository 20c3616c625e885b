#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: `Yota.grad` evals/sec of the 8-layer
4096-wide tanh MLP with softmax cross-entropy at batch 8192 (config C3),
one eval = one warm (val, grads-for-every-argument) call; plus the metric's second half, the
pullback HBM GB/s of the broadcast chain (C2) and of the scatter-add pullback (C5), as
`secondary`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1|c4]
                  [--gemm tf32|tf32x3|fp32|bf16] [--impl reference]

Prints ONE JSON line (rank 0).  N > 1: launched by torchrun, one rank per GPU; the global
batch of 8192 is sharded by columns ("strong" scaling), parameter gradients are all-reduced
over NVSwitch inside the run; `weak_scaling` repeats the measurement with 8192 columns per GPU.
`--impl reference` times the restated reference CPU path (the oracle: op-by-op NumPy/OpenBLAS,
not Julia -- no Julia toolchain exists in this image) on a bounded sample, all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], 0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                pw.append(float(f[2]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        # "under load": samples drawing more than half the maximum power seen (the sampler also
        # catches the idle edges of the timed region)
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * max(pw)]
        return {"sm_mhz": float(np.median(load or sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(load), "power_w_max": max(pw) if pw else None}


# ---- workloads --------------------------------------------------------------------------
def workload(name, rank, nranks, small=False, weak=False):
    """-> dict(f, args(host), n_params, label, flops, hbm_bytes, units) for this rank's shard."""
    import cases
    from yota_b200 import dp

    if name == "c3":
        L, W, B = (8, 4096, 8192) if not small else (3, 256, 512)
        args = list(cases.c3_args(L, W, B))
        if not weak:  # strong scaling: the 8192 columns are sharded; weak: every rank runs all 8192
            args[-2], args[-1] = dp.shard_columns(args[-2], rank, nranks), dp.shard_columns(args[-1], rank, nranks)
        Bl = args[-1].shape[1]
        Bg = B * nranks if weak else B
        return dict(f=cases.make_c3(L, Bg), args=args, n_params=2 * L, batch_args=(2 * L, 2 * L + 1), global_batch=Bg,
                    label=f"C3: {L}-layer {W}-wide tanh MLP, softmax-xent, batch {Bg} (Yota.grad, grads for all args)",
                    gemm_flops=3 * L * 2 * W * W * Bl, gemm_shape=(W, Bl, W), hbm_bytes=None)
    if name == "c2":
        n = 8192 if not small else 1024
        x, Wm, b = cases.c2_args(n)
        x, Wm = dp.shard_columns(x, rank, nranks), dp.shard_columns(Wm, rank, nranks)
        return dict(f=cases.c2_chain, args=[x, Wm, b], n_params=0, batch_args=(0, 1), ar_args=(2,),
                    label=f"C2: sum(tanh.(x .* W .+ b) .^ 2), {n}x{n}", gemm_flops=0, hbm_bytes=4 * x.size * 4,
                    hbm_step="region")
    if name == "c5":
        rows, nt, ni = (256, 1 << 20, 1 << 22) if not small else (256, 1 << 12, 1 << 14)
        E, idx, V = c5_args_fast(rows, nt, ni)
        return dict(f=cases.c5_embed, args=[E, idx, V], n_params=0, batch_args=(),
                    label=f"C5: sum(E[:, idx] .* V), E {rows}x{nt}, {ni} Int64 indices (bit-exact scatter-add pullback)",
                    gemm_flops=0, hbm_bytes=V.nbytes + idx.nbytes + E.nbytes, hbm_step="ungetindex",
                    # algorithmic bytes of the other steps of the eval (reported beside the scatter-add pullback)
                    # (fused pass: gathers E[:, idx] itself, reads V, writes dV; the loss is a sink)
                    hbm_other={"region": 3 * V.nbytes + idx.nbytes},
                    # the index sort of the scatter-add runs on a side stream beside the forward pass, so the
                    # steps overlap in the loop: kernel times are those of the serialised per-step pass
                    # (sort + fold back to back), not shares of the loop time
                    serialised_kernel_times=True)
    if name == "c1":
        return dict(f=cases.c1_mlp, args=list(cases.c1_args()), n_params=4, batch_args=(4, 5),
                    label="C1: MLP 784-128-10 tanh, batch 64", gemm_flops=3.9e7, hbm_bytes=None)
    if name == "c4":
        H, B, T = (1024, 128, 128) if not small else (128, 32, 8)
        return dict(f=cases.make_c4(T), args=list(cases.c4_args(H, B, T)), n_params=3, batch_args=(),
                    label=f"C4: unrolled Elman RNN H={H} T={T} B={B}", gemm_flops=T * 6 * 2 * H * H * B, hbm_bytes=None)
    raise SystemExit(f"unknown workload {name}")


def c5_args_fast(rows, nt, ni, seed=5):
    """C5 inputs (E ~ N(0,1), idx ~ U{1..nt}, V ~ N(0,1)): V is a 2^18-column block of normals
    repeated (4 GiB of normals take ~10 s of host RNG; the values only have to be non-trivial
    Float32 addends, the indices -- which decide the access pattern -- are all drawn)."""
    r = np.random.default_rng(seed)
    E = np.asfortranarray(r.standard_normal((rows, nt), dtype=np.float32))
    idx = r.integers(1, nt + 1, size=ni).astype(np.int64)
    blk = min(ni, 1 << 18)
    Vb = r.standard_normal((blk, rows), dtype=np.float32)
    V = np.empty((ni, rows), dtype=np.float32)
    for c0 in range(0, ni, blk):
        V[c0:c0 + blk] = Vb[:min(blk, ni - c0)]
    return E, idx, V.T  # V.T: (rows, ni), column-major


# ---- reference arm ------------------------------------------------------------------------
def run_reference(a):
    """Reference arm: the restated reference CPU path (oracle) on the host cores, rank 0 only.
    Every step is one oracle eval of a BOUNDED SAMPLE (a leading block of the batch columns /
    indices), sized after a probe so that warm-up + steps fit the time budget; `ms_per_step` is the
    measured time of such a step, `value` the full-workload rate it implies (cost is linear in the
    batch), and `config.sample` says which."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    blas = None
    limiter = None
    try:  # torchrun exports OMP_NUM_THREADS=1: raise the BLAS pool back to every core and say so
        from threadpoolctl import threadpool_info, threadpool_limits
        limiter = threadpool_limits(limits=cores)
        blas = [{"lib": i.get("internal_api"), "threads": i.get("num_threads")} for i in threadpool_info() if i.get("user_api") == "blas"]
    except Exception:
        pass
    from oracle import yota_oracle as O

    wl = workload(a.workload, 0, 1, a.small)
    f, args = wl["f"], wl["args"]
    budget = float(os.environ.get("YOTA_REF_BUDGET_S", "150"))
    n_warm = max(1, min(a.warmup, 2))
    n_evals = n_warm + a.steps

    def sample(sub):
        s = list(args)
        if a.workload == "c3":
            s[-2], s[-1] = args[-2][:, :sub], args[-1][:, :sub]
        elif a.workload == "c5":
            s[1], s[2] = args[1][:sub], np.asfortranarray(args[2][:, :sub])
        return s

    full = {"c3": args[-1].shape[1], "c5": args[1].shape[0]}.get(a.workload)
    scale, note, sargs = 1.0, "full workload", args
    if full is not None and not a.small:
        # probe on 1/16 of the units, then the largest power-of-two block that fits the budget
        probe = max(1, full // 16)
        pa = sample(probe)
        O.grad(f, *pa)  # first call: page-in, BLAS thread start-up
        t0 = time.perf_counter()
        O.grad(f, *pa)
        t_probe = time.perf_counter() - t0
        sub = probe
        while sub * 2 <= full and (t_probe * (sub * 2 / probe)) * n_evals <= budget:
            sub *= 2
        if sub < full:
            sargs, scale = sample(sub), full / sub
            unit = "batch columns" if a.workload == "c3" else "indices"
            note = (f"{unit} 1..{sub} of {full} per step (measured step time x {scale:g} = one full eval: "
                    "every cost of the eval is linear in the batch)")
    for _ in range(n_warm):
        O.grad(f, *sargs)
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        O.grad(f, *sargs)
        ts.append(time.perf_counter() - t0)
    t_step = float(np.mean(ts))
    val = 1.0 / (t_step * scale)
    line = {"impl": "reference", "metric": "Yota.grad evals/sec", "value": val, "unit": "evals/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": n_warm, "ms_per_step": 1e3 * t_step, "full_eval_ms": 1e3 * t_step * scale,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "sample": note,
                       "note": "restated reference CPU path (NumPy/OpenBLAS oracle), not Julia; ms_per_step is the "
                               "measured time of one sampled step, value = 1 / full_eval_ms"},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "blas": blas, "kind": "port", "sample": note},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    del limiter
    print(json.dumps(line))


# ---- helpers for the device arm -------------------------------------------------------------
class Timer:
    def __init__(self, ctx, lib, ffi):
        self.ctx, self.lib, self.ffi = ctx, lib, ffi
        self.e0, self.e1 = self._ev(), self._ev()

    def _ev(self):
        e = C.c_void_p()
        self.ffi.check(self.lib.yota_event_create(self.ctx.h, C.byref(e)))
        return e

    def start(self):
        self.ctx.sync()
        self.ffi.check(self.lib.yota_event_record(self.ctx.h, self.e0))

    def stop(self):
        """ms between start() and now on the compute stream (joins in-flight reductions first)."""
        self.ffi.check(self.lib.yota_event_record(self.ctx.h, self.e1))
        self.ctx.sync()
        ms = C.c_float()
        self.ffi.check(self.lib.yota_event_elapsed_ms(self.ctx.h, self.e0, self.e1, C.byref(ms)))
        return ms.value


def step_groups(prof):
    groups = {}
    for label, t in prof:
        key = label.split(" ")[0].split("[")[0]
        g = groups.setdefault(key, [0.0, 0])
        g[0] += t
        g[1] += 1
    return groups, sum(t for _, t in prof)


def hbm_leg(Y, ffi, ctx, lib, name, steps, pk, small):
    """One HBM-bound workload (C2 / C5) measured in this process: evals/s and the GB/s of the
    kernel BASELINE.json's metric names (its share of the step from the per-step event pass)."""
    wl = workload(name, 0, 1, small)
    dargs = [Y.cu(x) for x in wl["args"]]
    gf = Y.CompiledGrad(Y.gradtape(wl["f"], *dargs), ctx, 0)
    outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
    for _ in range(3):
        gf.run(dargs, out=outs)
    tm = Timer(ctx, lib, ffi)
    tm.start()
    for _ in range(steps):
        gf.run(dargs, out=outs)
    ms_step = tm.stop() / steps
    prof = None
    for _ in range(2):
        prof = gf.profile(dargs, out=outs)
    groups, total = step_groups(prof)
    dom = wl["hbm_step"]
    ser = bool(wl.get("serialised_kernel_times"))
    scale = 1.0 if ser else ms_step / total
    t_dom = groups[dom][0] * scale
    ach = wl["hbm_bytes"] / (t_dom * 1e-3) / 1e9
    out = {"workload": wl["label"], "evals_per_s": 1e3 / ms_step, "ms_per_step": ms_step, "kernel": dom,
           "algorithmic_bytes": wl["hbm_bytes"], "kernel_ms": t_dom, "achieved": ach, "unit": "GB/s", "peak": pk["hbm"],
           "frac": ach / pk["hbm"], "peak_source": f"{pk['src']} HBM copy", "launches_per_eval": gf.stats()["launches"],
           "kernel_ms_from": "serialised per-step pass (steps overlap in the loop)" if ser else "loop time x share of the per-step pass"}
    for k, b in (wl.get("hbm_other") or {}).items():
        if k in groups:
            t = groups[k][0] * scale
            out[f"other:{k}"] = {"achieved": b / (t * 1e-3) / 1e9, "frac": b / (t * 1e-3) / 1e9 / pk["hbm"]}
    del gf, outs, dargs
    return out


def cublas_peaks():
    """Kernel-independent tensor-core denominators measured in THIS run on THIS GPU: torch.matmul
    (cuBLAS) at 8192^3 in TF32 and BF16, best of 10 (the method of MEASURED_PEAKS.json)."""
    try:
        import torch

        dev = torch.device("cuda", torch.cuda.current_device())
        out = {}
        n = 8192
        for name, dt, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.randn(n, n, device=dev, dtype=dt)
            b = torch.randn(n, n, device=dev, dtype=dt)
            for _ in range(3):
                torch.matmul(a, b)
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = 2 * n ** 3 / (best * 1e-3) / 1e12
            if name == "tf32":  # ... and back to back for ~2 s: what the same library holds under the power cap
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = max(20, int(2000 / best))
                e0.record()
                for _ in range(reps):
                    torch.matmul(a, b)
                e1.record()
                e1.synchronize()
                out["tf32_sustained"] = 2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
            del a, b
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


MODE_DTYPE = {"fp32": "f32", "tf32": "tf32 (f32 storage, f32 accumulate)",
              "tf32x3": "3xtf32 (f32 storage split into tf32 hi + lo, three tensor-core products per term, f32 accumulate)",
              "bf16": "bf16 (f32 storage + bf16 operand shadows, f32 accumulate)"}
MODE_GATE = {"fp32": "rtol 1e-5 (reference gradcheck)", "tf32x3": "rtol 1e-5 (reference gradcheck)",
             "tf32": "5e-3 of max|g|", "bf16": "3e-2 of max|g|"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--gemm", default=os.environ.get("YOTA_GEMM_MODE", "tf32"),
                    help="contraction mode: tf32 (default: tcgen05, Float32 storage), tf32x3 (compensated, rtol 1e-5), "
                         "bf16 (tcgen05, BF16 shadows), fp32 (strict FFMA)")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other GEMM modes")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C5 HBM legs")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling leg")
    ap.add_argument("--small", action="store_true", help="tiny shapes (CI smoke of the bench itself)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (tuning sweeps): no e2e, roofline or CPU legs")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.impl == "reference":
        return run_reference(a)

    import yota_b200 as Y
    from yota_b200 import dp, ffi

    rank, nranks, local = dp.world()
    ctx = Y.context(local)
    lib = ffi.lib()
    MODE_FLAGS = {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16, "tf32x3": ffi.FLAG_GEMM_TF32X3}
    flags = MODE_FLAGS[a.gemm]
    if nranks > 1:
        dp.init_comm(ctx)
    tm = Timer(ctx, lib, ffi)

    def setup(wl):
        dargs = [Y.cu(x) for x in wl["args"]]
        tape = Y.gradtape(wl["f"], *dargs)
        gf = Y.CompiledGrad(tape, ctx, flags)
        outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
        ar_impl, symm_base = None, None
        if nranks > 1:
            slots = dp.param_grad_slots(gf, wl["n_params"]) + [gf.low.val[1]]
            for i in wl.get("ar_args", ()):
                slots.append(gf.low.grads[i][1])
            gf.set_allreduce(slots)
            # all-reduced outputs live in symmetric memory: reduced by the library's own NVSwitch kernel
            symm, symm_base = dp.symmetric_outputs(gf, slots, ctx)
            for s_, arr in symm.items():
                outs[s_] = arr
            ar_impl = "own multimem kernel over NCCL symmetric windows" if symm else "ncclAllReduce"
        return dargs, tape, gf, outs, ar_impl, symm_base

    def timed(gf, dargs, outs, steps, warmup):
        """K evals back to back, inputs resident in HBM; max over ranks of the device time."""
        for _ in range(warmup):
            gf.run(dargs, out=outs)
        ctx.sync()
        dp.barrier()
        tm.start()
        for _ in range(steps):
            gf.run(dargs, out=outs)
        ms = tm.stop()
        dp.barrier()
        return dp.allreduce_host_max(ms) / steps

    wl = workload(a.workload, rank, nranks, a.small)
    hargs = wl["args"]
    dargs, tape, gf, outs, ar_impl, symm_base = setup(wl)
    stats = gf.stats()

    # ---- device-resident timing: K evals, inputs already in HBM -------------------------
    clk = ClockSampler(local)
    for _ in range(a.warmup):
        gf.run(dargs, out=outs)
    ctx.sync()
    clk.start()
    ms_step = timed(gf, dargs, outs, a.steps, 0)
    clocks = clk.stop()
    value = 1e3 / ms_step
    loss = outs[gf.low.val[1]].to_host()

    if a.quick:
        if rank == 0:
            print(json.dumps({"metric": "Yota.grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": nranks,
                              "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "quick": True,
                              "config": {"workload": wl["label"], "gemm_mode": a.gemm}, "clocks": clocks}))
        return

    # ---- end to end: each step uploads its batch from pinned host memory, reads the loss --
    bidx = wl["batch_args"]
    pinned, h2d = [], 0
    for i in bidx:
        p = C.c_void_p()
        ffi.check(lib.yota_host_alloc(ctx.h, hargs[i].nbytes, C.byref(p)))
        hcol = np.asfortranarray(hargs[i])  # (named: a temporary would be freed before memmove reads it)
        C.memmove(p, hcol.ctypes.data, hcol.nbytes)
        pinned.append((i, p, hargs[i].nbytes))
        h2d += hargs[i].nbytes
    # the loss of every step is read back into pinned host memory by an asynchronous copy ordered
    # after that step (and after the all-reduce of the loss, N > 1), one slot per step parity, so
    # the host is already enqueueing step i+1 while step i runs; checked after the final sync
    hl = C.c_void_p()
    ffi.check(lib.yota_host_alloc(ctx.h, 8, C.byref(hl)))
    hloss = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_float)), shape=(2,))
    # double-buffered batch: the upload of step i+1 (copy stream) overlaps the sweep of step i
    bufs = [list(dargs), list(dargs)]
    for i in bidx:
        bufs[1][i] = Y.B200Array(dargs[i].dims, dargs[i].dtype, ctx)

    def upload(k):
        for i, p, nb in pinned:
            ffi.check(lib.yota_h2d_async(ctx.h, bufs[k][i].ptr, p, nb))

    def e2e_steps(n):
        upload(0)
        ffi.check(lib.yota_stream_fence(ctx.h, 1, 0))
        for it in range(n):
            k = it & 1
            ffi.check(lib.yota_stream_fence(ctx.h, 0, 1))  # copy stream: runs <= it-1 are done with buffer k^1
            val, _ = gf.run(bufs[k], out=outs)
            if it + 1 < n:
                upload(k ^ 1)
            ffi.check(lib.yota_d2h_async(ctx.h, hl.value + 4 * k, val.ptr, 4))  # the step's result, read every step
            ffi.check(lib.yota_stream_fence(ctx.h, 1, 0))  # next run waits for its batch
        ctx.sync()
        assert np.isfinite(hloss).all(), hloss

    e2e_steps(4)
    dp.barrier()
    tm.start()
    e2e_steps(a.steps)
    e2e_ms = tm.stop()
    dp.barrier()
    e2e_ms = dp.allreduce_host_max(e2e_ms) / a.steps
    for _, p, _nb in pinned:
        lib.yota_host_free(ctx.h, p)
    e2e_loss = float(hloss[(a.steps - 1) & 1])
    lib.yota_host_free(ctx.h, hl)
    del bufs

    # ---- roofline of the dominant kernel: per-step CUDA events (serialised pass) ----------
    pk = peaks()
    prof = None
    for _ in range(3):
        prof = gf.profile(dargs, out=outs)
    groups, total_prof = step_groups(prof)
    dom = max(groups, key=lambda k: groups[k][0])
    roof = None
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_tab = {}
    # tensor-core denominators of this very run (cuBLAS TF32 / BF16 at 8192^3), N = 1 only
    cub = cublas_peaks() if (nranks == 1 and dom == "matmul" and not a.small and a.gemm != "fp32") else {}
    if dom == "matmul" and wl["gemm_flops"]:
        n_g = groups[dom][1]
        share = groups[dom][0] / total_prof
        # duration inside the timed loop = the step time measured above x the kernel's share of the
        # step (from the per-step event pass); the event pass itself serialises the steps and runs
        # cooler under the power cap, so its own durations are reported as "isolated"
        avg_ms = ms_step * share / n_g
        ach = wl["gemm_flops"] / n_g / (avg_ms * 1e-3) / 1e12
        tf32_like = a.gemm in ("tf32", "tf32x3")
        derived_burst = pk["tc_burst"] * (0.5 if tf32_like else 1.0)
        derived_sust = pk["tc_sust"] * (0.5 if tf32_like else 1.0)
        own = cub.get("tf32" if tf32_like else "bf16")
        # denominator: the cuBLAS GEMM of the same arithmetic type measured in this run (burst: best
        # of 10, the stricter of the two regimes); MEASURED_PEAKS-derived fractions beside it
        peak = own or derived_burst
        tr = traffic_tab.get(f"gemm ({a.gemm})") if a.workload == "c3" and not a.small and nranks == 1 else None
        roof = {"bound": "tensor", "kernel": f"gemm ({a.gemm})", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "algorithmic_flops_per_launch": wl["gemm_flops"] / n_g,
                "frac_of_bf16_burst_derived": ach / derived_burst, "frac_of_bf16_sustained_derived": ach / derived_sust,
                "frac_of_cublas_tf32_sustained_in_run": (ach / cub["tf32_sustained"]) if tf32_like and cub.get("tf32_sustained") else None,
                "avg_launch_ms": avg_ms, "avg_launch_ms_isolated": groups[dom][0] / n_g, "launches_per_step": n_g,
                "share_of_step": share,
                "peak_source": (f"cuBLAS {'TF32' if tf32_like else 'BF16'} 8192^3 best of 10, measured in this run" if own else
                                f"{pk['src']} bf16 burst" + (" x 0.5 (TF32 runs at half the BF16 rate)" if tf32_like else ""))
                               + ("; 3xTF32 issues 3 tensor-core products per algorithmic multiply-add, achieved counts algorithmic FLOPs" if a.gemm == "tf32x3" else "")
                               + ("; strict-FP32 FFMA kernel, not a tensor-core kernel" if a.gemm == "fp32" else ""),
                "cublas_in_run_tflops": cub or None}
    elif wl.get("hbm_bytes"):
        dom = wl.get("hbm_step", dom)  # the kernel BASELINE.json names for this workload
        t_iso = groups[dom][0]
        ser = bool(wl.get("serialised_kernel_times"))
        hscale = 1.0 if ser else ms_step / total_prof
        t_dom = t_iso * hscale  # in-loop duration: step time x share (see the GEMM branch), unless steps overlap
        ach = wl["hbm_bytes"] / (t_dom * 1e-3) / 1e9
        tr = traffic_tab.get(f"{dom}:{a.workload}") if not a.small else None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "avg_launch_ms": t_dom / groups[dom][1], "avg_launch_ms_isolated": t_iso / groups[dom][1],
                "launches_per_step": groups[dom][1], "share_of_step": t_iso / total_prof, "peak_source": f"{pk['src']} HBM copy"}
        if wl.get("hbm_other"):
            roof["other_steps"] = {k: {"achieved": b / (groups[k][0] * hscale * 1e-3) / 1e9,
                                       "frac": b / (groups[k][0] * hscale * 1e-3) / 1e9 / pk["hbm"]}
                                   for k, b in wl["hbm_other"].items() if k in groups}

    # ---- the other contraction modes, briefly (same workload, device-resident, 5 evals) -----
    other = {}
    if a.workload == "c3" and nranks == 1 and not a.no_other_modes:
        for mode in ("fp32", "tf32x3", "tf32", "bf16"):
            if mode == a.gemm:
                continue
            g2 = Y.CompiledGrad(tape, ctx, MODE_FLAGS[mode])
            t = timed(g2, dargs, outs, 5, 3)
            other[mode] = {"evals_per_s": 1e3 / t, "ms_per_step": t, "loss": float(outs[g2.low.val[1]].to_host()),
                           "parity_gate": MODE_GATE[mode]}
            del g2

    # ---- weak scaling (N > 1): 8192 columns per GPU, global batch 8192 N ---------------------
    weak = None
    del gf, outs, dargs
    if nranks > 1 and a.workload == "c3" and not a.no_weak and not a.small:
        if symm_base is not None:
            ffi.check(lib.yota_comm_symm_free(ctx.h, symm_base))
        wlw = workload("c3", rank, nranks, a.small, weak=True)
        dw, _tw, gw, ow, _ai, wbase = setup(wlw)
        t = timed(gw, dw, ow, max(5, a.steps // 2), 3)
        weak = {"global_batch": wlw["global_batch"], "batch_per_gpu": wlw["args"][-1].shape[1], "ms_per_step": t,
                "evals_per_s": 1e3 / t, "samples_per_s": wlw["global_batch"] * 1e3 / t,
                "batch_8192_equivalents_per_s": nranks * 1e3 / t, "loss": float(ow[gw.low.val[1]].to_host())}
        del gw, ow, dw
        if wbase is not None:
            ffi.check(lib.yota_comm_symm_free(ctx.h, wbase))

    # ---- the metric's second half: pullback HBM GB/s (C2 region, C5 scatter-add), N = 1 ------
    secondary = None
    if nranks == 1 and a.workload == "c3" and not a.no_secondary:
        secondary = {}
        for name, k in (("c2", 20), ("c5", 5)):
            try:
                secondary[name] = hbm_leg(Y, ffi, ctx, lib, name, k, pk, a.small)
            except Exception as e:  # noqa: BLE001
                secondary[name] = {"error": str(e)[:300]}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample -----------------
    cpu = None
    if rank == 0 and nranks == 1 and not a.no_cpu_baseline:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", a.workload,
                            "--steps", "2", "--warmup", "1"] + (["--small"] if a.small else []),
                           capture_output=True, text=True, env=dict(os.environ, YOTA_REF_BUDGET_S="30"))
        try:
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception:
            cpu = {"value": None, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + r.stderr[-200:]}

    if rank == 0:
        par = f"dp{nranks} (batch columns sharded, all-reduce of parameter grads"
        par += f": {ar_impl}; the last reductions of eval i overlap the forward sweep of eval i+1)" if ar_impl else ")"
        line = {"metric": "Yota.grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": nranks, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": MODE_DTYPE[a.gemm],
                "data": "synthetic",
                "config": {"workload": wl["label"], "gemm_mode": a.gemm, "parity_gate": MODE_GATE[a.gemm],
                           "global_batch": wl.get("global_batch"), "parallelism": par,
                           "l2": "working set (>=1 GiB of weights/activations per eval) is far larger than the 126 MB L2; no flush needed",
                           "loss": float(loss)},
                "clocks": clocks,
                "e2e": {"value": 1e3 / e2e_ms, "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4,
                        "loss": e2e_loss,
                        "note": "per step: batch (x, Y) uploaded from pinned host memory on the copy stream (double-buffered, overlapping the previous sweep), loss copied back to pinned host memory every step (asynchronous copy ordered after the step, synchronised inside the timed region); parameters and gradients stay on the device"},
                "gpu_launches": int(stats["launches"]) * a.steps,
                "program": stats,
                "roofline": roof, "cpu_baseline": cpu, "other_gemm_modes": other, "secondary": secondary,
                "weak_scaling": weak,
                "step_breakdown_ms": {k: round(v[0], 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])}}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
