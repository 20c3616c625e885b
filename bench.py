#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: `Yota.grad` evals/sec of the 8-layer
4096-wide tanh MLP with softmax cross-entropy at batch 8192 (config C3),
one eval = one warm (val, grads-for-every-argument) call.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1|c4]
                  [--gemm tf32|fp32|bf16] [--impl reference]

Prints ONE JSON line (rank 0).  N > 1: launched by torchrun, one rank per GPU; the global
batch of 8192 is sharded by columns ("strong" scaling), parameter gradients are all-reduced
by NCCL inside the run.  `--impl reference` times the restated reference CPU path (the
oracle: op-by-op NumPy/OpenBLAS, not Julia -- no Julia toolchain exists in this image).
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
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- workloads --------------------------------------------------------------------------
def workload(name, rank, nranks, small=False):
    """-> dict(f, args(host), n_params, label, flops, hbm_bytes, units) for this rank's shard."""
    import cases
    from yota_b200 import dp

    if name == "c3":
        L, W, B = (8, 4096, 8192) if not small else (3, 256, 512)
        args = list(cases.c3_args(L, W, B))
        args[-2], args[-1] = dp.shard_columns(args[-2], rank, nranks), dp.shard_columns(args[-1], rank, nranks)
        Bl = args[-1].shape[1]
        return dict(f=cases.make_c3(L, B), args=args, n_params=2 * L, batch_args=(2 * L, 2 * L + 1),
                    label=f"C3: {L}-layer {W}-wide tanh MLP, softmax-xent, batch {B} (Yota.grad, grads for all args)",
                    gemm_flops=3 * L * 2 * W * W * Bl, gemm_shape=(W, Bl, W), hbm_bytes=None)
    if name == "c2":
        n = 8192 if not small else 1024
        x, Wm, b = cases.c2_args(n)
        x, Wm = dp.shard_columns(x, rank, nranks), dp.shard_columns(Wm, rank, nranks)
        return dict(f=cases.c2_chain, args=[x, Wm, b], n_params=0, batch_args=(0, 1), ar_args=(2,),
                    label=f"C2: sum(tanh.(x .* W .+ b) .^ 2), {n}x{n}", gemm_flops=0, hbm_bytes=4 * x.size * 4)
    if name == "c5":
        rows, nt, ni = (256, 1 << 20, 1 << 22) if not small else (256, 1 << 12, 1 << 14)
        E, idx, V = cases.c5_args(rows, nt, ni)
        return dict(f=cases.c5_embed, args=[E, idx, V], n_params=0, batch_args=(),
                    label=f"C5: sum(E[:, idx] .* V), E {rows}x{nt}, {ni} Int64 indices (bit-exact scatter-add pullback)",
                    gemm_flops=0, hbm_bytes=V.nbytes + idx.nbytes + E.nbytes, hbm_step="ungetindex",
                    # algorithmic bytes of the other steps of the eval (reported beside the scatter-add pullback)
                    # (fused pass: gathers E[:, idx] itself, reads V, writes dV; the loss is a sink)
                    hbm_other={"region": 3 * V.nbytes + idx.nbytes})
    if name == "c1":
        return dict(f=cases.c1_mlp, args=list(cases.c1_args()), n_params=4, batch_args=(4, 5),
                    label="C1: MLP 784-128-10 tanh, batch 64", gemm_flops=3.9e7, hbm_bytes=None)
    if name == "c4":
        H, B, T = (1024, 128, 128) if not small else (128, 32, 8)
        return dict(f=cases.make_c4(T), args=list(cases.c4_args(H, B, T)), n_params=3, batch_args=(),
                    label=f"C4: unrolled Elman RNN H={H} T={T} B={B}", gemm_flops=T * 6 * 2 * H * H * B, hbm_bytes=None)
    raise SystemExit(f"unknown workload {name}")


def run_reference(a):
    """Reference arm: the restated reference CPU path (oracle) on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import yota_oracle as O

    cores = os.cpu_count() or 1
    blas = None
    try:  # BLAS library and thread count actually in use (north_star: state both)
        from threadpoolctl import threadpool_info
        blas = [{"lib": i.get("internal_api"), "threads": i.get("num_threads")} for i in threadpool_info() if i.get("user_api") == "blas"]
    except Exception:
        pass
    wl = workload(a.workload, 0, 1, a.small)
    scale = 1.0
    sample = "full workload"
    if a.workload == "c3" and not a.small:
        # bounded sample: the first 512 of the 8192 batch columns; cost is linear in the batch
        sub = 512
        wl["args"][-2], wl["args"][-1] = wl["args"][-2][:, :sub], wl["args"][-1][:, :sub]
        scale = 8192 / sub
        sample = f"batch columns 1..{sub} of 8192 (time x {scale:g}; all GEMM/elementwise costs are linear in the batch)"
    if a.workload == "c5" and not a.small:
        sub = 1 << 19
        wl["args"][1], wl["args"][2] = wl["args"][1][:sub], wl["args"][2][:, :sub]
        scale = (1 << 22) / sub
        sample = f"first {sub} of 4194304 indices (time x {scale:g})"
    f, args = wl["f"], wl["args"]
    for _ in range(max(1, min(a.warmup, 1))):
        O.grad(f, *args)
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        O.grad(f, *args)
        ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts)) * scale
    val = 1.0 / t
    line = {"impl": "reference", "metric": "Yota.grad evals/sec", "value": val, "unit": "evals/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "note": "restated reference CPU path (NumPy/OpenBLAS oracle), not Julia"},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "blas": blas, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--gemm", default=os.environ.get("YOTA_GEMM_MODE", "tf32"),
                    help="contraction mode: tf32 (default: tcgen05, Float32 storage), bf16 (tcgen05, BF16 shadows), fp32 (strict FFMA)")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other GEMM modes")
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
    wl = workload(a.workload, rank, nranks, a.small)
    f, hargs = wl["f"], wl["args"]
    flags = {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16}[a.gemm]
    dargs = [Y.cu(x) for x in hargs]
    tape = Y.gradtape(f, *dargs)
    gf = Y.CompiledGrad(tape, ctx, flags)
    outs = [Y.B200Array(gf._out_dims[s], "f32", ctx) for s in range(gf.n_out)]
    ar_impl = None
    if nranks > 1:
        dp.init_comm(ctx)
        slots = dp.param_grad_slots(gf, wl["n_params"]) + [gf.low.val[1]]
        for i in wl.get("ar_args", ()):
            slots.append(gf.low.grads[i][1])
        gf.set_allreduce(slots)
        # all-reduced outputs live in symmetric memory: reduced by the library's own NVSwitch kernel
        symm, _symm_base = dp.symmetric_outputs(gf, slots, ctx)
        for s_, arr in symm.items():
            outs[s_] = arr
        ar_impl = "own multimem kernel over NCCL symmetric windows" if symm else "ncclAllReduce"
    stats = gf.stats()

    def ev():
        e = C.c_void_p()
        ffi.check(lib.yota_event_create(ctx.h, C.byref(e)))
        return e

    def elapsed(e0, e1):
        ms = C.c_float()
        ffi.check(lib.yota_event_elapsed_ms(ctx.h, e0, e1, C.byref(ms)))
        return ms.value

    # ---- device-resident timing: K evals, inputs already in HBM -------------------------
    for _ in range(a.warmup):
        gf.run(dargs, out=outs)
    ctx.sync()
    dp.barrier()
    clk = ClockSampler(local)
    clk.start()
    e0, e1 = ev(), ev()
    ctx.sync()
    ffi.check(lib.yota_event_record(ctx.h, e0))
    for _ in range(a.steps):
        gf.run(dargs, out=outs)
    ffi.check(lib.yota_event_record(ctx.h, e1))
    ctx.sync()
    dp.barrier()
    ms = dp.allreduce_host_max(elapsed(e0, e1))
    clocks = clk.stop()
    ms_step = ms / a.steps
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
        C.memmove(p, np.asfortranarray(hargs[i]).ctypes.data, hargs[i].nbytes)
        pinned.append((i, p, hargs[i].nbytes))
        h2d += hargs[i].nbytes
    # the loss of every step is read back into pinned host memory by an asynchronous copy ordered
    # after that step (one slot per step parity), so the host is already enqueueing step i+1
    # while step i runs; the values are checked after the final synchronisation
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
    ctx.sync()
    ffi.check(lib.yota_event_record(ctx.h, e0))
    e2e_steps(a.steps)
    ffi.check(lib.yota_event_record(ctx.h, e1))
    ctx.sync()
    dp.barrier()
    e2e_ms = dp.allreduce_host_max(elapsed(e0, e1)) / a.steps
    for _, p, _nb in pinned:
        lib.yota_host_free(ctx.h, p)
    e2e_loss = float(hloss[(a.steps - 1) & 1])
    lib.yota_host_free(ctx.h, hl)

    # ---- roofline of the dominant kernel: per-step CUDA events (serialised pass) ----------
    pk = peaks()
    prof = None
    for _ in range(3):
        prof = gf.profile(dargs, out=outs)
    groups = {}
    for label, t in prof:
        key = label.split(" ")[0].split("[")[0]
        g = groups.setdefault(key, [0.0, 0])
        g[0] += t
        g[1] += 1
    total_prof = sum(t for _, t in prof)
    dom = max(groups, key=lambda k: groups[k][0])
    roof = None
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        traffic_tab = {}
    if dom == "matmul" and wl["gemm_flops"]:
        n_g = groups[dom][1]
        share = groups[dom][0] / total_prof
        # duration inside the timed loop = the step time measured above x the kernel's share of the
        # step (from the per-step event pass); the event pass itself serialises the steps and runs
        # cooler under the power cap, so its own durations are reported as "isolated"
        avg_ms = ms_step * share / n_g
        ach = wl["gemm_flops"] / n_g / (avg_ms * 1e-3) / 1e12
        peak = pk["tc_sust"] * (0.5 if a.gemm == "tf32" else 1.0)
        roof = {"bound": "tensor", "kernel": f"gemm ({a.gemm})", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic_tab.get(f"gemm ({a.gemm})") if a.workload == "c3" and not a.small else None,
                "avg_launch_ms": avg_ms, "avg_launch_ms_isolated": groups[dom][0] / n_g, "launches_per_step": n_g,
                "share_of_step": share,
                "peak_source": f"{pk['src']} bf16 sustained" + (" x 0.5 (TF32 runs at half the BF16 rate)" if a.gemm == "tf32" else "")
                               + ("; strict-FP32 FFMA kernel, not a tensor-core kernel" if a.gemm == "fp32" else "")}
    elif wl.get("hbm_bytes"):
        dom = wl.get("hbm_step", dom)  # the kernel BASELINE.json names for this workload
        t_iso = groups[dom][0]
        t_dom = ms_step * t_iso / total_prof  # in-loop duration: step time x share (see the GEMM branch)
        ach = wl["hbm_bytes"] / (t_dom * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": traffic_tab.get(f"{dom}:{a.workload}") if not a.small else None, "avg_launch_ms": t_dom / groups[dom][1], "avg_launch_ms_isolated": t_iso / groups[dom][1],
                "launches_per_step": groups[dom][1], "share_of_step": t_iso / total_prof, "peak_source": f"{pk['src']} HBM copy"}
        if wl.get("hbm_other"):
            roof["other_steps"] = {k: {"achieved": b / (ms_step * groups[k][0] / total_prof * 1e-3) / 1e9,
                                       "frac": b / (ms_step * groups[k][0] / total_prof * 1e-3) / 1e9 / pk["hbm"]}
                                   for k, b in wl["hbm_other"].items() if k in groups}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample -----------------
    cpu = None
    if rank == 0 and nranks == 1 and not a.no_cpu_baseline:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", a.workload,
                            "--steps", "2", "--warmup", "1"] + (["--small"] if a.small else []),
                           capture_output=True, text=True)
        try:
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception:
            cpu = {"value": None, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + r.stderr[-200:]}

    # ---- the other contraction modes, briefly (same workload, device-resident, 5 evals) -----
    other = {}
    if a.workload == "c3" and nranks == 1 and not a.no_other_modes:
        for mode in ("fp32", "tf32", "bf16"):
            if mode == a.gemm:
                continue
            g2 = Y.CompiledGrad(tape, ctx, {"fp32": 0, "tf32": ffi.FLAG_GEMM_TF32, "bf16": ffi.FLAG_GEMM_BF16}[mode])
            for _ in range(3):
                g2.run(dargs, out=outs)
            ctx.sync()
            ffi.check(lib.yota_event_record(ctx.h, e0))
            for _ in range(5):
                g2.run(dargs, out=outs)
            ffi.check(lib.yota_event_record(ctx.h, e1))
            ctx.sync()
            t = elapsed(e0, e1) / 5
            other[mode] = {"evals_per_s": 1e3 / t, "ms_per_step": t, "loss": float(outs[g2.low.val[1]].to_host())}
            del g2

    if rank == 0:
        line = {"metric": "Yota.grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": nranks, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32 (f32 storage, f32 accumulate)", "bf16": "bf16 (f32 storage + bf16 operand shadows, f32 accumulate)"}[a.gemm],
                "data": "synthetic",
                "config": {"workload": wl["label"], "gemm_mode": a.gemm, "global_batch": 8192 if a.workload == "c3" and not a.small else None,
                           "parallelism": f"dp{nranks} (batch columns sharded, all-reduce of parameter grads" + (f": {ar_impl})" if ar_impl else ")"),
                           "l2": "working set (>=1 GiB of weights/activations per eval) is far larger than the 126 MB L2; no flush needed",
                           "loss": float(loss)},
                "clocks": clocks,
                "e2e": {"value": 1e3 / e2e_ms, "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4,
                        "loss": e2e_loss,
                        "note": "per step: batch (x, Y) uploaded from pinned host memory on the copy stream (double-buffered, overlapping the previous sweep), loss copied back to pinned host memory every step (asynchronous copy ordered after the step, synchronised inside the timed region); parameters and gradients stay on the device"},
                "gpu_launches": int(stats["launches"]) * a.steps,
                "program": stats,
                "roofline": roof, "cpu_baseline": cpu, "other_gemm_modes": other,
                "step_breakdown_ms": {k: round(v[0], 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])}}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
