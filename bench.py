#!/usr/bin/env python
"""Headline benchmark: converted frames/sec at 20k-exemplar KL-NMF (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode 3xtf32|tf32|fp32]
                    [--workload single_utterance_20k|large_dictionary_200k|context_stacked_50k]

One "step" = one pass of the hot path over one batch of synthetic input: 500 KL multiplicative-update
iterations of the activation solve over the resident dictionary pair, then the conversion product Y = B H.
N = 1 runs BASELINE.json configs[1] (F=513, N=20000, T=1000).  N > 1 (torchrun, one rank per GPU) shards
whole utterances: every rank converts its own T=1000 utterance against its replica of the dictionary, no
data-path collective (weak scaling); `--workload large_dictionary_200k` instead shards the exemplar
dimension with a per-iteration NCCL all-reduce of the partial A*H (total work fixed: strong scaling).

Prints ONE JSON line (rank 0).  `value` is device-timed with the inputs resident in HBM; `e2e` is the same
metric through the public Python API with host buffers (H2D of the frames and D2H of H and Y inside the timed
region).  `--impl reference` times the reference's own CPU implementation (its exact scikit-learn call,
04_align_n_nmf.py:212-213 with KL, plus np.matmul for Y) on the host cores, on a bounded sample of the same
workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converted frames/sec at 20k-exemplar KL-NMF"
UNIT = "frames/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "_source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm
def reference_step_seconds(X, A, B, iterations, sample_iters, dtype):
    """Seconds one full step would take on the CPU, from a bounded sample: the reference's exact call with
    max_iter = 1 and max_iter = 1 + sample_iters (difference = sample_iters iterations, best of 2 each after a
    warm-up call so thread-pool start-up and first-touch page faults are not attributed to the iterations),
    extrapolated linearly to `iterations` (the per-iteration cost is constant), plus the measured Y = W @ B."""
    from oracle import nmf_oracle as o
    Xc, Ac, Bc = X.astype(dtype), A.astype(dtype), B.astype(dtype)

    def timed(k):
        best, W = None, None
        for _ in range(2):
            t = time.perf_counter()
            W, _n = o.reference_call(Xc, Ac, tol=0.0, max_iter=k)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        return best, W

    o.reference_call(Xc, Ac, tol=0.0, max_iter=1)          # warm-up
    t_one, _ = timed(1)
    t_many, W = timed(1 + sample_iters)
    t = time.perf_counter(); o.convert(W, Bc); t_conv = time.perf_counter() - t
    per_iter = max((t_many - t_one) / sample_iters, 1e-9)
    fixed = max(t_one - per_iter, 0.0)          # validation, W0, initial objective
    return fixed + per_iter * iterations + t_conv, per_iter


def run_reference(args, wl, X, A, B):
    cores = os.cpu_count() or 1
    secs = []
    for _ in range(min(args.warmup, 1)):
        reference_step_seconds(X, A, B, wl.iterations, 1, np.float32)
    for _ in range(args.steps):
        s, per_iter = reference_step_seconds(X, A, B, wl.iterations, args.ref_sample_iters, np.float32)
        secs.append(s)
    sec = float(np.mean(secs))
    value = wl.T / sec
    sample = (f"{args.ref_sample_iters} of {wl.iterations} KL iterations of sklearn non_negative_factorization "
              f"(float32, solver='mu', update_H=False) at the full shape, extrapolated linearly, + np.matmul for Y")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "F": wl.F, "N": wl.N, "T": wl.T, "iterations": wl.iterations,
                       "extrapolated": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="3xtf32", choices=["3xtf32", "tf32", "bf16", "fp32"])
    ap.add_argument("--workload", default="single_utterance_20k")
    ap.add_argument("--iterations", type=int, default=None, help="override the workload's iteration count")
    ap.add_argument("--ref-sample-iters", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-float64", action="store_true", help="skip the float64 leg of the CPU baseline")
    ap.add_argument("--no-p2p", action="store_true", help="exemplar sharding: ncclAllReduce instead of the peer-memory kernel")
    args = ap.parse_args()

    from exemplars_vc_b200 import synth
    wl = synth.CONFIGS[args.workload]
    if args.iterations:
        wl = synth.Workload(wl.name, wl.F, wl.N, wl.T, args.iterations, wl.n_utt, wl.description)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    seed = synth.BASE_SEED + 1

    if args.impl == "reference":
        if rank != 0:
            return
        A, B = synth.dictionaries(seed, wl.F, wl.N)
        X = synth.frames(seed, A, wl.T)
        run_reference(args, wl, X, A, B)
        return

    import torch
    import torch.distributed as dist
    from exemplars_vc_b200 import ExemplarDictionary, _lib, sharding

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    exemplar_sharded = world > 1 and args.workload == "large_dictionary_200k"

    # ---- synthetic inputs (host), dictionary upload (not timed: it stays resident across utterances)
    if exemplar_sharded:
        n0, n1 = sharding.exemplar_range(wl.N, rank, world)
        rngA = np.random.default_rng(seed)
        # every rank draws the same stream and keeps its rows, so the dictionary equals the 1-GPU one
        A_full, B_full = synth.dictionaries(seed, wl.F, wl.N)
        X_host = synth.frames(seed, A_full, wl.T)
        d = sharding.make_exemplar_sharded(lambda a, b: A_full[a:b], lambda a, b: B_full[a:b], wl.N, mode=args.mode,
                                           p2p=not args.no_p2p, max_frames=wl.T)
        del A_full, B_full, rngA
    else:
        A, B = synth.dictionaries(seed, wl.F, wl.N)
        X_host = synth.frames(seed + 10 * rank, A, wl.T)     # each rank converts its own utterance
        d = ExemplarDictionary(A, B, mode=args.mode)
    x_pinned = torch.from_numpy(X_host).pin_memory()
    x_dev = x_pinned.to(dev)
    kw = dict(tol=0.0, max_iter=wl.iterations)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        act = d.solve(x_dev, **kw)
        y = d.convert(act.H)
        return act, y

    def step_e2e():
        x = x_pinned.to(dev, non_blocking=True)                 # H2D of this step's frames
        act = d.solve(x, **kw)
        y = d.convert(act.H)
        yh = d.to_host(y, key="Y")                              # D2H of the converted frames
        hh = d.to_host(act.H, key="H")                          # D2H of the activations (the reference returns them)
        return act, yh, hh

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    enqueue_ms = 0.0
    for _ in range(args.steps):
        act, y = step_resident()
        enqueue_ms += float(_lib.lib().evc_last_enqueue_ms())
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.kernel_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)

    # ---- end to end through the public API, host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        act_e, yh, hh = step_e2e()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)

    # ---- per-kernel-class device time of one more step (CUDA events around every launch)
    d.profile(True)
    step_resident()
    prof = d.profile_read()
    d.profile(False)

    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    e2e_step = float(e2e_ms.item()) / args.steps
    frames_total = wl.T if exemplar_sharded else wl.T * world
    value = frames_total / (ms_step * 1e-3)
    e2e_value = frames_total / (e2e_step * 1e-3)

    def measured_matmul_tflops(dtype):
        """Dense library GEMM rate in this run (MEASURED_PEAKS.json has no TF32 figure): torch.matmul
        8192^3, a cross-check of the 1/2 x bf16 denominator, not the denominator itself."""
        n = 8192
        a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            for _ in range(3):
                a @ b
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                a @ b
            t1.record(); torch.cuda.synchronize()
            return 10 * 2.0 * n ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old

    if rank == 0:
        peaks = load_peaks()
        try:
            lib_gemm = {"tf32_matmul_tflops": round(measured_matmul_tflops(torch.float32), 1),
                        "bf16_matmul_tflops": round(measured_matmul_tflops(torch.bfloat16), 1)}
        except Exception as e:      # diagnostics only
            lib_gemm = {"error": str(e)[:100]}
        n_local = (wl.N if not exemplar_sharded else (n1 - n0))
        # dominant kernel: contraction 2 (R A^T with the fused multiplicative update); algorithmic work per
        # launch = 2*T*F*N_local flop (4*F*N per frame per iteration, half in each contraction)
        c2_ms, c2_n = prof["contraction2_update"]
        c1_ms, c1_n = prof["contraction1"]
        per_launch_s = (c2_ms / max(c2_n, 1)) * 1e-3
        flop = 2.0 * wl.T * wl.F * n_local
        achieved = flop / per_launch_s / 1e12 if per_launch_s > 0 else 0.0
        passes = max(1, int(_lib.lib().evc_mma_passes_per_product(_lib.MODES[args.mode])))
        tf32_peak = peaks["bf16_tflops_sustained"] / (1.0 if args.mode == "bf16" else 2.0)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.mode, {}).get("contraction2_update")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel<contraction 2 + MU epilogue>" if args.mode != "fp32"
                    else "simt gemm_kernel<MU epilogue>", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "traffic": traffic,
                    "peak_source": ("dense BF16 = bf16_tflops_sustained, " if args.mode == "bf16" else
                                    "dense TF32 = 1/2 of bf16_tflops_sustained, ") + peaks["_source"],
                    # what the tensor pipe executes: `passes` MMAs per product, bf16 (kind::f16) MMAs in the
                    # fp32-accurate split mode and in bf16 mode, tf32 MMAs in tf32 mode
                    "executed_tflops": achieved * passes,
                    "executed_frac": achieved * passes / (tf32_peak if args.mode == "tf32" else peaks["bf16_tflops_sustained"]),
                    "executed_peak": "dense TF32" if args.mode == "tf32" else "dense BF16 (bf16_tflops_sustained)",
                    "mma_passes_per_product": passes, "library_gemm_this_run": lib_gemm,
                    "us_per_launch": per_launch_s * 1e6,
                    "step_share": {k: round(v[0] / max(sum(x[0] for x in prof.values()), 1e-9), 4) for k, v in prof.items()},
                    "contraction1_us_per_launch": c1_ms / max(c1_n, 1) * 1e3,
                    "class_ms_launches": {k: [round(v[0], 3), v[1]] for k, v in prof.items()}}
        total_flop = (4.0 * wl.iterations + 2.0) * wl.T * wl.F * wl.N * (1 if exemplar_sharded else world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if exemplar_sharded else "weak", "vs_baseline": None,
                "dtype": {"3xtf32": "bf16x3 split products, fp32 accumulate (fp32-accurate mode)",
                          "tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.mode],
                "data": "synthetic",
                "config": {"workload": wl.name, "F": wl.F, "N": wl.N, "T": wl.T, "iterations": wl.iterations,
                           "mode": args.mode, "sharding": ("exemplar" if exemplar_sharded else "utterance") if world > 1 else "none",
                           "all_reduce": getattr(d, "all_reduce", None),
                           "l2": "working set (H 80 MB + dictionary operands > 160 MB) exceeds the 126 MB L2; no flush needed"},
                "tflops_algorithmic": total_flop / (ms_step * 1e-3) / 1e12,
                "objective": act.objective, "host_enqueue_ms_per_step": enqueue_ms / args.steps, "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_step,
                        "h2d_bytes_per_step": int(x_pinned.numel() * 4),
                        "d2h_bytes_per_step": int(yh.size * 4 + hh.size * 4)}}
        if world == 1 and not args.no_cpu_baseline:
            A_h, B_h = (A, B)
            secs, per_iter = reference_step_seconds(X_host, A_h, B_h, wl.iterations, args.ref_sample_iters, np.float32)
            line["cpu_baseline"] = {
                "value": wl.T / secs, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                "sample": f"{args.ref_sample_iters} of {wl.iterations} KL iterations of the reference's sklearn call "
                          f"(float32) at the full shape, extrapolated linearly, + np.matmul for Y",
                "s_per_iteration": per_iter}
            if not args.no_cpu_float64:
                # the reference's WORLD branch is float64 (SURVEY 8d asks for both); a shorter sample keeps the run bounded
                n64 = max(2, args.ref_sample_iters // 2)
                secs64, per_iter64 = reference_step_seconds(X_host, A_h, B_h, wl.iterations, n64, np.float64)
                line["cpu_baseline"].update(value_float64=wl.T / secs64, s_per_iteration_float64=per_iter64,
                                            sample_float64=f"{n64} of {wl.iterations} iterations, float64")
        print(json.dumps(line))
    d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
