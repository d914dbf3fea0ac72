#!/usr/bin/env python
"""Headline benchmark: converted frames/sec at 20k-exemplar KL-NMF (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode 3xtf32|tf32|bf16|fp32]
                    [--workload single_utterance_20k|batch_256utt_20k|large_dictionary_200k|context_stacked_50k|
                                reference_default]

One "step" = one pass of the hot path over one batch of synthetic input: 500 KL multiplicative-update
iterations of the activation solve over the resident dictionary pair, then the conversion product Y = B H.
N = 1 runs BASELINE.json configs[1] (F=513, N=20000, T=1000).  N > 1 (torchrun, one rank per GPU) shards
whole utterances: every rank converts its own T=1000 utterance against its replica of the dictionary, no
data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-timed with the inputs resident in HBM; `e2e` is the same
metric through the public Python API with host buffers (H2D of the frames and D2H of H and Y inside the timed
region; `e2e.copy_ms` is the copy leg timed on its own).  `extra.exemplar_sharded` is BASELINE configs[3]
(F=513, N=200000, T=2000) run at the SAME N with the exemplar dimension sharded and a fixed iteration count:
ms per iteration, the exchange's share, and an in-run parity check of the objective against the stored 1-GPU
value -- strong scaling of that line over N is the exemplar-sharded curve.

Other workloads (one JSON line each, same keys): `--workload batch_256utt_20k` (configs[2], utterance-sharded
at N > 1), `--workload large_dictionary_200k` (configs[3] alone, all 500 iterations), `--workload
context_stacked_50k` (configs[4]), `--workload reference_default` (the reference's real call: T=688, N=20727,
max_iter=150, tol=1e-4, through the script-level drop-in, dictionary upload inside the timed region).

`--impl reference` times the reference's own CPU implementation (its exact scikit-learn call,
04_align_n_nmf.py:212-213 with KL, plus np.matmul for Y) on the host cores with every BLAS thread the process
may use (`blas_threads` in the line), on a bounded sample of the same workload.
"""
from __future__ import annotations

import os
import sys


def _argv(flag, default):
    return sys.argv[sys.argv.index(flag) + 1] if flag in sys.argv[:-1] else default


def _usable_cpus() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; OpenBLAS reads it when numpy is imported, which
# throttled the reference arm 7x at N > 1 in round 1.  The CPU legs get every core the process may use, set here,
# before numpy loads (and reported as `blas_threads`).
if _argv("--impl", "ours") == "reference" or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_usable_cpus())

import argparse  # noqa: E402
import gc  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converted frames/sec at 20k-exemplar KL-NMF"
UNIT = "frames/s"
SHARDED_PARITY = os.path.join(ROOT, "tests", "golden", "large_dictionary_200k_objective.json")


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


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

    def start(self, wait_first_sample_s: float = 3.0):
        """Starts the sampler and waits until its FIRST sample is on disk: loading NVML and the first query stall the
        driver for tens of milliseconds, which must not land in a timed region (the periodic samples after it do not)."""
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < wait_first_sample_s:
            try:
                if os.path.getsize(self.f.name) > 0:
                    break
            except OSError:
                break
            time.sleep(0.02)

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


def reference_default_seconds(X, A, B, wl):
    """The reference's real call, un-extrapolated: max_iter = 150, tol = 1e-4 (it stops by its own rule), + Y."""
    from oracle import nmf_oracle as o
    t = time.perf_counter()
    W, n_iter = o.reference_call(X, A, tol=wl.tol, max_iter=wl.iterations)
    o.convert(W, B)
    return time.perf_counter() - t, n_iter


def run_reference(args, wl, X, A, B):
    cores = _usable_cpus()
    secs, note = [], {}
    if wl.name == "reference_default":
        for _ in range(args.steps):
            s, n_iter = reference_default_seconds(X, A, B, wl)
            secs.append(s)
        note = {"n_iter": int(n_iter), "extrapolated": False}
        sample = (f"the reference's whole call (float32, solver='mu', update_H=False, max_iter={wl.iterations}, "
                  f"tol={wl.tol}: stopped after {n_iter} iterations) + np.matmul for Y, not extrapolated")
    else:
        for _ in range(min(args.warmup, 1)):
            reference_step_seconds(X, A, B, wl.iterations, 1, np.float32)
        for _ in range(args.steps):
            s, per_iter = reference_step_seconds(X, A, B, wl.iterations, args.ref_sample_iters, np.float32)
            secs.append(s)
        note = {"extrapolated": True}
        sample = (f"{args.ref_sample_iters} of {wl.iterations} KL iterations of sklearn non_negative_factorization "
                  f"(float32, solver='mu', update_H=False) at the full shape, extrapolated linearly, + np.matmul for Y")
    sec = float(np.mean(secs))
    value = wl.T / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": wl.name, "F": wl.F, "N": wl.N, "T": wl.T, "iterations": wl.iterations}, **note),
            "blas_threads": blas_threads(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def measured_matmul_tflops(torch, dev, dtype):
    """Dense library GEMM rate in this run (MEASURED_PEAKS.json has no TF32 figure): torch.matmul 8192^3."""
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


def exemplar_sharded_extra(args, torch, dist, dev, rank, world):
    """BASELINE configs[3] at this N: F=513, N=200000, T=2000, exemplar dimension sharded over the ranks (one
    all-reduce of the partial A*H per iteration), `--sharded-iterations` iterations + Y.  Returns the dict that
    goes under extra.exemplar_sharded (rank 0) -- ms per iteration is device-timed, max over ranks."""
    from exemplars_vc_b200 import ExemplarDictionary, sharding, synth
    wl = synth.CONFIGS["large_dictionary_200k"]
    iters = args.sharded_iterations
    seed = synth.BASE_SEED + 7
    # every rank draws the same stream and keeps its rows, so the dictionary equals the 1-GPU one
    A_full, B_full = synth.dictionaries(seed, wl.F, wl.N)
    X = synth.frames(seed, A_full, wl.T)
    if world > 1:
        d = sharding.make_exemplar_sharded(lambda a, b: A_full[a:b], lambda a, b: B_full[a:b], wl.N, mode=args.mode,
                                           p2p=not args.no_p2p, max_frames=wl.T)
    else:
        d = ExemplarDictionary(A_full, B_full, mode=args.mode)
    del A_full, B_full
    x_dev = torch.from_numpy(X).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d.solve(x_dev, tol=0.0, max_iter=5)
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    act = d.solve(x_dev, tol=0.0, max_iter=iters)
    e1.record()
    d.convert(act.H)
    e2.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev, dtype=torch.float64)
    d.profile(True)
    d.solve(x_dev, tol=0.0, max_iter=10)
    prof = d.profile_read()
    d.profile(False)
    ex = torch.tensor([prof["exchange"][0] / max(prof["exchange"][1], 1) * 1e3,
                       sum(v[0] for v in prof.values())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ex, op=dist.ReduceOp.MAX)
    out = {"workload": wl.name, "F": wl.F, "N": wl.N, "T": wl.T, "iterations": iters, "mode": args.mode,
           "n_gpus": world, "all_reduce": getattr(d, "all_reduce", None) if world > 1 else None,
           "ms_per_iteration": float(t[0].item()) / iters, "convert_ms": float(t[1].item()),
           "frames_per_s_at_500_iterations": wl.T / ((float(t[0].item()) / iters * 500 + float(t[1].item())) * 1e-3),
           "exchange_us": float(ex[0].item()) if world > 1 else 0.0,
           "exchange_share_of_profiled_step": (prof["exchange"][0] / max(sum(v[0] for v in prof.values()), 1e-9))
           if world > 1 else 0.0,
           "objective": act.objective}
    expected = None
    if os.path.exists(SHARDED_PARITY):
        try:
            expected = json.load(open(SHARDED_PARITY)).get(args.mode, {}).get(str(iters))
        except Exception:
            expected = None
    if expected:
        rel = abs(act.objective - expected) / expected
        # (the shards accumulate K ranges of different lengths in TMEM, so the sums differ in the last bits: measured
        # 7e-6 at 2 GPUs after 50 iterations; a wrong exchange is off by orders of magnitude more)
        out["parity"] = {"objective_1gpu": expected, "rel_diff": rel, "tolerance": 5e-5, "ok": bool(rel < 5e-5)}
        if rel >= 5e-5 and rank == 0:       # reported in the line (parity.ok = false), not fatal for the headline number
            print(f"bench.py: exemplar-sharded objective {act.objective} differs from the 1-GPU value {expected} by "
                  f"{rel:.2e}", file=sys.stderr)
    else:
        out["parity"] = {"objective_1gpu": None, "note": "no stored 1-GPU value for this mode / iteration count"}
    d.close()
    return out


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
    ap.add_argument("--sharded-iterations", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-float64", action="store_true", help="skip the float64 leg of the CPU baseline")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.exemplar_sharded (configs[3] at this N)")
    ap.add_argument("--no-p2p", action="store_true", help="exemplar sharding: ncclAllReduce instead of the peer-memory kernel")
    args = ap.parse_args()

    from exemplars_vc_b200 import synth
    wl = synth.CONFIGS[args.workload]
    if args.iterations:
        wl = synth.Workload(wl.name, wl.F, wl.N, wl.T, args.iterations, wl.n_utt, wl.description, wl.tol)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    seed = synth.BASE_SEED + 1

    if args.impl == "reference":
        if rank != 0:
            return
        A, B = synth.dictionaries(seed, wl.F, wl.N)
        T_ref = wl.T
        if wl.n_utt > 1:
            # per-frame cost is constant and a 130k-frame float32 call needs tens of GB of host memory: the CPU arm
            # times the first utterance and reports frames/s on it
            T_ref = int(synth.utterance_lengths(synth.BASE_SEED + 2, wl.n_utt)[0])
            wl = synth.Workload(wl.name, wl.F, wl.N, T_ref, wl.iterations, wl.n_utt, wl.description, wl.tol)
        X = synth.frames(seed, A, T_ref)
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
    batch = wl.n_utt > 1
    script_level = wl.name == "reference_default"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs (host), dictionary upload (not timed: it stays resident across utterances)
    offs = None
    if exemplar_sharded:
        n0, n1 = sharding.exemplar_range(wl.N, rank, world)
        # every rank draws the same stream and keeps its rows, so the dictionary equals the 1-GPU one
        A, B = synth.dictionaries(seed, wl.F, wl.N)
        X_host = synth.frames(seed, A, wl.T)
        d = sharding.make_exemplar_sharded(lambda a, b: A[a:b], lambda a, b: B[a:b], wl.N, mode=args.mode,
                                           p2p=not args.no_p2p, max_frames=wl.T)
    else:
        A, B = synth.dictionaries(seed, wl.F, wl.N)
        if batch:
            # configs[2]: the 256 utterances are dealt to the ranks by frame count (sharding.partition_utterances),
            # each rank stacks its own along T and solves them in one batched call
            lens = synth.utterance_lengths(synth.BASE_SEED + 2, wl.n_utt)
            mine = sharding.partition_utterances(lens, world)[rank]
            offs = np.concatenate([[0], np.cumsum([lens[i] for i in mine])]).astype(int).tolist()
            X_host = synth.frames(seed + 100 * rank, A, offs[-1])
        else:
            X_host = synth.frames(seed + 10 * rank, A, wl.T)     # each rank converts its own utterance
        d = None if script_level else ExemplarDictionary(A, B, mode=args.mode)
    x_pinned = torch.from_numpy(X_host).pin_memory()
    x_dev = x_pinned.to(dev)
    kw = dict(tol=wl.tol, max_iter=wl.iterations)
    frames_local = int(X_host.shape[0])
    n_iter_seen = [wl.iterations]

    if script_level:
        # the drop-in a reference user calls: numpy in, numpy out, dictionary handling inside
        from exemplars_vc_b200 import align_n_nmf
        align_n_nmf.beta_override = "kullback-leibler"
        align_n_nmf.mode = args.mode
        align_n_nmf.max_iter = wl.iterations

        def step_e2e():
            H_nt = align_n_nmf._factorize(X_host, A, tol=wl.tol)
            y = align_n_nmf._product(H_nt, B)
            return None, y, H_nt

        step_resident = None
    else:
        def solve(x):
            if batch:
                acts = d.solve_batched(x, offs, per_utterance_stop=True, **kw)
                return acts[0], acts[0].H_stacked
            act = d.solve(x, **kw)
            n_iter_seen[0] = act.n_iter
            return act, act.H

        def step_resident():
            act, H = solve(x_dev)
            y = d.convert(H)
            return act, H, y

        def step_e2e():
            x = x_pinned.to(dev, non_blocking=True)                 # H2D of this step's frames
            act, H = solve(x)
            y = d.convert(H)
            yh = d.to_host(y, key="Y")                              # D2H of the converted frames
            # D2H of the activations (the reference returns them); the 10 GB stack of the batch workload stays on
            # the device -- the product of a batch conversion is Y
            hh = d.to_host(H, key="H") if not batch else np.empty((0,), np.float32)
            return act, yh, hh

    launches = 0
    enqueue_ms = 0.0
    clocks = None
    act = H_last = y = None
    # a full collection of a torch process's heap takes tens of milliseconds: none inside the timed regions (the
    # wall-clocked end-to-end loop showed single steps twice as long as their neighbours)
    gc.collect()
    gc.disable()
    if step_resident is not None:
        # the sampler is started BEFORE the warm-up: spawning nvidia-smi and its first NVML queries stall the driver
        # for tens of milliseconds (seen as one timed step twice as long as its neighbours), which belongs in the
        # warm-up, not in the timed region; it keeps sampling every 200 ms through the timed steps
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            act, H_last, y = step_resident()      # (held like the timed steps' results: same allocation pattern)
        barrier()
        launches0 = _lib.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            act, H_last, y = step_resident()
            enqueue_ms += float(_lib.lib().evc_last_enqueue_ms())
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = _lib.kernel_launch_count() - launches0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)

    # ---- end to end through the public API, host buffers
    # warm-up of the end-to-end loop: at least two steps whose results are HELD exactly like the timed ones, so that
    # the second buffer each step needs while the previous step's result is still referenced (80 MB of H on the
    # device, a page-locked host block for the script-level path) is allocated here and not in the timed region
    # (seen as a second timed step 60 ms longer than its neighbours)
    act_e = yh = hh = None
    for _ in range(2 if step_resident is not None else max(args.warmup, 2)):
        act_e, yh, hh = step_e2e()
    barrier()
    if step_resident is None:
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.kernel_launch_count()
    t0 = time.perf_counter()
    e2e_each, e2e_enqueue = [], 0.0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        act_e, yh, hh = step_e2e()
        e2e_each.append(round((time.perf_counter() - t1) * 1e3, 3))
        e2e_enqueue += float(_lib.lib().evc_last_enqueue_ms())
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if step_resident is None:
        clocks = sampler.stop() if rank == 0 else None
        launches = _lib.kernel_launch_count() - launches0
        ms = e2e_ms.clone()

    gc.enable()
    # ---- the copy leg of the end-to-end path on its own (H2D of X, D2H of Y and H), so the gap can be checked
    copy_ms = None
    if step_resident is not None:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            x_pinned.to(dev, non_blocking=True)
            d.to_host(y, key="Y")
            if not batch:
                d.to_host(H_last, key="H")
        torch.cuda.synchronize()
        copy_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    # ---- per-kernel-class device time of one more step (CUDA events around every launch)
    prof = None
    if d is not None:
        H_last = y = None
        d.profile(True)
        step_resident()
        prof = d.profile_read()
        d.profile(False)

    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    e2e_step = float(e2e_ms.item()) / args.steps
    ft = torch.tensor([frames_local], device=dev, dtype=torch.float64)
    if world > 1 and not exemplar_sharded:
        dist.all_reduce(ft, op=dist.ReduceOp.SUM)
    frames_total = float(ft.item())
    value = frames_total / (ms_step * 1e-3)
    e2e_value = frames_total / (e2e_step * 1e-3)

    # ---- configs[3] at this N with the exemplar dimension sharded (all ranks take part)
    extra = {}
    if not args.no_extras and args.workload == "single_utterance_20k" and args.mode != "fp32":
        if d is not None:
            d.close()
            d = None
        del x_dev
        torch.cuda.empty_cache()
        extra["exemplar_sharded"] = exemplar_sharded_extra(args, torch, dist, dev, rank, world)

    if rank == 0:
        peaks = load_peaks()
        try:
            lib_gemm = {"tf32_matmul_tflops": round(measured_matmul_tflops(torch, dev, torch.float32), 1),
                        "bf16_matmul_tflops": round(measured_matmul_tflops(torch, dev, torch.bfloat16), 1)}
        except Exception as e:      # diagnostics only
            lib_gemm = {"error": str(e)[:100]}
        roofline = None
        iters_run = n_iter_seen[0]
        if prof is not None:
            n_local = (wl.N if not exemplar_sharded else (n1 - n0))
            # dominant kernel: contraction 2 (R A^T with the fused multiplicative update); algorithmic work per
            # launch = 2*T*F*N_local flop (4*F*N per frame per iteration, half in each contraction)
            c2_ms, c2_n = prof["contraction2_update"]
            c1_ms, c1_n = prof["contraction1"]
            per_launch_s = (c2_ms / max(c2_n, 1)) * 1e-3
            flop = 2.0 * frames_local * wl.F * n_local
            achieved = flop / per_launch_s / 1e12 if per_launch_s > 0 else 0.0
            passes = max(1, int(_lib.lib().evc_mma_passes_per_product(_lib.MODES[args.mode])))
            bf16_peak = peaks["bf16_tflops_sustained"]
            tf32_peak = bf16_peak / 2.0
            alg_peak = bf16_peak if args.mode == "bf16" else tf32_peak
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp) and args.workload == "single_utterance_20k":
                try:
                    traffic = json.load(open(tp)).get(args.mode, {}).get("contraction2_update")
                except Exception:
                    traffic = None
            exec_peak = tf32_peak if args.mode == "tf32" else bf16_peak
            roofline = {
                "bound": "tensor",
                "kernel": "tc_gemm_kernel<contraction 2 + MU epilogue>" if args.mode != "fp32" else "simt gemm_kernel<MU epilogue>",
                "achieved": achieved, "peak": alg_peak, "unit": "TFLOP/s", "frac": achieved / alg_peak, "traffic": traffic,
                "peak_source": ("dense BF16 = bf16_tflops_sustained, " if args.mode == "bf16" else
                                "dense TF32 = 1/2 of bf16_tflops_sustained, ") + peaks["_source"],
                # the same fraction against the library GEMM rates measured IN THIS RUN (cuBLAS through torch.matmul)
                "frac_vs_library_this_run": (achieved / lib_gemm[("bf16" if args.mode == "bf16" else "tf32") + "_matmul_tflops"])
                if "error" not in lib_gemm else None,
                # what the tensor pipe executes: `passes` MMAs per product -- bf16 (kind::f16) MMAs in the
                # fp32-accurate split mode and in bf16 mode, tf32 MMAs in tf32 mode
                "executed_tflops": achieved * passes, "executed_frac": achieved * passes / exec_peak,
                "executed_peak": "dense TF32" if args.mode == "tf32" else "dense BF16 (bf16_tflops_sustained)",
                "executed_frac_vs_library_this_run": (achieved * passes / lib_gemm[("tf32" if args.mode == "tf32" else "bf16") + "_matmul_tflops"])
                if "error" not in lib_gemm else None,
                "mma_passes_per_product": passes, "library_gemm_this_run": lib_gemm,
                "us_per_launch": per_launch_s * 1e6,
                "step_share": {k: round(v[0] / max(sum(x[0] for x in prof.values()), 1e-9), 4) for k, v in prof.items()},
                "contraction1_us_per_launch": c1_ms / max(c1_n, 1) * 1e3,
                "class_ms_launches": {k: [round(v[0], 3), v[1]] for k, v in prof.items()}}
        total_flop = (4.0 * iters_run + 2.0) * frames_total * wl.F * wl.N
        objective = next((a.objective for a in (act, act_e) if a is not None), None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if exemplar_sharded else "weak", "vs_baseline": None,
                "dtype": {"3xtf32": "bf16x3 split products, fp32 accumulate (fp32-accurate mode)",
                          "tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.mode],
                "data": "synthetic",
                "config": {"workload": wl.name, "F": wl.F, "N": wl.N, "T": int(frames_total) if batch else wl.T,
                           "iterations": wl.iterations, "tol": wl.tol, "iterations_run": iters_run, "n_utt": wl.n_utt,
                           "mode": args.mode,
                           "sharding": ("exemplar" if exemplar_sharded else "utterance") if world > 1 else "none",
                           "all_reduce": getattr(d, "all_reduce", None) if d is not None else None,
                           "value_is": "device-timed, inputs resident" if step_resident is not None else
                                       "script-level drop-in, numpy in / numpy out, dictionary upload inside (same as e2e)",
                           "l2": "working set (H + dictionary operands) exceeds the 126 MB L2; no flush needed"},
                "tflops_algorithmic": total_flop / (ms_step * 1e-3) / 1e12,
                "objective": objective, "host_enqueue_ms_per_step": enqueue_ms / args.steps,
                "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_step, "copy_ms": copy_ms,
                        "ms_each_step": e2e_each, "host_enqueue_ms_per_step": e2e_enqueue / args.steps,
                        "h2d_bytes_per_step": int(x_pinned.numel() * 4),
                        "d2h_bytes_per_step": int(np.asarray(yh).size * 4 + np.asarray(hh).size * 4)}}
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            if script_level:
                secs, n_it = reference_default_seconds(X_host, A, B, wl)
                line["cpu_baseline"] = {"value": wl.T / secs, "unit": UNIT, "cores": _usable_cpus(), "kind": "reference",
                                        "blas_threads": blas_threads(),
                                        "sample": f"the reference's whole call (float32, max_iter={wl.iterations}, tol={wl.tol}: "
                                                  f"{n_it} iterations run) + np.matmul for Y, not extrapolated"}
            else:
                Xs = X_host if not batch else X_host[: offs[1]]
                secs, per_iter = reference_step_seconds(Xs, A, B, wl.iterations, args.ref_sample_iters, np.float32)
                line["cpu_baseline"] = {
                    "value": Xs.shape[0] / secs, "unit": UNIT, "cores": _usable_cpus(), "kind": "reference",
                    "blas_threads": blas_threads(),
                    "sample": f"{args.ref_sample_iters} of {wl.iterations} KL iterations of the reference's sklearn call "
                              f"(float32) on {Xs.shape[0]} frames at the full dictionary, extrapolated linearly, + np.matmul for Y",
                    "s_per_iteration": per_iter}
                if not args.no_cpu_float64:
                    # the reference's WORLD branch is float64 (SURVEY 8d asks for both); a shorter sample keeps the run bounded
                    n64 = max(2, args.ref_sample_iters // 2)
                    secs64, per_iter64 = reference_step_seconds(Xs, A, B, wl.iterations, n64, np.float64)
                    line["cpu_baseline"].update(value_float64=Xs.shape[0] / secs64, s_per_iteration_float64=per_iter64,
                                                sample_float64=f"{n64} of {wl.iterations} iterations, float64")
        print(json.dumps(line))
    if d is not None:
        d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
