"""Run the BASELINE.json configs that are not the bench line at (near) full size on one GPU: a few iterations,
timing per iteration, and a float64 spot check of a handful of frames.  Usage: python tests/manual/config_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import ExemplarDictionary, synth  # noqa: E402
from oracle import nmf_oracle as o  # noqa: E402


def spot_check(d, X, A, B, frames, iters):
    idx = np.linspace(0, X.shape[0] - 1, frames).astype(int)
    Xs = X[idx]
    W_ref, _, obj = o.kl_mu(Xs.astype(np.float64), A.astype(np.float64), tol=0.0, max_iter=iters)
    act = d.solve(Xs, tol=0.0, max_iter=iters)
    H = d.to_host(act.H).astype(np.float64)
    Y = d.to_host(d.convert(act.H)).astype(np.float64)
    Yr = W_ref @ B.astype(np.float64)
    return (np.linalg.norm(H - W_ref) / np.linalg.norm(W_ref), np.linalg.norm(Y - Yr) / np.linalg.norm(Yr),
            abs(act.objective - obj) / obj)


def run(name, F, N, T, mode, iters=20, offsets=None):
    seed = synth.BASE_SEED + 3
    A, B = synth.dictionaries(seed, F, N)
    X = synth.frames(seed, A, T)
    with ExemplarDictionary(A, B, mode=mode) as d:
        eh, ey, eo = spot_check(d, X, A, B, 6, 10)
        xd = torch.from_numpy(X).cuda()
        if offsets is None:
            d.solve(xd, tol=0.0, max_iter=2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if offsets is None:
            act = d.solve(xd, tol=0.0, max_iter=iters)
            obj = act.objective
        else:
            acts = d.solve_batched(xd, offsets, tol=0.0, max_iter=iters, per_utterance_stop=True)
            obj = acts[0].objective
        y = d.convert(act.H if offsets is None else torch.cat([a.H for a in acts[:4]]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    fl = 4.0 * F * N * T * iters
    print(f"{name:28s} [{mode:6s}] F={F} N={N} T={T}: {dt / iters * 1e6:9.1f} us/iter ({fl / dt / 1e12:6.1f} algorithmic TFLOP/s) "
          f"| 6-frame check vs float64: H {eh:.1e} Y {ey:.1e} obj {eo:.1e} | objective {obj:.4f}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["batch", "stacked", "large"]
    if "batch" in which:
        lens = synth.utterance_lengths(synth.BASE_SEED + 2, 256)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(int).tolist()
        run("batch_256utt_20k", 513, 20000, offs[-1], "3xtf32", iters=10, offsets=offs)
    if "stacked" in which:
        run("context_stacked_50k", 2565, 50000, 1000, "3xtf32", iters=10)
        run("context_stacked_50k", 2565, 50000, 1000, "tf32", iters=10)
        run("context_stacked_50k", 2565, 50000, 1000, "bf16", iters=10)
    if "large" in which:
        run("large_dictionary_200k(1gpu)", 513, 200000, 2000, "3xtf32", iters=6)
