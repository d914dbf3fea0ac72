"""GPU diagnostics (not a test): per-mode errors of both contractions and of a short solve, printed even
when something is badly off, so one gpurun call tells what to fix.  Usage: python tests/manual/gpu_diag.py"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import ExemplarDictionary, synth  # noqa: E402
from oracle import nmf_oracle as o  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def contraction_errors(mode, F, N, T, seed=1):
    rng = np.random.default_rng(seed)
    A = (rng.standard_normal((N, F)) ** 2 + 1e-3).astype(np.float32)
    B = (rng.standard_normal((N, F)) ** 2 + 1e-3).astype(np.float32)
    H = rng.random((T, N)).astype(np.float32)
    with ExemplarDictionary(A, B, mode=mode) as d:
        y = d.to_host(d.convert(H))
        wh = d.to_host(d.reconstruct(H))
        X = (H.astype(np.float64) @ A.astype(np.float64) * (0.5 + rng.random((T, F)))).astype(np.float32)
        H0 = rng.random((T, N)).astype(np.float32) + 0.1
        act = d.solve(X, tol=0.0, max_iter=1, H0=H0)
        h1 = d.to_host(act.H)
    yr = H.astype(np.float64) @ B.astype(np.float64)
    whr = H.astype(np.float64) @ A.astype(np.float64)
    WH0 = np.maximum(H0.astype(np.float64) @ A.astype(np.float64), o.EPSILON)
    h1r = H0 * (((X / WH0) @ A.T.astype(np.float64)) / A.astype(np.float64).sum(1))
    e = (rel(y, yr), rel(wh, whr), rel(h1, h1r))
    print(f"  [{mode:6s}] F={F:5d} N={N:6d} T={T:5d}: convert {e[0]:.2e}  reconstruct {e[1]:.2e}  one MU step {e[2]:.2e}",
          flush=True)
    if e[2] > 1e-2:
        bad = np.argwhere(np.abs(h1 - h1r) > 1e-2 * np.abs(h1r).max())
        print("    worst rows/cols of the MU step:", bad[:6].tolist(), "count", len(bad))
    if e[1] > 1e-2:
        bad = np.argwhere(np.abs(wh - whr) > 1e-2 * np.abs(whr).max())
        print("    bad reconstruct entries (t,f):", bad[:6].tolist(), "count", len(bad),
              "sample", wh[0, :4].tolist(), whr[0, :4].tolist())
    return e


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    shapes = [(13, 32, 8), (64, 128, 32), (201, 777, 37), (513, 2000, 64), (513, 20000, 1000), (600, 4100, 300)]
    for mode in ("fp32", "tf32", "bf16", "3xtf32"):
        for (F, N, T) in shapes:
            if mode == "fp32" and N * T > 4e6:
                continue
            try:
                contraction_errors(mode, F, N, T)
            except Exception:
                traceback.print_exc()
                print(f"  [{mode}] F={F} N={N} T={T}: FAILED", flush=True)
    # timing of the headline config, per mode
    A, B = synth.dictionaries(synth.BASE_SEED + 1, 513, 20000)
    X = synth.frames(synth.BASE_SEED + 1, A, 1000)
    for mode in ("tf32", "bf16", "3xtf32"):
        try:
            with ExemplarDictionary(A, B, mode=mode) as d:
                d.solve(X, tol=0.0, max_iter=5)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                act = d.solve(X, tol=0.0, max_iter=100)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                fl = 4.0 * 513 * 20000 * 1000 * 100
                print(f"  [{mode}] 100 iterations: {dt * 1e3:.1f} ms  ({dt * 1e4:.1f} us/iter, {fl / dt / 1e12:.1f} algorithmic TFLOP/s) "
                      f"objective {act.objective:.6f}", flush=True)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
