"""The reference script's literal setting (beta forced to 'frobenius', 04_align_n_nmf.py:210) at the headline shape:
time per iteration and a float64 spot check on a few frames.  python tests/manual/frobenius_fullsize.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import ExemplarDictionary, synth  # noqa: E402
from oracle import nmf_oracle as o  # noqa: E402

F, N, T = 513, 20000, 1000
A, B = synth.dictionaries(synth.BASE_SEED + 1, F, N)
X = synth.frames(synth.BASE_SEED + 1, A, T)
for mode in ("3xtf32", "tf32"):
    with ExemplarDictionary(A, B, mode=mode) as d:
        Xs = X[:4]
        W_ref, _, obj = o.frobenius_mu(Xs.astype(np.float64), A.astype(np.float64), tol=0.0, max_iter=10)
        act = d.solve(Xs, beta_loss="frobenius", tol=0.0, max_iter=10)
        H = d.to_host(act.H).astype(np.float64)
        eh = np.linalg.norm(H - W_ref) / np.linalg.norm(W_ref)
        xd = torch.from_numpy(X).cuda()
        d.solve(xd, beta_loss="frobenius", tol=0.0, max_iter=3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        act = d.solve(xd, beta_loss="frobenius", tol=0.0, max_iter=50)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"frobenius [{mode}]: {dt / 50 * 1e6:.1f} us/iteration at F={F} N={N} T={T}; 4-frame check vs float64: "
              f"H {eh:.1e}, objective {abs(act.objective) :.4f} (spot obj rel {abs(d.objective(Xs, H.astype(np.float32), 'frobenius') - obj) / obj:.1e})", flush=True)
