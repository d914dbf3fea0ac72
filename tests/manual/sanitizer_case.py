"""Small end-to-end case for compute-sanitizer (memcheck): every tensor-core kernel, ragged shapes, batched stop rule.
compute-sanitizer --tool memcheck python tests/manual/sanitizer_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import ExemplarDictionary  # noqa: E402
from oracle import nmf_oracle as o  # noqa: E402

for (F, N, T) in [(513, 700, 70), (201, 300, 37)]:
    X, A, B = o.gen(5, F, N, T, np.float32)
    for mode in ("3xtf32", "tf32"):
        with ExemplarDictionary(A, B, mode=mode) as d:
            act = d.solve(X, tol=1e-3, max_iter=20)
            y = d.to_host(d.convert(act.H))
            acts = d.solve_batched(X, [0, 10, 10, T], tol=1e-2, max_iter=30, per_utterance_stop=True)
            fro = d.solve(X, beta_loss="frobenius", tol=0.0, max_iter=5)
            print(mode, F, N, T, act.n_iter, float(y.sum()), [a.n_iter for a in acts], fro.objective, flush=True)
print("sanitizer case done")
