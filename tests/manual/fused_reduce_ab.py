"""Digest of a short solve at several shapes, for comparing two builds / environments bit for bit:
    EVC_FUSED_REDUCE=1 python tests/manual/fused_reduce_ab.py      (split-K sum + ratio inside contraction 1 at every
                                                                    shape; without the variable only where K is not split)
    EVC_NO_FUSED_REDUCE=1 python tests/manual/fused_reduce_ab.py   (always the separate reduce_partials_kernel launch)
Both must print identical lines: the in-kernel pass sums the partials in the same order (sklearn _nmf.py:554-571)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from exemplars_vc_b200 import ExemplarDictionary, synth  # noqa: E402

SHAPES = [(513, 20000, 1000, 12), (513, 6821, 688, 12), (257, 3000, 130, 12), (2565, 5000, 300, 6), (513, 20000, 5000, 4),
          (513, 4000, 19000, 3)]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    modes = sys.argv[1:] or ["3xtf32", "tf32", "bf16"]
    for F, N, T, iters in SHAPES:
        A, B = synth.dictionaries(7 + F + N, F, N)
        X = synth.frames(11 + T, A, T)
        for mode in modes:
            with ExemplarDictionary(A, B, mode=mode) as d:
                act = d.solve(X, tol=0.0, max_iter=iters)
                H = d.to_host(act.H)
                Y = d.to_host(d.convert(act.H))
            print(f"{F}x{N}x{T} {mode} H {digest(H)} Y {digest(Y)} obj {act.objective!r}", flush=True)


if __name__ == "__main__":
    main()
