"""Seeded synthetic workloads of the five BASELINE.json configs.

chi^2-like strictly positive dictionaries, frames that are 5-sparse combinations of source exemplars plus
a small uniform floor -- the generator the survey used for its known-answer table, vectorised so the big
configs build in seconds.  (tests/ check that the small cases equal oracle.nmf_oracle.gen bit for bit.)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BASE_SEED = 20190123


@dataclass(frozen=True)
class Workload:
    name: str
    F: int
    N: int
    T: int
    iterations: int
    n_utt: int = 1
    description: str = ""
    tol: float = 0.0          # 0: run all `iterations` (the bench configs); > 0: the reference's stop rule


def utterance_lengths(seed: int, n_utt: int, lo: int = 400, hi: int = 600):
    return np.random.default_rng(seed + 4).integers(lo, hi + 1, size=n_utt).astype(np.int64)


CONFIGS = {
    # configs[1] of BASELINE.json: the configuration the metric is quoted on
    "single_utterance_20k": Workload("single_utterance_20k", 513, 20000, 1000, 500, 1,
                                     "synthetic single utterance F=513 N=20000 T=1000, 500 KL-MU iterations"),
    # configs[2]: 256 utterances, T ~ U[400,600]
    "batch_256utt_20k": Workload("batch_256utt_20k", 513, 20000, int(utterance_lengths(BASE_SEED + 2, 256).sum()),
                                 500, 256,
                                 "256 synthetic utterances (T~U[400,600]) against one shared 513x20k dictionary pair"),
    # configs[3]
    "large_dictionary_200k": Workload("large_dictionary_200k", 513, 200000, 2000, 500, 1,
                                      "F=513 N=200000 T=2000, exemplar-sharded"),
    # configs[4]
    "context_stacked_50k": Workload("context_stacked_50k", 2565, 50000, 1000, 500, 1,
                                    "+-2-frame stacked exemplars F=2565 N=50000 T=1000"),
    # configs[0] at the size the reference's own logs show (20 files -> N ~ 20.7k exemplars,
    # utterance 100162 -> T = 688 frames) with the reference's settings: max_iter = 150, tol = 1e-4
    # (04_align_n_nmf.py:194, 213), called through the script-level drop-in with the dictionary upload included
    "reference_default": Workload("reference_default", 513, 20727, 688, 150, 1,
                                  "the reference's real call: T=688, N=20727, max_iter=150, tol=1e-4", 1e-4),
}


def dictionaries(seed: int, F: int, N: int, dtype=np.float32):
    A = np.random.default_rng(seed).standard_normal((N, F)) ** 2 + 1e-3
    B = np.random.default_rng(seed + 1).standard_normal((N, F)) ** 2 + 1e-3
    return A.astype(dtype), B.astype(dtype)


def frames(seed: int, A: np.ndarray, T: int, dtype=np.float32, exact: bool = False):
    """X (T,F): each frame = 5 random exemplars with U(0,1) weights + 0.01*U(0,1).

    exact=True reproduces oracle.nmf_oracle.gen draw for draw (slow python loop, dense matmul);
    the default draws the same distribution with vectorised index sampling."""
    N, F = A.shape
    r = np.random.default_rng(seed + 2)
    if exact:
        Ht = np.zeros((T, N))
        for t in range(T):
            idx = r.choice(N, 5, replace=False)
            Ht[t, idx] = r.random(5)
        X = Ht @ A.astype(np.float64)
    else:
        idx = r.integers(0, N, size=(T, 5))
        w = r.random((T, 5))
        X = np.einsum("tk,tkf->tf", w, A[idx].astype(np.float64))
    X = X + 0.01 * np.random.default_rng(seed + 3).random((T, F))
    return X.astype(dtype)
