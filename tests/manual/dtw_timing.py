"""Time of the device DTW (features.dtw_alignment, drop-in for 01_make_dict_parallel.dtw_alignment) on a set of
parallel utterances of the reference's size (20 file pairs, 600-800 frames of 513-dim features each; its log: 70 s for
20 files with a multiprocessing pool), next to the restated recursion on one pair (pure-Python loops, like the package)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from exemplars_vc_b200 import features  # noqa: E402
from oracle import dtw_oracle as o  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    n_files, dim = 20, 513
    A = [rng.random((dim, int(rng.integers(600, 800)))) for _ in range(n_files)]
    B = []
    for a in A:  # a time-warped, noisy copy of every file
        c = int(rng.integers(600, 800))
        idx = np.clip(np.round(np.linspace(0, a.shape[1] - 1, c) + rng.normal(0, 2, c)).astype(int), 0, a.shape[1] - 1)
        B.append(np.ascontiguousarray(a[:, np.sort(idx)] + 0.05 * rng.random((dim, c))))
    features.dtw_alignment(A[:2], B[:2])  # warm-up (module load, allocations)
    t0 = time.perf_counter()
    paths = features.dtw_alignment(A, B)[0]
    t_gpu = time.perf_counter() - t0
    cells = sum(a.shape[1] * b.shape[1] for a, b in zip(A, B))
    print(f"device DTW: {n_files} file pairs, {cells / 1e6:.1f} M cells, dim {dim}: {t_gpu * 1e3:.1f} ms "
          f"(host features in, host paths out); path lengths {len(paths[0][0])}..")
    a, b = A[0][:, :200], B[0][:, :200]
    t0 = time.perf_counter()
    dist, _, _, (p, q) = o.dtw(a.T, b.T)
    t_cpu = time.perf_counter() - t0
    (pg, qg), = features.dtw_alignment([a], [b])[0]
    assert np.array_equal(pg, p) and np.array_equal(qg, q)
    print(f"restated recursion (oracle, Python loops) on one 200 x 200 pair: {t_cpu:.2f} s -> "
          f"{t_cpu / 4e4 * cells:.0f} s extrapolated to the {cells / 1e6:.1f} M cells above; paths identical")


if __name__ == "__main__":
    main()
