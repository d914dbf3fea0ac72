"""CPU restatement of the DTW the reference runs to build its exemplar index paths.  TEST INFRASTRUCTURE ONLY.

Call site: ``dist, cost, cum_cost, path = dtw(feat_A.T, feat_B.T, lambda x, y: sum(np.square(x - y)))``
(01_make_dict_parallel.py:226; 01_make_dict.py has the same call).  ``dtw`` is the un-vendored, un-pinned PyPI
package of that name (pierre-rouanet/dtw; the reference has no requirements file -- its 2018/19 vintage is the 1.3.x
series) and is NOT installed in this image, so this file restates the package's published algorithm
(``dtw/dtw.py``: ``dtw`` with ``warp=1`` and ``_traceback``):

    D0 = zeros((r+1, c+1)); D0[0, 1:] = inf; D0[1:, 0] = inf; D1 = D0[1:, 1:]
    D1[i, j] = dist(x[i], y[j])
    D1[i, j] += min(D0[i, j], D0[i+1, j], D0[i, j+1])          # diagonal, left, up
    path = _traceback(D0): from (r-1, c-1), tb = argmin((D0[i, j], D0[i, j+1], D0[i+1, j])):
           0 -> (i-1, j-1), 1 -> i-1, 2 -> j-1, until i == j == 0;  returns (array(p), array(q))
    return D1[-1, -1] / sum(D1.shape), C, D1, path

PARITY UNPINNED against the package itself (it cannot be run here and the reference holds no golden paths); the
restatement is anchored on the reference's call site and local cost, and checked for the properties any DTW path
has (tests/test_oracle.py).  `local_cost` reproduces the lambda's left-to-right float64 summation.
"""
import numpy as np


def local_cost(x, y):
    """C[i, j] = sum(np.square(x[i] - y[j])) with python's left-to-right sum (01_make_dict_parallel.py:226)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    C = np.zeros((len(x), len(y)))
    for d in range(x.shape[1]):
        e = x[:, d, None] - y[None, :, d]
        C = C + e * e
    return C


def dtw(x, y):
    """x (r, dim), y (c, dim) -> (dist, C, D1, (p, q)) as the package returns them."""
    C = local_cost(x, y)
    r, c = C.shape
    D0 = np.zeros((r + 1, c + 1))
    D0[0, 1:] = np.inf
    D0[1:, 0] = np.inf
    D1 = D0[1:, 1:]
    D1[:, :] = C
    for i in range(r):
        for j in range(c):
            D1[i, j] += min(D0[i, j], D0[i + 1, j], D0[i, j + 1])
    i, j = r - 1, c - 1
    p, q = [i], [j]
    while i > 0 or j > 0:
        tb = int(np.argmin((D0[i, j], D0[i, j + 1], D0[i + 1, j])))
        if tb == 0:
            i -= 1; j -= 1
        elif tb == 1:
            i -= 1
        else:
            j -= 1
        p.insert(0, i); q.insert(0, j)
    return D1[-1, -1] / (r + c), C, D1.copy(), (np.array(p), np.array(q))
