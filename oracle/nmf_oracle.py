"""CPU oracle for the activation-estimation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``exemplars_vc_b200``) never does: it has no CPU path at all.

What is restated, and from where (paths relative to the reference checkout; ``sklearn:``
is scikit-learn 1.9.0 ``sklearn/decomposition/_nmf.py``, the un-vendored, un-pinned
dependency that holds the arithmetic -- the reference has no requirements file):

* ``reference_call``        -- the reference's own call, verbatim in its arguments:
                               ``04_align_n_nmf.py:212-213`` (with the beta the caller asks
                               for; the script body hard-codes "frobenius" at ``:210``).
* ``kl_mu``                 -- numpy restatement of ``sklearn:521-626`` (beta=1 branch of
                               ``_multiplicative_update_w``), ``sklearn:726-886`` (loop and
                               stop rule), ``sklearn:1224-1226`` (W0 rule) and
                               ``sklearn:78-182`` (objective).
* ``frobenius_mu``          -- same for the beta=2 branch (``sklearn:535-549``).
* ``kl_objective``          -- ``sklearn:139-154`` + ``:176-180``.
* ``convert``               -- ``04_align_n_nmf.py:391`` (``Y = H.T @ B``).
* ``nmf_tool_euclidean_mu`` -- ``nmf_tool/nmf.py:33-40`` with ``initW=True`` (fixed W).

Parity pinning: the reference holds NO golden vectors or tests for this path
(SURVEY.md section 8c), so the oracle is pinned against *outputs of the reference's call
run in the authoring container* (sklearn 1.9.0, numpy 2.3.5): see
``oracle/make_golden.py`` and ``tests/test_oracle.py``.  ``kl_mu`` is bit-identical to
sklearn at lambda = 0.

Layout follows the reference: frames are rows.  X (T,F), A (N,F), W == H^T (T,N), B (N,F).
"""
from __future__ import annotations

import warnings

import numpy as np

# sklearn:32 -- EPSILON is float32 eps whatever the dtype of the data
EPSILON = np.finfo(np.float32).eps


def gen(seed: int, F: int, N: int, T: int, dtype=np.float64):
    """Synthetic problem of SURVEY.md section 8(c)/(d): chi^2-like dictionaries, 5-sparse frames."""
    A = np.random.default_rng(seed).standard_normal((N, F)) ** 2 + 1e-3
    B = np.random.default_rng(seed + 1).standard_normal((N, F)) ** 2 + 1e-3
    r = np.random.default_rng(seed + 2)
    Ht = np.zeros((T, N))
    for t in range(T):
        idx = r.choice(N, 5, replace=False)
        Ht[t, idx] = r.random(5)
    X = Ht @ A + 0.01 * np.random.default_rng(seed + 3).random((T, F))
    return X.astype(dtype), A.astype(dtype), B.astype(dtype)


def reference_call(X, A, beta_loss="kullback-leibler", tol=1e-4, max_iter=150):
    """The reference's exact operator call (04_align_n_nmf.py:212-213). Returns (W, n_iter)."""
    from sklearn.decomposition import non_negative_factorization

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _W, _H, n_iter = non_negative_factorization(
            X=X, H=A, init="custom", update_H=False, n_components=A.shape[0],
            beta_loss=beta_loss, solver="mu", tol=tol, max_iter=max_iter, verbose=0)
    return _W, n_iter


def initial_activation(X, N):
    """sklearn:1224-1226 -- W0 = sqrt(mean(X) / N) everywhere (computed in X's dtype)."""
    avg = np.sqrt(X.mean() / N)
    return np.full((X.shape[0], N), avg, dtype=X.dtype)


def kl_objective(X, W, A):
    """sqrt(2 * KL(X || W A)) exactly as sklearn:118-154,176-180."""
    WH_data = np.dot(W, A).ravel()
    X_data = X.ravel()
    indices = X_data > EPSILON
    WH_data = WH_data[indices]
    X_data = X_data[indices]
    WH_data[WH_data < EPSILON] = EPSILON
    sum_WH = np.dot(np.sum(W, axis=0), np.sum(A, axis=1))
    div = X_data / WH_data
    res = np.dot(X_data, np.log(div))
    res += sum_WH - X_data.sum()
    res = max(res, 0)
    return np.sqrt(2 * res)


def frobenius_objective(X, W, A):
    """sklearn:112-127 -- ||X - W A||_F."""
    d = X - np.dot(W, A)
    return np.sqrt(np.sum(d * d))


def kl_mu(X, A, lam=0.0, tol=1e-4, max_iter=150, W0=None, sklearn_l1_accumulate=False,
          trace=False):
    """KL multiplicative updates with a fixed dictionary.

    Follows sklearn:554-624 line by line.  ``lam`` is the constant L1 penalty of the north
    star (den = A^T 1 + lam).  ``sklearn_l1_accumulate=True`` reproduces sklearn 1.9.0's
    quirk Q1: ``denominator += l1_reg_W`` is applied in place to a view of
    the cached ``H_sum``, so iteration k sees A^T 1 + k*lam.

    Returns (W, n_iter, objective) or (W, n_iter, objective, [trace]) -- objective is
    sqrt(2 KL) of the returned W.
    """
    N = A.shape[0]
    W = initial_activation(X, N) if W0 is None else np.array(W0, dtype=X.dtype, copy=True)
    err0 = kl_objective(X, W, A)
    prev = err0
    H_sum = np.sum(A, axis=1)
    hist = [err0]
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        WH = np.dot(W, A)                          # sklearn:554
        WH[WH < EPSILON] = EPSILON                 # sklearn:568
        np.divide(X, WH, out=WH)                   # sklearn:571
        numerator = np.dot(WH, A.T)                # sklearn:585
        if sklearn_l1_accumulate:
            denominator = H_sum[np.newaxis, :]     # a view: += below mutates H_sum
            if lam > 0:
                denominator += lam                 # sklearn:611-612
        else:
            denominator = (H_sum + lam)[np.newaxis, :] if lam > 0 else H_sum[np.newaxis, :]
            denominator = denominator.astype(X.dtype, copy=True)
        denominator[denominator == 0] = EPSILON    # sklearn:615
        numerator /= denominator                   # sklearn:617
        W *= numerator                             # sklearn:624
        if tol > 0 and n_iter % 10 == 0:           # sklearn:867-879
            err = kl_objective(X, W, A)
            hist.append(err)
            if (prev - err) / err0 < tol:
                break
            prev = err
    obj = kl_objective(X, W, A)
    if trace:
        return W, n_iter, obj, hist
    return W, n_iter, obj


def frobenius_mu(X, A, lam=0.0, tol=1e-4, max_iter=150, W0=None):
    """Frobenius multiplicative updates with a fixed dictionary, sklearn:535-549, 611-624.

    This is what 04_align_n_nmf.py:210 really runs.  Returns (W, n_iter, ||X - W A||_F).
    """
    N = A.shape[0]
    W = initial_activation(X, N) if W0 is None else np.array(W0, dtype=X.dtype, copy=True)
    err0 = frobenius_objective(X, W, A)
    prev = err0
    XHt = np.dot(X, A.T)
    HHt = np.dot(A, A.T)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        numerator = XHt.copy()
        denominator = np.dot(W, HHt)
        if lam > 0:
            denominator += lam
        denominator[denominator == 0] = EPSILON
        numerator /= denominator
        W *= numerator
        if tol > 0 and n_iter % 10 == 0:
            err = frobenius_objective(X, W, A)
            if (prev - err) / err0 < tol:
                break
            prev = err
    return W, n_iter, frobenius_objective(X, W, A)


def convert(W, B):
    """04_align_n_nmf.py:391 -- converted frames Y (T,F) = H^T B = W B."""
    return np.matmul(W, B)


def nmf_tool_euclidean_mu(V, W, H0, max_iter):
    """nmf_tool/nmf.py:33-40 with initW=True: H <- H * (W^T V) / ((W^T W) H); no epsilon.

    North-star orientation here (as in nmf_tool): V (m,n), W (m,r) fixed, H (r,n).
    TensorFlow draws H0 ~ U(0,1) from its own RNG (not reproducible), so parity for this
    API is "same algorithm from the same H0".
    """
    H = np.array(H0, copy=True)
    Wt = W.T
    WtV = Wt @ V
    WtW = Wt @ W
    for _ in range(max_iter):
        H = H * WtV / (WtW @ H)
    cost = float(np.sum((V - W @ H) ** 2))     # nmf_tool/nmf.py:34
    return H, cost
