"""Operator-level drop-in for the call the reference makes (04_align_n_nmf.py:212-213):

    _W, _H, n_iter = non_negative_factorization(X=X, H=W, init="custom", update_H=False,
                         n_components=W.shape[0], beta_loss=beta_loss, solver='mu', tol=tol,
                         max_iter=150, verbose=1)

Same name, argument meaning, return triple and error behaviour as
``sklearn.decomposition.non_negative_factorization`` (1.9.0, _nmf.py:900-1126) for the one
configuration on the path -- fixed dictionary (``update_H=False``, ``init='custom'``), multiplicative
updates (``solver='mu'``), beta in {KL, Frobenius}.  The arithmetic runs on the GPU through
libevc_b200 (no CPU path); numpy in, numpy out, output dtype = dtype of X.

Precision note: float64 inputs (the dtype of the reference's WORLD branch, pyworld returns float64) are
COMPUTED in float32 on the device (fp32-accurate split products, fp32 accumulation and update) and the
result is cast back to float64 -- inside the 1e-3 / 1e-4 tolerances the parity tests assert against the
float64 reference, but not a float64 computation: quantities that depend on the sign of tiny differences
(e.g. which entries of ``log(H^T A - X)`` at 04_align_n_nmf.py:292 come out NaN) can differ.
"""
from __future__ import annotations

import time
import warnings
from typing import Optional

import numpy as np

from .dictionary import ExemplarDictionary, dictionary_cache

try:  # so that filters written for the reference keep working when scikit-learn is around
    from sklearn.exceptions import ConvergenceWarning
except Exception:  # pragma: no cover - sklearn is not a dependency of the product

    class ConvergenceWarning(UserWarning):
        """Raised when max_iter is reached with tol > 0 (sklearn _nmf.py:1722-1727)."""


DEFAULT_MODE = "3xtf32"
cache_dictionaries = True      # keep uploaded dictionaries resident between calls (content-checked, see DictionaryCache)


def _check_dictionary_dtype(X: np.ndarray, H: np.ndarray):
    # sklearn _nmf.py:1216-1221
    if H.dtype != X.dtype:
        raise TypeError("H should have the same dtype as X. Got H.dtype = {}.".format(H.dtype))


def non_negative_factorization(X, W=None, H=None, n_components="auto", *, init=None, update_H=True, solver="cd",
                               beta_loss="frobenius", tol=1e-4, max_iter=200, alpha_W=0.0, alpha_H="same",
                               l1_ratio=0.0, random_state=None, verbose=0, shuffle=False,
                               mode: str = DEFAULT_MODE, dictionary: Optional[ExemplarDictionary] = None,
                               sklearn_l1_accumulate: bool = True, _device_result: Optional[dict] = None):
    """Fixed-dictionary NMF activations on the GPU.

    Returns ``(W, H, n_iter)`` with ``W`` (n_samples, n_components) the activations (frames are rows,
    the reference transposes it afterwards, 04_align_n_nmf.py:215) and ``H`` the very object passed in.

    Extra keyword arguments (not in scikit-learn):
      mode        -- arithmetic of the contractions, see ExemplarDictionary.
      dictionary  -- a resident ExemplarDictionary built from ``H``; skips the upload of ``H``.
      sklearn_l1_accumulate -- with ``alpha_W*l1_ratio > 0``: True reproduces scikit-learn 1.9.0, whose
                     denominator grows by l1_reg_W every iteration; False applies
                     the constant penalty of the north-star formula.
    """
    if update_H:
        raise NotImplementedError("update_H=True (learning the dictionary) is not on the exemplar-VC path; "
                                  "the reference always calls with update_H=False")
    if init != "custom":
        raise ValueError("init must be 'custom' when the dictionary H is given (update_H=False)")
    if solver != "mu":
        raise NotImplementedError("solver=%r: only the multiplicative-update solver 'mu' is implemented "
                                  "(see SURVEY.md section 2, footnote to row 2)" % (solver,))
    if beta_loss not in ("kullback-leibler", "frobenius", 1, 2, 1.0, 2.0):
        raise NotImplementedError("beta_loss=%r is not on the path" % (beta_loss,))
    if H is None:
        raise ValueError("H (the exemplar dictionary) is required when update_H=False")
    if max_iter < 0 or tol < 0:
        raise ValueError("max_iter and tol must be non-negative")
    X_in = np.asarray(X)
    if X_in.ndim != 2:
        raise ValueError("Expected 2D array, got array with shape %r" % (X_in.shape,))
    if X_in.dtype not in (np.float64, np.float32):
        X_in = X_in.astype(np.float64)          # sklearn check_array(dtype=[float64, float32])
    H_arr = np.asarray(H)
    # sklearn check_array(force_all_finite=True) runs before the sign check (NaN < 0 is False, so a NaN would
    # otherwise slip through): same exception type and wording
    def _check_finite(name, arr):
        if arr.dtype.kind == "f" and arr.size and not np.isfinite(arr).all():
            if np.isnan(arr).any():
                raise ValueError("Input %s contains NaN." % name)
            raise ValueError("Input %s contains infinity or a value too large for %r." % (name, arr.dtype))

    _check_finite("X", X_in)
    use_cache = dictionary is None and cache_dictionaries
    if not use_cache and dictionary is None:
        _check_finite("H", H_arr)          # (a cached dictionary was scanned when it was built; see DictionaryCache.get)
    if X_in.size and X_in.min() < 0:
        raise ValueError("Negative values in data passed to NMF (input X)")
    _check_dictionary_dtype(X_in, H_arr)
    n_samples, n_features = X_in.shape
    if H_arr.ndim != 2 or H_arr.shape[1] != n_features:
        raise ValueError("Array with wrong second dimension passed to NMF (input H). Expected %d, but got %r."
                         % (n_features, H_arr.shape))
    if n_components not in ("auto", None) and int(n_components) != H_arr.shape[0]:
        raise ValueError("Array with wrong first dimension passed to NMF (input H). Expected %s, but got %d."
                         % (n_components, H_arr.shape[0]))
    if W is not None:
        warnings.warn("When update_H=False, the provided initial W is not used.", RuntimeWarning)  # :1206-1210

    beta = "kullback-leibler" if beta_loss in ("kullback-leibler", 1, 1.0) else "frobenius"
    # sklearn _nmf.py:1255-1257 (l2 term is not on the reference's path)
    l1_reg_W = n_features * alpha_W * l1_ratio
    if alpha_W * (1.0 - l1_ratio) != 0.0:
        raise NotImplementedError("l2 regularisation (alpha_W with l1_ratio < 1) is not implemented")
    lam, lam_step = (0.0, l1_reg_W) if (sklearn_l1_accumulate and beta == "kullback-leibler") else (l1_reg_W, 0.0)

    # the dictionary stays resident across calls with the same (unchanged) array: dictionary.DictionaryCache
    own = dictionary is None and not cache_dictionaries
    d = dictionary if dictionary is not None else (
        ExemplarDictionary(H_arr, None, mode=mode) if own
        else dictionary_cache.get(H_arr, None, mode, validate=lambda: _check_finite("H", H_arr)))
    try:
        t0 = time.time()
        act = d.solve(X_in, beta_loss=beta, tol=tol, max_iter=max_iter, lam=lam, lambda_step=lam_step)
        W_out = d.to_host(act.H).astype(X_in.dtype, copy=False)
        if _device_result is not None and not own:     # script-level callers go on with the device-resident result
            _device_result.update(dictionary=d, activation=act)
        if verbose:
            print("Epoch %02d reached after %.3f seconds, error: %f" % (act.n_iter, time.time() - t0, act.objective))
    finally:
        if own:
            d.close()
    if act.n_iter == max_iter and tol > 0:     # sklearn _nmf.py:1722-1727
        warnings.warn("Maximum number of iterations %d reached. Increase it to improve convergence." % max_iter,
                      ConvergenceWarning)
    return W_out, H, act.n_iter
