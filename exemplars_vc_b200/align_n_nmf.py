"""Drop-in for the hot-path functions of the reference's ``04_align_n_nmf.py``.

Same names, signatures and return shapes:

    _factorize(X, W, beta_loss="kullback-leibler", tol=1e-4) -> H (N,T)       04_align_n_nmf.py:194-215
    factorize(tobe_converted, src_feat) -> (H dict, R dict | None)            04_align_n_nmf.py:218-333
    convert(H, tar_feat, residual) -> dict | ndarray                          04_align_n_nmf.py:336-393

The reference module cannot be imported (its name starts with a digit and importing it opens log files,
parses config/config and imports dtw/pyworld/librosa), so the path lives here under an importable name.
Module-level knobs replace what the reference reads from ``config/config`` at import time:

    use_stft        -- [VAR] use_stft (config/config:12); 1 = |real(stft)| branch, 0 = WORLD sp/ap/f0
    beta_override   -- the reference BODY overwrites its beta_loss argument with "frobenius"
                       (04_align_n_nmf.py:210), so the script as committed runs Frobenius updates whatever
                       the caller passes.  None (default) honours the argument instead, i.e. the signature's
                       default "kullback-leibler" runs (the KL path this package is built around); that is a
                       DIFFERENT result from the committed script, so the first such call warns.  Set
                       beta_override = "frobenius" (or EVC_SCRIPT_LITERAL=1 in the environment) to reproduce
                       the script literally (tests/test_parity_round2_gpu.py checks that path against a
                       golden from the reference's own call).
    mode            -- arithmetic of the contractions ("3xtf32" fp32-accurate | "tf32" | "bf16" | "fp32")
    cache_dir       -- None (default): never reuse a stale H.  The reference's pickle cache
                       (04_align_n_nmf.py:251-255, keyed by feature type and file count only) is unsafe.
"""
from __future__ import annotations

import hashlib
import logging
import os
import pickle
import warnings

import numpy as np

from .dictionary import dictionary_cache
from .nmf import non_negative_factorization

use_stft = 1
beta_override = "frobenius" if os.environ.get("EVC_SCRIPT_LITERAL") == "1" else None
mode = "3xtf32"
cache_dir = None
max_iter = 150          # 04_align_n_nmf.py:213

# The F = 1 f0 track (04_align_n_nmf.py:288) has no use for tensor cores: route it to the FFMA kernels.
_SMALL_F = 8
_warned_beta = False


def _mode_for(F: int) -> str:
    return "fp32" if F < _SMALL_F else mode


def _factorize(X, W, beta_loss="kullback-leibler", tol=1e-4):
    """Calculate matrix ``H`` with ``W x H ~ X`` for the fixed exemplar dictionary ``W``.

    :param X: frames to decompose, (T, F) -- rows are frames, like the reference's caller passes them
    :param W: exemplar dictionary, (N, F)
    :return: H (N, T): the transposed activations, as 04_align_n_nmf.py:215 returns ``_W.T``
    """
    global _warned_beta
    if beta_override is not None:
        beta_loss = beta_override
    elif beta_loss != "frobenius" and not _warned_beta:
        _warned_beta = True
        warnings.warn("exemplars_vc_b200.align_n_nmf._factorize runs beta_loss=%r as passed; the reference script "
                      "overwrites it with 'frobenius' (04_align_n_nmf.py:210). Set align_n_nmf.beta_override = "
                      "'frobenius' (or EVC_SCRIPT_LITERAL=1) for the script-literal result." % (beta_loss,),
                      stacklevel=2)
    X = np.asarray(X)
    W = np.asarray(W)
    _W, _H, n_iter = non_negative_factorization(
        X=X, H=W, init="custom", update_H=False, n_components=W.shape[0], beta_loss=beta_loss, solver="mu",
        tol=tol, max_iter=max_iter, verbose=0, mode=_mode_for(W.shape[1]), _device_result=_keep)
    return _W.T


_keep = None    # set by _factorize_with_residual: receives the device-resident dictionary + activations of the call


def _factorize_with_residual(X, W):
    """H (N,T) as `_factorize`, plus the WORLD-branch residual log(H^T W - X) (04_align_n_nmf.py:292-294) formed on the
    device from the activations that are still resident there (no second upload, no host matmul)."""
    global _keep
    _keep = {}
    try:
        H = _factorize(X=X, W=W)
        d, act = _keep.get("dictionary"), _keep.get("activation")
    finally:
        _keep = None
    if d is None:          # dictionary cache switched off: the handle is gone, go through a fresh one
        d = dictionary_cache.get(np.asarray(W), None, _mode_for(np.asarray(W).shape[1]))
        R = d.to_host(d.residual(X, np.ascontiguousarray(H.T)))
    else:
        R = d.to_host(d.residual(X, act.H))
    return H, R.astype(np.result_type(H.dtype, np.asarray(X).dtype), copy=False)


def _stack(feats, key, absolute=False):
    rows = []
    for f in feats:                                   # 04_align_n_nmf.py:236-239: list.extend over files
        v = np.asarray(f[key])
        rows.append(np.abs(v) if absolute else v)
    out = np.concatenate([r if r.ndim > 1 else r[:, np.newaxis] for r in rows], axis=0)
    return out


def _cache_path(tag, arrays):
    if cache_dir is None:
        return None
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    h.update(repr((beta_override, mode, max_iter)).encode())
    return os.path.join(cache_dir, "H_%s_%s.pkl" % (tag, h.hexdigest()[:24]))


def factorize(tobe_converted, src_feat):
    """Audio-file level activation estimation (04_align_n_nmf.py:218-333).

    :param tobe_converted: dict with 'sp','ap','f0' (WORLD branch) or 'real' (STFT branch) of the utterance
    :param src_feat: list of per-file dicts of aligned source exemplars with the same keys
    :return: (H, R): H = {'H_sp','H_ap','H_f0'} or {'H_stft'}, each (N, T); R = residual dict or None
    """
    logging.info("Start calculating the activation matrix H ...")
    if not use_stft:
        conv_sp, conv_ap = np.asarray(tobe_converted["sp"]), np.asarray(tobe_converted["ap"])
        conv_f0 = np.asarray(tobe_converted["f0"])[:, np.newaxis]
        A_sp, A_ap, A_f0 = _stack(src_feat, "sp"), _stack(src_feat, "ap"), _stack(src_feat, "f0")
        path = _cache_path("sp_ap_f0", [conv_sp, conv_ap, conv_f0, A_sp, A_ap, A_f0])
        if path and os.path.isfile(path):
            with open(path, "rb") as f:
                return pickle.load(f)
        # activations + residual compensation log(H^T A - X), 04_align_n_nmf.py:284-294 (NaN wherever H^T A < X, by
        # construction); the product and the logarithm run on the device, next to the activations
        H, R = {}, {}
        for k, X_k, A_k in (("sp", conv_sp, A_sp), ("ap", conv_ap, A_ap), ("f0", conv_f0, A_f0)):
            H["H_" + k], R["r_" + k] = _factorize_with_residual(X_k, A_k)
        if path:
            os.makedirs(cache_dir, exist_ok=True)
            with open(path, "wb") as f:
                pickle.dump((H, R), f)
        return H, R
    conv_stft = np.abs(np.asarray(tobe_converted["real"]))           # 04_align_n_nmf.py:315
    A_stft = _stack(src_feat, "real", absolute=True)                  # :319-323
    path = _cache_path("stft", [conv_stft, A_stft])
    if path and os.path.isfile(path):
        with open(path, "rb") as f:
            return pickle.load(f), None
    H = {"H_stft": _factorize(X=conv_stft, W=A_stft)}
    if path:
        os.makedirs(cache_dir, exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(H, f)
    return H, None


def _product(H_nt, B, residual=None):
    """np.matmul(H.T, B) on the GPU (04_align_n_nmf.py:391); with `residual` the whole WORLD-branch expression
    exp(log(H.T @ B) + log(residual)) of 04_align_n_nmf.py:371-373 in one pass (NaN -> 0 rule of :363-365 included).
    The target dictionary stays resident between calls (dictionary.DictionaryCache)."""
    H_nt = np.asarray(H_nt)
    B = np.asarray(B)
    d = dictionary_cache.get(B, B, _mode_for(B.shape[1]))
    y = d.to_host(d.convert(np.ascontiguousarray(H_nt.T), residual=residual))
    return y.astype(np.result_type(H_nt.dtype, B.dtype), copy=False)


def convert(H, tar_feat, residual):
    """From H and the aligned target exemplars, calculate the converted feature (04_align_n_nmf.py:336-393)."""
    logging.info("Using H for conversion ...")
    if not use_stft:
        B_sp, B_ap, B_f0 = _stack(tar_feat, "sp"), _stack(tar_feat, "ap"), _stack(tar_feat, "f0")
        for k in ("r_sp", "r_ap", "r_f0"):
            residual[k][np.isnan(residual[k])] = 0          # :363-365 (in place, like the reference)
        # exp(log(H^T B) + log(r)), :371-373, as the epilogue of the conversion product on the device
        converted_sp = _product(H["H_sp"], B_sp, residual["r_sp"])
        converted_ap = _product(H["H_ap"], B_ap, residual["r_ap"])
        converted_f0 = _product(H["H_f0"], B_f0, residual["r_f0"])
        return {"sp": converted_sp, "ap": converted_ap, "f0": np.squeeze(converted_f0)}
    B_stft = _stack(tar_feat, "real", absolute=True)
    return _product(H["H_stft"], B_stft)                  # :391
