"""Drop-in for the hot-path functions of ``04_align_n_nmf_pytorch.py`` (which contains no PyTorch).

    _factorize(X, W, beta_loss="kullback-leibler", tol=1e-4) -> H (N,T)       04_align_n_nmf_pytorch.py:189-210
    factorize(tobe_converted, src_feat) -> H dict                              :213-289
    convert(H, tar_feat) -> dict                                               :292-327

That variant asks scikit-learn for ``solver='cd'`` with ``max_iter=200`` (:207-208).  Coordinate descent
is a different, inherently sequential algorithm (footnote to row 2) and is out of
scope; the signatures are kept and routed to the multiplicative-update solver with this variant's
iteration budget.  No residual is produced and ``convert`` takes two arguments, as in the reference.
"""
from __future__ import annotations

import logging
import warnings

import numpy as np

from . import align_n_nmf as _mu
from .nmf import non_negative_factorization

use_stft = 0
beta_override = None
mode = "3xtf32"
max_iter = 200          # 04_align_n_nmf_pytorch.py:208
_warned_solver = False


def _factorize(X, W, beta_loss="kullback-leibler", tol=1e-4):
    global _warned_solver
    if not _warned_solver:
        _warned_solver = True
        warnings.warn("exemplars_vc_b200.align_n_nmf_pytorch runs multiplicative updates (solver='mu'); the reference "
                      "variant asks scikit-learn for solver='cd' with beta_loss='frobenius' "
                      "(04_align_n_nmf_pytorch.py:205-208), a different algorithm with a different fixed point.",
                      stacklevel=2)
    if beta_override is not None:
        beta_loss = beta_override
    X = np.asarray(X)
    W = np.asarray(W)
    m = "fp32" if W.shape[1] < _mu._SMALL_F else mode
    _W, _H, n_iter = non_negative_factorization(
        X=X, H=W, init="custom", update_H=False, n_components=W.shape[0], beta_loss=beta_loss, solver="mu",
        tol=tol, max_iter=max_iter, verbose=0, mode=m)
    return _W.T


def factorize(tobe_converted, src_feat):
    logging.info("Start calculating the activation matrix H ...")
    if not use_stft:
        conv_sp, conv_ap = np.asarray(tobe_converted["sp"]), np.asarray(tobe_converted["ap"])
        conv_f0 = np.asarray(tobe_converted["f0"])[:, np.newaxis]
        A_sp, A_ap, A_f0 = _mu._stack(src_feat, "sp"), _mu._stack(src_feat, "ap"), _mu._stack(src_feat, "f0")
        return {"H_sp": _factorize(X=conv_sp, W=A_sp), "H_ap": _factorize(X=conv_ap, W=A_ap),
                "H_f0": _factorize(X=conv_f0, W=A_f0)}
    conv_stft = np.asarray(tobe_converted["real"])                   # :270 (no abs in this variant)
    A_stft = _mu._stack(src_feat, "stft")                             # :274-277
    return {"H_stft": _factorize(X=conv_stft, W=A_stft)}


def convert(H, tar_feat):
    logging.info("Using H for conversion ...")
    B_sp, B_ap, B_f0 = _mu._stack(tar_feat, "sp"), _mu._stack(tar_feat, "ap"), _mu._stack(tar_feat, "f0")
    old = _mu.mode
    _mu.mode = mode
    try:
        return {"sp": _mu._product(H["H_sp"], B_sp), "ap": _mu._product(H["H_ap"], B_ap),
                "f0": np.squeeze(_mu._product(H["H_f0"], B_f0))}
    finally:
        _mu.mode = old
