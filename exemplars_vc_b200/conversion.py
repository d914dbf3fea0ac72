"""Drop-in for ``05_conversion.py``: the exploratory single-frame decomposition.

    _get_conversion_data(audiodatum, fs, refine_f0) -> features dict       05_conversion.py:47-70
    io_load_from_pickle(speaker) -> (exemplar, exemplar_W)                 05_conversion.py:73-81
    decompose_frame(frame, _W) -> h                                        05_conversion.py:94-107 (the body of __main__)

``_get_conversion_data`` needs pyworld, which is not part of this image; it is imported lazily so the
module itself always imports.  The reference's ``__main__`` builds ``NMF(n_components=N)`` and calls
``fit_transform(frame.T[:, None], W=_W.T)`` (one (513,1) frame against the stacked dictionary) and then
drops into pdb; ``decompose_frame`` is that computation on the GPU.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from .nmf import non_negative_factorization

feature_path = "data/vc"        # config/config [PATH] feature_path
mode = "3xtf32"


def _get_conversion_data(audiodatum, fs, refine_f0):
    """WORLD features (sp, ap, f0) of one utterance -- 05_conversion.py:47-70."""
    import pyworld as pw  # noqa: deferred, absent from this image

    _f0, t = pw.dio(audiodatum, fs)
    f0 = pw.stonemask(audiodatum, _f0, t, fs) if refine_f0 else _f0
    sp = pw.cheaptrick(audiodatum, f0, t, fs)
    ap = pw.d4c(audiodatum, f0, t, fs)
    return {"sp": sp, "ap": ap, "f0": f0, "fs": fs, "sr": fs}


def io_load_from_pickle(speaker):
    """05_conversion.py:73-81 (the reference loads the same pickle twice; kept)."""
    pickle_path = os.path.join(feature_path, "exem_dict")
    with open(os.path.join(pickle_path, "{}_feat_sp_ap_f0.pkl".format(speaker)), "rb") as f:
        exemplar_speaker = pickle.load(f)
    with open(os.path.join(pickle_path, "{}_feat_sp_ap_f0.pkl".format(speaker)), "rb") as f:
        exemplar_W_speaker = pickle.load(f)
    return exemplar_speaker, exemplar_W_speaker


def stack_dictionary(exemplar_W, drop_last=150, key="sp"):
    """05_conversion.py:94-98: ``_W.extend(exemplar_W_A[i]['sp'])`` over all but the last 150 files."""
    files = exemplar_W[:-drop_last] if drop_last else exemplar_W
    return np.concatenate([np.asarray(f[key]) for f in files], axis=0)


def decompose_frame(frame, _W, beta_loss="kullback-leibler", max_iter=200, tol=1e-4):
    """Activations h (N,) of ONE frame (F,) over the stacked dictionary _W (N,F) -- 05_conversion.py:100-106."""
    frame = np.asarray(frame, dtype=np.asarray(_W).dtype).reshape(1, -1)
    W_act, _, _ = non_negative_factorization(X=frame, H=np.asarray(_W), init="custom", update_H=False,
                                             n_components=np.asarray(_W).shape[0], beta_loss=beta_loss,
                                             solver="mu", tol=tol, max_iter=max_iter, mode=mode)
    return W_act[0]
