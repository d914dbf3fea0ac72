"""Dictionary construction on the device.

The reference gathers DTW-aligned frames with a Python double loop (04_align_n_nmf.py:113-124) and stacks the
per-file lists with ``list.extend`` / ``np.asarray`` (04_align_n_nmf.py:230-246).  Here the per-file feature
matrices are uploaded once and the aligned (and optionally +-context-stacked) exemplar matrices A and B are
built by one gather kernel each, directly in the layout ``ExemplarDictionary`` consumes.  ``context=2`` gives the
stacked exemplars of BASELINE.json's F = 5*513 = 2565 configuration; the frames to convert go through
``stack_frames`` with the same context.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .dictionary import ExemplarDictionary, _ptr, _stream


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("exemplars_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def gather_stack(frames: torch.Tensor, idx: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, context: int = 0):
    """out[k] = concat_{d=-context..context} frames[clamp(idx[k]+d, lo[k], hi[k]-1)]  -> (len(idx), (2c+1)*F)."""
    dev = frames.device
    frames = frames.to(torch.float32).contiguous()
    idx, lo, hi = (t.to(device=dev, dtype=torch.int32).contiguous() for t in (idx, lo, hi))
    n_frames, F = frames.shape
    n_out = int(idx.numel())
    width = (2 * context + 1) * F
    ld_out = (width + 3) // 4 * 4
    out = torch.zeros((max(n_out, 1), ld_out), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evc_gather_stack(_ptr(frames), F, n_frames, F, _ptr(idx), _ptr(lo), _ptr(hi), n_out,
                                               context, _ptr(out), ld_out, _stream(dev)))
    return out[:n_out, :width]


def stack_frames(X, context: int = 0) -> torch.Tensor:
    """Context-stack the frames of ONE utterance: row t = [x[t-c], ..., x[t+c]] (clamped at the ends)."""
    dev = _device()
    x = torch.as_tensor(np.asarray(X) if not isinstance(X, torch.Tensor) else X).to(dev, torch.float32)
    T = x.shape[0]
    ar = torch.arange(T, device=dev, dtype=torch.int32)
    return gather_stack(x, ar, torch.zeros_like(ar), torch.full_like(ar, T), context)


def build_dictionaries(src_files: Sequence, tar_files: Sequence, src_paths: Sequence, tar_paths: Sequence,
                       context: int = 0, mode: str = "3xtf32", key: Optional[str] = None) -> ExemplarDictionary:
    """Aligned exemplar pair (A, B) from per-file features and DTW index paths, resident on the GPU.

    src_files[i], tar_files[i] : (n_i, F) / (m_i, F) feature matrices of file i (or dicts holding them under `key`,
                                 like the reference's per-file feature dicts with 'sp' / 'ap' / 'real')
    src_paths[i], tar_paths[i] : equal-length integer index paths from DTW (exemplar_W_A / exemplar_W_B,
                                 01_make_dict_parallel.py:325-339): exemplar k of file i pairs
                                 src_files[i][src_paths[i][k]] with tar_files[i][tar_paths[i][k]]
    """
    dev = _device()

    def mats(files):
        out = []
        for f in files:
            m = f[key] if key is not None else f
            m = np.asarray(m, dtype=np.float32)
            out.append(np.abs(m) if key == "real" else m)       # 04_align_n_nmf.py:323 takes |real(stft)|
        return out

    S, Tg = mats(src_files), mats(tar_files)
    if not (len(S) == len(Tg) == len(src_paths) == len(tar_paths)):
        raise ValueError("src_files, tar_files, src_paths and tar_paths must have one entry per file")

    def side(files, paths):
        offs = np.concatenate([[0], np.cumsum([len(m) for m in files])]).astype(np.int64)
        idx, lo, hi = [], [], []
        for i, p in enumerate(paths):
            p = np.asarray(p, dtype=np.int64)
            if p.size and (p.min() < 0 or p.max() >= len(files[i])):
                raise ValueError(f"alignment path of file {i} indexes outside the file")
            idx.append(p + offs[i]); lo.append(np.full(p.shape, offs[i])); hi.append(np.full(p.shape, offs[i + 1]))
        frames = torch.from_numpy(np.concatenate(files, axis=0)).to(dev)
        to_t = lambda a: torch.from_numpy(np.concatenate(a).astype(np.int32)).to(dev)  # noqa: E731
        return gather_stack(frames, to_t(idx), to_t(lo), to_t(hi), context)

    for i, (a, b) in enumerate(zip(src_paths, tar_paths)):
        if len(a) != len(b):
            raise ValueError(f"file {i}: source and target alignment paths differ in length")
    A = side(S, src_paths)
    B = side(Tg, tar_paths)
    return ExemplarDictionary(A, B, mode=mode)


def dtw_alignment(feat_full_A: Sequence, feat_full_B: Sequence):
    """DTW index paths of parallel utterances: drop-in for ``01_make_dict_parallel.dtw_alignment`` (:239-249), which
    runs ``dtw(feat_A.T, feat_B.T, lambda x, y: sum(np.square(x - y)))`` (:226) per file pair in a process pool.

    feat_full_A[i], feat_full_B[i] : (order, n_frames) feature matrices of file i, the reference's layout (:219-220)
    returns (dtw_paths, None, None) like the reference; dtw_paths[i] = (p, q): equal-length int arrays,
    exemplar k of file i pairs frame p[k] of A with frame q[k] of B (what make_exemplar_dict_* index with, :205-206).
    All file pairs are aligned by one launch (one block per pair, float64, bit-identical to the package's recursion)."""
    dev = _device()
    if len(feat_full_A) != len(feat_full_B):
        raise ValueError("feat_full_A and feat_full_B must hold the same number of files")
    n = len(feat_full_A)
    if n == 0:
        return [], None, None
    fa = [np.ascontiguousarray(np.asarray(a, dtype=np.float64).T) for a in feat_full_A]      # (frames, order)
    fb = [np.ascontiguousarray(np.asarray(b, dtype=np.float64).T) for b in feat_full_B]
    dim = fa[0].shape[1]
    if any(m.ndim != 2 or m.shape[1] != dim for m in fa + fb):
        raise ValueError("every file must be a 2-D (order, n_frames) matrix of the same order")
    ra, cb = np.array([len(m) for m in fa], np.int64), np.array([len(m) for m in fb], np.int64)
    if (ra < 1).any() or (cb < 1).any():
        raise ValueError("every file needs at least one frame")
    a_off, b_off = np.concatenate([[0], np.cumsum(ra)]), np.concatenate([[0], np.cumsum(cb)])
    dir_off = np.concatenate([[0], np.cumsum(ra * cb)])
    path_off = np.concatenate([[0], np.cumsum(ra + cb)])
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)    # noqa: E731
    A, B = t(np.concatenate(fa), torch.float64), t(np.concatenate(fb), torch.float64)
    a_o, b_o, d_o, p_o = (t(x, torch.int64) for x in (a_off, b_off, dir_off, path_off))
    dirs = torch.empty(int(dir_off[-1]), dtype=torch.uint8, device=dev)
    pa = torch.empty(int(path_off[-1]), dtype=torch.int32, device=dev)
    pb = torch.empty_like(pa)
    plen = torch.empty(n, dtype=torch.int32, device=dev)
    dist = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evc_dtw(_ptr(A), _ptr(a_o), _ptr(B), _ptr(b_o), n, dim, int(max(ra.max(), 1)), _ptr(dirs),
                                      _ptr(d_o), _ptr(pa), _ptr(pb), _ptr(p_o), _ptr(plen), _ptr(dist), _stream(dev)))
    pa_h, pb_h, len_h = pa.cpu().numpy(), pb.cpu().numpy(), plen.cpu().numpy()
    paths = []
    for i in range(n):
        s0, L = int(path_off[i]), int(len_h[i])
        paths.append((pa_h[s0:s0 + L][::-1].astype(np.int64), pb_h[s0:s0 + L][::-1].astype(np.int64)))
    dtw_alignment.last_distances = dist.cpu().numpy()
    return paths, None, None
