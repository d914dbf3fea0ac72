"""Drop-in for the Griffin-Lim part of the reference's ``zz_audio_utilities.py`` (the vocoder of the STFT branch,
called as ``reconstruct_signal_griffin_lim(np.abs(stft_mag), frame_length, hop_length, 300)`` at
04_align_n_nmf.py:187 with frame_length = 400, hop_length = 80):

    stft_for_reconstruction(x, fft_size, hopsamp) -> (T, fft_size/2+1) complex         zz_audio_utilities.py:181-196
    istft_for_reconstruction(X, fft_size, hopsamp) -> (T*hopsamp + fft_size,) float64   zz_audio_utilities.py:199-218
    reconstruct_signal_griffin_lim(magnitude_spectrogram, fft_size, hopsamp, iterations) zz_audio_utilities.py:258-292

Same names, argument meaning and return shapes; the work runs on the GPU in double precision (libevc_b200,
csrc/audio_kernels.cuh), every frame's STFT -> phase -> inverse STFT fused in one block.  The reference starts from
``np.random.randn(len_samples)``; pass ``x0`` to pin the start signal (that is how parity is tested), otherwise the
same numpy draw is made here.  ``verbose=True`` prints the reference's per-iteration RMSE line.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("exemplars_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _window(fft_size, dev):
    return torch.from_numpy(np.hanning(int(fft_size))).to(dev)       # zz_audio_utilities.py:192, 211


def stft_for_reconstruction(x, fft_size, hopsamp):
    """STFT of the time-domain signal x: rows are time slices, columns frequency bins (zz_audio_utilities.py:181-196)."""
    dev = _dev()
    fft_size, hopsamp = int(fft_size), int(hopsamp)
    xd = torch.as_tensor(np.asarray(x, dtype=np.float64)).to(dev)
    n = int(xd.numel())
    T = len(range(0, n - fft_size, hopsamp))
    if T < 1:
        return np.zeros((0, fft_size // 2 + 1), dtype=np.complex128)
    spec = torch.empty((T, fft_size // 2 + 1, 2), dtype=torch.float64, device=dev)
    check(_lib.lib().evc_stft(_ptr(xd), n, fft_size, hopsamp, _ptr(_window(fft_size, dev)), _ptr(spec), _stream(dev)))
    return torch.view_as_complex(spec).cpu().numpy()


def istft_for_reconstruction(X, fft_size, hopsamp):
    """Invert an STFT (rows = time slices) into a time-domain signal (zz_audio_utilities.py:199-218)."""
    dev = _dev()
    fft_size, hopsamp = int(fft_size), int(hopsamp)
    Xc = np.ascontiguousarray(np.asarray(X, dtype=np.complex128))
    T = int(Xc.shape[0])
    out = torch.zeros(T * hopsamp + fft_size, dtype=torch.float64, device=dev)
    if T < 1:
        return out.cpu().numpy()
    spec = torch.view_as_real(torch.from_numpy(Xc)).contiguous().to(dev)
    check(_lib.lib().evc_istft(_ptr(spec), T, fft_size, hopsamp, _ptr(_window(fft_size, dev)), _ptr(out), _stream(dev)))
    return out.cpu().numpy()


def reconstruct_signal_griffin_lim(magnitude_spectrogram, fft_size, hopsamp, iterations, x0=None, verbose=False):
    """Reconstruct an audio signal from a magnitude spectrogram (Griffin & Lim 1984), zz_audio_utilities.py:258-292.

    magnitude_spectrogram: (T, fft_size/2+1), numpy or a CUDA tensor (e.g. the converted |STFT| straight from
    ExemplarDictionary.convert).  Returns the float64 signal of T*hopsamp + fft_size samples."""
    dev = _dev()
    fft_size, hopsamp, iterations = int(fft_size), int(hopsamp), int(iterations)
    mag = magnitude_spectrogram
    if not isinstance(mag, torch.Tensor):
        mag = torch.from_numpy(np.ascontiguousarray(np.asarray(mag, dtype=np.float32)))
    mag = mag.to(device=dev, dtype=torch.float32).contiguous()
    T = int(mag.shape[0])
    if int(mag.shape[1]) != fft_size // 2 + 1:
        raise ValueError(f"magnitude_spectrogram has {mag.shape[1]} bins, fft_size {fft_size} needs {fft_size // 2 + 1}")
    len_samples = int(T * hopsamp + fft_size)
    if x0 is None:
        x0 = np.random.randn(len_samples)                         # :279, same draw as the reference
    x0d = torch.as_tensor(np.asarray(x0, dtype=np.float64)).to(dev)
    if int(x0d.numel()) != len_samples:
        raise ValueError(f"x0 must have {len_samples} samples")
    out = torch.empty(len_samples, dtype=torch.float64, device=dev)
    sq = torch.zeros(max(iterations, 1), dtype=torch.float64, device=dev)
    check(_lib.lib().evc_griffin_lim(_ptr(mag), int(mag.stride(0)), T, fft_size, hopsamp, iterations,
                                     _ptr(_window(fft_size, dev)), _ptr(x0d), _ptr(out), _ptr(sq), _stream(dev)))
    if verbose:
        for i, v in enumerate(torch.sqrt(sq[:iterations] / len_samples).cpu().tolist()):
            print("Reconstruction iteration: {}/{} RMSE: {} ".format(i + 1, iterations, v))       # :290
    return out.cpu().numpy()
