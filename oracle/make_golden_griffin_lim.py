"""Golden vectors for the Griffin-Lim vocoder from the reference's OWN code.  TEST INFRASTRUCTURE; run in the
authoring container (it reads /root/reference).

``zz_audio_utilities.py`` starts with ``from pylab import *`` (matplotlib is not in this image); the three functions
on the path only use numpy names from it (``sqrt``, ``sum``), so a stand-in ``pylab`` module that re-exports numpy is
injected and the UNMODIFIED reference file is imported and run:

    reconstruct_signal_griffin_lim(mag, 400, 80, iterations)     zz_audio_utilities.py:258-292
    stft_for_reconstruction / istft_for_reconstruction            :181-218

with fft_size = 400, hop = 80 as at 04_align_n_nmf.py:46-47, 187.  The start signal is the reference's own draw
(np.random.seed(7); np.random.randn(len)), stored so the device path can start from the same x0.
-> tests/golden/griffin_lim_400_80.npz
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_module():
    shim = types.ModuleType("pylab")
    shim.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
    sys.modules.setdefault("pylab", shim)
    sys.path.insert(0, "/root/reference")
    import zz_audio_utilities as z
    return z


def main():
    z = reference_module()
    fft, hop, T = 400, 80, 24
    rng = np.random.default_rng(3)
    # a magnitude spectrogram of a real-ish signal: two chirping partials + noise
    n = T * hop + fft
    t = np.arange(n) / 16000.0
    sig = np.sin(2 * np.pi * (300 + 2000 * t) * t) + 0.5 * np.sin(2 * np.pi * 1200 * t) + 0.05 * rng.standard_normal(n)
    S = z.stft_for_reconstruction(sig, fft, hop)
    mag = np.abs(S).astype(np.float32)
    assert mag.shape == (T, fft // 2 + 1)
    out = {}
    for iters in (1, 3, 30):
        np.random.seed(7)
        x0 = np.random.randn(n)
        np.random.seed(7)
        with contextlib.redirect_stdout(io.StringIO()):
            out[iters] = z.reconstruct_signal_griffin_lim(mag.astype(np.float64), fft, hop, iters)
    np.savez_compressed(os.path.join(OUT, "griffin_lim_400_80.npz"), mag=mag, x0=x0, sig=sig, stft_re=S.real, stft_im=S.imag,
                        istft=z.istft_for_reconstruction(S, fft, hop), x1=out[1], x3=out[3], x30=out[30],
                        fft_size=fft, hop=hop)
    print("griffin-lim golden:", mag.shape, "len", n, "rms x30", float(np.sqrt(np.mean(out[30] ** 2))))


if __name__ == "__main__":
    main()
