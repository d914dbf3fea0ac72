"""Griffin-Lim at the reference's real call shape (688 frames x 201 bins, fft 400, hop 80, 300 iterations,
04_align_n_nmf.py:187): device time next to the numpy oracle's (a restatement of the reference's per-frame
np.fft.rfft list comprehension; its own log shows ~9 s).  Usage: python tests/manual/griffin_lim_timing.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import audio_utilities as au  # noqa: E402
from oracle import griffin_lim_oracle as o  # noqa: E402

T, fft, hop, iters = 688, 400, 80, 300
rng = np.random.default_rng(1)
n = T * hop + fft
sig = np.sin(np.arange(n) * 0.03) + 0.3 * np.sin(np.arange(n) * 0.11) + 0.01 * rng.standard_normal(n)
mag = np.abs(o.stft_for_reconstruction(sig, fft, hop)).astype(np.float32)
x0 = rng.standard_normal(n)
au.reconstruct_signal_griffin_lim(mag, fft, hop, 3, x0=x0)
torch.cuda.synchronize()
t0 = time.perf_counter()
x = au.reconstruct_signal_griffin_lim(mag, fft, hop, iters, x0=x0)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
ref = o.reconstruct_signal_griffin_lim(mag.astype(np.float64), fft, hop, iters, x0)
t_cpu = time.perf_counter() - t0
print(f"griffin-lim T={T} fft={fft} hop={hop} {iters} iterations: device {t_dev * 1e3:.1f} ms (host buffers in/out), "
      f"numpy oracle {t_cpu:.2f} s, max rel diff {np.abs(x - ref).max() / np.abs(ref).max():.2e}")
