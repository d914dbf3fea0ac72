"""CPU restatement of the reference's Griffin-Lim vocoder.  TEST INFRASTRUCTURE ONLY (tests/, smoke, bench CPU legs).

Follows ``zz_audio_utilities.py`` of the reference line by line (vectorised over frames where the reference loops):
    stft_for_reconstruction          :181-196   window * x[i:i+fft] for i in range(0, len(x)-fft, hop), np.fft.rfft
    istft_for_reconstruction         :199-218   x[i:i+fft] += window * real(irfft(X[n])), len = T*hop + fft
    reconstruct_signal_griffin_lim   :258-292   x <- istft(mag * exp(1j * angle(stft(x)))), `iterations` times
Pinned against outputs of the reference's own functions (oracle/make_golden_griffin_lim.py ->
tests/golden/griffin_lim_400_80.npz) in tests/test_oracle.py.
"""
import numpy as np


def stft_for_reconstruction(x, fft_size, hopsamp):
    window = np.hanning(fft_size)
    fft_size, hopsamp = int(fft_size), int(hopsamp)
    starts = range(0, len(x) - fft_size, hopsamp)
    return np.array([np.fft.rfft(window * x[i:i + fft_size]) for i in starts])


def istft_for_reconstruction(X, fft_size, hopsamp):
    fft_size, hopsamp = int(fft_size), int(hopsamp)
    window = np.hanning(fft_size)
    time_slices = X.shape[0]
    x = np.zeros(int(time_slices * hopsamp + fft_size))
    for n, i in enumerate(range(0, len(x) - fft_size, hopsamp)):
        x[i:i + fft_size] += window * np.real(np.fft.irfft(X[n]))
    return x


def reconstruct_signal_griffin_lim(magnitude_spectrogram, fft_size, hopsamp, iterations, x0):
    """x0 replaces the reference's np.random.randn(len_samples) (:279)."""
    x = np.array(x0, dtype=np.float64, copy=True)
    for _ in range(int(iterations)):
        S = stft_for_reconstruction(x, fft_size, hopsamp)
        proposal = magnitude_spectrogram * np.exp(1.0j * np.angle(S))
        x = istft_for_reconstruction(proposal, fft_size, hopsamp)
    return x
