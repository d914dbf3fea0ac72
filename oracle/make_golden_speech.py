"""Reduced reproduction of BASELINE.json configs[0] (SF1 -> TF1, utterance 100162) from the wav files the
reference ships, for a real-speech parity fixture.  TEST INFRASTRUCTURE; run in the authoring container.

The reference's own dictionaries are unavailable (npy/SF1.npy, npy/TF1.npy are missing large blobs and
data/vc/exem_dict/*.pkl was never committed) and pyworld is not installed, so this script
rebuilds a SMALL dictionary with scipy-only restatements of the upstream steps:
  * 513-bin magnitude spectra: scipy.signal.stft, n_fft = 1024, hop = 80 (5 ms at 16 kHz) -- the stand-in for
    WORLD cheaptrick's 513-bin envelope (03_a_b_r_parallel.py:86-100),
  * per-file DTW with squared-L2 local cost (01_make_dict_parallel.py:226) on 24 log-band energies,
  * aligned-frame gather (04_align_n_nmf.py:113-124): A[i] = source frame, B[i] = its DTW partner.
Then the reference's exact operator call (KL, max_iter = 150, tol = 1e-4) gives the golden activations.
Only the arrays (float32) are stored: tests/golden/speech_sf1_tf1_100162.npz.
"""
import os
import sys

import numpy as np
from scipy.io import wavfile
from scipy.signal import stft

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nmf_oracle as o  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def spectrum(path):
    sr, x = wavfile.read(path)
    x = x.astype(np.float64) / 32768.0
    _, _, Z = stft(x, fs=sr, window="hann", nperseg=1024, noverlap=1024 - 80, nfft=1024, boundary=None, padded=False)
    return np.abs(Z).T          # (frames, 513)


def bands(S, n=24):
    edges = np.linspace(0, S.shape[1], n + 1).astype(int)
    return np.log(np.stack([S[:, a:b].mean(1) for a, b in zip(edges, edges[1:])], 1) + 1e-6)


def dtw_path(a, b):
    ta, tb = len(a), len(b)
    cost = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)       # squared L2, 01_make_dict_parallel.py:226
    D = np.full((ta + 1, tb + 1), np.inf)
    D[0, 0] = 0.0
    for i in range(1, ta + 1):
        for j in range(1, tb + 1):
            D[i, j] = cost[i - 1, j - 1] + min(D[i - 1, j - 1], D[i - 1, j], D[i, j - 1])
    i, j, path = ta, tb, []
    while i > 0 and j > 0:
        path.append((i - 1, j - 1))
        k = int(np.argmin([D[i - 1, j - 1], D[i - 1, j], D[i, j - 1]]))
        i, j = (i - 1, j - 1) if k == 0 else ((i - 1, j) if k == 1 else (i, j - 1))
    return path[::-1]


def main_full():
    """BASELINE configs[0] at (nearly) its real size: the dictionary from ALL 8 parallel SF1/TF1 pairs the reference
    ships (about 6k aligned exemplar pairs), the WHOLE utterance 100162 (T = 688 frames), the reference's defaults
    (max_iter = 150, tol = 1e-4).  To keep the fixture small the inputs are rounded to float16 FIRST and the
    reference call runs on exactly those values (so the stored float16 arrays are the bit-exact inputs); of the
    activations only every 8th frame is stored (frames are independent once n_iter is fixed), plus Y, n_iter and the
    objective.  -> tests/golden/speech_sf1_tf1_100162_full.npz"""
    A_rows, B_rows = [], []
    for k in range(1, 9):
        f = "10000%d" % k
        Sa, Sb = spectrum(f"{REF}/data/SF1/{f}.wav"), spectrum(f"{REF}/data/TF1/{f}.wav")
        for i, j in dtw_path(bands(Sa), bands(Sb)):
            A_rows.append(Sa[i]); B_rows.append(Sb[j])
    A, B = np.asarray(A_rows) + 1e-6, np.asarray(B_rows) + 1e-6
    X = spectrum(f"{REF}/wav/SF1_100162.wav")
    A16, B16, X16 = A.astype(np.float16), B.astype(np.float16), X.astype(np.float16)
    assert np.isfinite(A16).all() and np.isfinite(B16).all() and np.isfinite(X16).all()
    A64, B64, X64 = A16.astype(np.float64), B16.astype(np.float64), X16.astype(np.float64)
    W0 = o.initial_activation(X64, A64.shape[0])
    obj0 = o.kl_objective(X64, W0, A64)
    W, n_iter = o.reference_call(X64, A64, tol=1e-4, max_iter=150)
    obj = o.kl_objective(X64, W, A64)
    idx = np.arange(0, X.shape[0], 8)
    import sklearn
    np.savez_compressed(os.path.join(OUT, "speech_sf1_tf1_100162_full.npz"), X16=X16, A16=A16, B16=B16,
                        frame_idx=idx, W_sub=W[idx].astype(np.float32), n_iter=n_iter, objective=obj,
                        objective_at_init=obj0, Y=(W @ B64).astype(np.float32), tol=1e-4, max_iter=150,
                        versions=np.array([f"sklearn={sklearn.__version__}", f"numpy={np.__version__}"]))
    print("full: A", A.shape, "X", X.shape, "n_iter", n_iter, "objective", obj, "at init", obj0,
          "sparsity(H<1e-6*max)", float((W < 1e-6 * W.max()).mean()))


def main():
    A_rows, B_rows = [], []
    for f in ("100002", "100004", "100007"):
        Sa, Sb = spectrum(f"{REF}/data/SF1/{f}.wav"), spectrum(f"{REF}/data/TF1/{f}.wav")
        for i, j in dtw_path(bands(Sa), bands(Sb)):
            A_rows.append(Sa[i]); B_rows.append(Sb[j])
    A, B = np.asarray(A_rows), np.asarray(B_rows)
    keep = np.linspace(0, len(A) - 1, 768).astype(int)          # reduced dictionary: every k-th aligned pair
    A, B = A[keep] + 1e-7, B[keep] + 1e-7
    Xfull = spectrum(f"{REF}/wav/SF1_100162.wav")
    e = Xfull.sum(1)
    start = int(np.argmax(np.convolve(e, np.ones(64), "valid")))   # the 64 most energetic consecutive frames
    X = Xfull[start:start + 64]
    A, B, X = A.astype(np.float32), B.astype(np.float32), X.astype(np.float32)
    W, n_iter = o.reference_call(X.astype(np.float64), A.astype(np.float64), tol=1e-4, max_iter=150)
    obj = o.kl_objective(X.astype(np.float64), W, A.astype(np.float64))
    import sklearn
    np.savez_compressed(os.path.join(OUT, "speech_sf1_tf1_100162.npz"), X=X, A=A, B=B, W=W.astype(np.float32),
                        n_iter=n_iter, objective=obj, Y=(W @ B.astype(np.float64)).astype(np.float32), tol=1e-4,
                        max_iter=150, frames_total=Xfull.shape[0], frame_start=start,
                        versions=np.array([f"sklearn={sklearn.__version__}", f"numpy={np.__version__}"]))
    print("A", A.shape, "X", X.shape, "of", Xfull.shape, "n_iter", n_iter, "objective", obj,
          "sparsity(H<1e-6*max)", float((W < 1e-6 * W.max()).mean()))


if __name__ == "__main__":
    (main_full if "--full" in sys.argv else main)()
