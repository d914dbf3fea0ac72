"""SURVEY 8f-2 on the device: the WORLD-branch residual epilogues (04_align_n_nmf.py:292-294, 363-373) and the
Griffin-Lim vocoder (zz_audio_utilities.py:181-292), against goldens made by the reference's OWN functions
(oracle/make_golden_griffin_lim.py) and against numpy restatements of the reference expressions.

Tolerances: Griffin-Lim runs in double precision on the device; one STFT / ISTFT agrees with numpy's FFT to 1e-11
relative, 30 iterations from the same start signal to 1e-6 (the phase of near-silent bins is ill-conditioned, so
rounding differences grow slowly with the iteration count).  The residual expressions are elementwise on fp32 products:
1e-4 relative on entries where the difference H A - X is not a cancellation (|H A - X| > 1e-2 |X|)."""
import numpy as np
import pytest

from conftest import load_golden, rel_fro

pytestmark = pytest.mark.gpu


def test_stft_istft_match_the_reference_functions():
    from exemplars_vc_b200 import audio_utilities as au
    g = load_golden("griffin_lim_400_80")
    fft, hop = int(g["fft_size"]), int(g["hop"])
    S = au.stft_for_reconstruction(g["sig"], fft, hop)
    S_ref = g["stft_re"] + 1j * g["stft_im"]
    assert S.shape == S_ref.shape and S.dtype == np.complex128
    assert np.abs(S - S_ref).max() < 1e-11 * np.abs(S_ref).max()
    x = au.istft_for_reconstruction(S_ref, fft, hop)
    assert x.shape == g["istft"].shape
    assert np.abs(x - g["istft"]).max() < 1e-11 * np.abs(g["istft"]).max()


@pytest.mark.parametrize("iters,tol", [(1, 1e-10), (3, 1e-9), (30, 1e-6)])
def test_griffin_lim_matches_the_reference_function(iters, tol):
    from exemplars_vc_b200 import audio_utilities as au
    g = load_golden("griffin_lim_400_80")
    x = au.reconstruct_signal_griffin_lim(g["mag"], int(g["fft_size"]), int(g["hop"]), iters, x0=g["x0"])
    ref = g["x%d" % iters]
    err = np.abs(x - ref).max() / np.abs(ref).max()
    print(f"griffin-lim {iters} iterations: max rel err {err:.2e}")
    assert x.shape == ref.shape and err < tol


def test_griffin_lim_full_utterance_against_oracle():
    """The reference's real call shape: 688 frames x 201 bins, fft 400, hop 80 (04_align_n_nmf.py:187), a few
    iterations against the numpy oracle, and the documented behaviour for a zero spectrogram (silence stays silence)."""
    from exemplars_vc_b200 import audio_utilities as au
    from oracle import griffin_lim_oracle as o
    rng = np.random.default_rng(11)
    T, fft, hop = 688, 400, 80
    n = T * hop + fft
    sig = np.cumsum(rng.standard_normal(n)) * 0.01 + np.sin(np.arange(n) * 0.05)
    mag = np.abs(o.stft_for_reconstruction(sig, fft, hop)).astype(np.float32)
    x0 = rng.standard_normal(n)
    x = au.reconstruct_signal_griffin_lim(mag, fft, hop, 5, x0=x0)
    ref = o.reconstruct_signal_griffin_lim(mag.astype(np.float64), fft, hop, 5, x0)
    assert np.abs(x - ref).max() < 1e-8 * np.abs(ref).max()
    # consistency improves: the spectrogram of the result is closer to the target than that of the start signal
    e0 = rel_fro(np.abs(o.stft_for_reconstruction(x0, fft, hop)), mag)
    e5 = rel_fro(np.abs(o.stft_for_reconstruction(x, fft, hop)), mag)
    assert e5 < 0.8 * e0
    z = au.reconstruct_signal_griffin_lim(np.zeros((8, 201), np.float32), fft, hop, 2, x0=rng.standard_normal(8 * hop + fft))
    assert np.all(z == 0.0)
    with pytest.raises(ValueError):
        au.reconstruct_signal_griffin_lim(mag[:, :100], fft, hop, 1)


def test_residual_epilogues_match_the_reference_expressions():
    """R = log(H^T A - X) and converted = exp(log(H^T B) + log(R')) with NaN -> 0 (04_align_n_nmf.py:292-294, 363-373)."""
    from exemplars_vc_b200 import ExemplarDictionary
    rng = np.random.default_rng(4)
    N, F, T = 700, 257, 40
    A = (rng.random((N, F)) ** 2 + 1e-3).astype(np.float32)
    B = (rng.random((N, F)) ** 2 + 1e-3).astype(np.float32)
    H = (rng.random((T, N)) * (rng.random((T, N)) < 0.05)).astype(np.float32)
    WH = H.astype(np.float64) @ A.astype(np.float64)
    A *= 50.0                                                           # WORLD-sized magnitudes: H A - X exceeds 1, so log > 0
    WH = H.astype(np.float64) @ A.astype(np.float64)
    X = (WH * (0.3 + 1.4 * rng.random((T, F)))).astype(np.float32)      # about half the entries above the model
    X[0, :5] = 0.0
    with ExemplarDictionary(A, B, mode="3xtf32") as d:
        R = d.to_host(d.residual(X, H))
        Y = d.to_host(d.convert(H, residual=R))
        Y_plain = d.to_host(d.convert(H))
    with np.errstate(invalid="ignore", divide="ignore"):
        R_ref = np.log(WH - X.astype(np.float64))
        Rz = R_ref.copy()
        Rz[np.isnan(Rz)] = 0
        Y_ref = np.exp(np.log(H.astype(np.float64) @ B.astype(np.float64)) + np.log(Rz))
    well = np.abs(WH - X) > 1e-2 * np.abs(X)
    assert np.array_equal(np.isnan(R)[well], np.isnan(R_ref)[well])       # same NaN pattern away from cancellations
    ok = well & np.isfinite(R_ref)
    assert np.abs(R[ok] - R_ref[ok]).max() < 1e-4 * max(1.0, np.abs(R_ref[ok]).max())
    # converted: 0 where the residual was NaN (-> 0 -> log 0 = -inf -> exp = 0), NaN where it is negative, Y * r else
    nan_r = np.isnan(R)
    assert np.all(Y[nan_r] == 0.0)
    neg = (~nan_r) & (R < 0)
    assert neg.any() and np.all(np.isnan(Y[neg]))
    pos = ok & (R_ref > 0) & (R > 0)
    assert pos.any()
    # exactly the expression applied to the device's own residual ...
    assert np.abs(Y[pos] / (Y_plain[pos].astype(np.float64) * R[pos]) - 1).max() < 1e-6
    # ... and the float64 reference within the residual's own (absolute) accuracy
    assert np.abs(Y[pos] - Y_ref[pos]).max() < 2e-4 * np.abs(Y_plain[pos]).max() * max(1.0, np.abs(R_ref[pos]).max())
    assert rel_fro(Y_plain, H.astype(np.float64) @ B.astype(np.float64)) < 1e-5


def test_script_level_world_branch_uses_the_device_epilogues():
    """factorize / convert of the WORLD branch end to end (sp, ap, f0) against the numpy restatement of
    04_align_n_nmf.py:284-294, 363-373 fed with the SAME activations."""
    import warnings
    from exemplars_vc_b200 import align_n_nmf as m
    rng = np.random.default_rng(6)
    N, F, T = 260, 129, 11
    As, Aa = 100 * (rng.random((N, F)) ** 2 + 1e-3), rng.random((N, F)) ** 2 + 1e-3
    Bs, Ba = 100 * (rng.random((N, F)) ** 2 + 1e-3), rng.random((N, F)) ** 2 + 1e-3
    f0d = np.where(rng.random(N) < 0.3, 0.0, 100 + 100 * rng.random(N))
    f0t = np.where(f0d > 0, f0d * 1.2, 0.0)
    Hs = rng.random((T, N)) * (rng.random((T, N)) < 0.04)
    # frames that are NOT an exact model (x 0.5 .. 1.5 per bin), so H^T A - X has both signs and exceeds 1 in places
    Xs = (Hs @ As) * (0.5 + rng.random((T, F))) + 0.01
    Xa = Hs @ Aa + 0.01 * rng.random((T, F))
    f0x = np.where(rng.random(T) < 0.3, 0.0, 150 + 50 * rng.random(T))
    old = m.use_stft
    m.use_stft = 0
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            H, R = m.factorize({"sp": Xs, "ap": Xa, "f0": f0x}, [{"sp": As, "ap": Aa, "f0": f0d}])
            Rc = {k: v.copy() for k, v in R.items()}
            out = m.convert(H, [{"sp": Bs, "ap": Ba, "f0": f0t}], R)
    finally:
        m.use_stft = old
    assert set(R) == {"r_sp", "r_ap", "r_f0"} and R["r_sp"].shape == (T, F) and R["r_f0"].shape == (T, 1)
    assert not np.isnan(R["r_sp"]).any()                       # convert() zeroed the NaNs in place, like the reference
    with np.errstate(invalid="ignore", divide="ignore"):
        WH = H["H_sp"].T @ As
        r_ref = np.log(WH - Xs)
        well = np.abs(WH - Xs) > 1e-2 * np.abs(Xs)
        fin = well & np.isfinite(r_ref) & np.isfinite(Rc["r_sp"])
        assert np.array_equal(np.isnan(Rc["r_sp"])[well], np.isnan(r_ref)[well])
        assert np.abs(Rc["r_sp"][fin] - r_ref[fin]).max() < 1e-3 * max(1.0, np.abs(r_ref[fin]).max())
        rz = Rc["r_sp"].copy(); rz[np.isnan(rz)] = 0
        conv_ref = np.exp(np.log(H["H_sp"].T @ Bs) + np.log(rz))
    both = np.isfinite(conv_ref) & np.isfinite(out["sp"]) & (conv_ref > 0)
    assert both.any() and np.isnan(conv_ref).any() and (conv_ref == 0).any()      # all three regimes of the expression occur
    assert np.array_equal(np.isnan(conv_ref), np.isnan(out["sp"]))
    assert np.abs(out["sp"][both] / conv_ref[both] - 1).max() < 1e-3
    assert out["f0"].shape == (T,)


def test_dictionary_cache_reuses_and_never_goes_stale():
    """The numpy-level entry points keep the uploaded dictionary resident between calls (identity + content checksum);
    an in-place edit of the array rebuilds it."""
    import warnings
    from exemplars_vc_b200 import nmf
    from exemplars_vc_b200.dictionary import dictionary_cache
    from oracle import nmf_oracle as o
    X, A, _ = o.gen(91, 129, 300, 10)
    dictionary_cache.clear()
    h0, m0 = dictionary_cache.hits, dictionary_cache.misses
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        W1, _, _ = nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="mu",
                                                  beta_loss="kullback-leibler", max_iter=20, tol=0)
        W2, _, _ = nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="mu",
                                                  beta_loss="kullback-leibler", max_iter=20, tol=0)
        assert (dictionary_cache.hits - h0, dictionary_cache.misses - m0) == (1, 1)
        assert np.array_equal(W1, W2)
        A *= 2.0                                               # in place: same object, different content
        W3, _, _ = nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="mu",
                                                  beta_loss="kullback-leibler", max_iter=20, tol=0)
        assert dictionary_cache.misses - m0 == 2
    W3_ref, _, _ = o.kl_mu(X, A, tol=0, max_iter=20)
    assert rel_fro(W3, W3_ref) < 1e-3 and rel_fro(W3, W1) > 1e-2
    dictionary_cache.clear()


def test_dtw_alignment_on_the_device_matches_the_oracle():
    """features.dtw_alignment (drop-in for 01_make_dict_parallel.dtw_alignment, :239-249): ragged files, the
    reference's (order, n_frames) layout, paths and distances bit-identical to the restated `dtw` recursion."""
    from exemplars_vc_b200 import features
    from oracle import dtw_oracle as o
    rng = np.random.default_rng(31)
    lens = [(37, 41), (1, 9), (64, 64), (120, 77), (5, 1)]
    A = [np.cumsum(rng.standard_normal((24, r)), axis=1) for r, _ in lens]        # (order, n_frames)
    B = [np.cumsum(rng.standard_normal((24, c)), axis=1) for _, c in lens]
    A[2] = B[2].copy()                                                              # identical pair: pure diagonal, all ties
    paths, none1, none2 = features.dtw_alignment(A, B)
    assert none1 is None and none2 is None and len(paths) == len(lens)
    for i, (a, b) in enumerate(zip(A, B)):
        dist, _, _, (p, q) = o.dtw(a.T, b.T)
        assert np.array_equal(paths[i][0], p) and np.array_equal(paths[i][1], q), i
        assert features.dtw_alignment.last_distances[i] == dist
    assert np.array_equal(paths[2][0], np.arange(64)) and np.array_equal(paths[2][1], np.arange(64))
    # the paths index the per-file features exactly as make_exemplar_dict does (:205-206) and feed build_dictionaries
    src = [np.abs(a.T).astype(np.float32) for a in A]
    tar = [np.abs(b.T).astype(np.float32) for b in B]
    d = features.build_dictionaries(src, tar, [p for p, _ in paths], [q for _, q in paths], mode="fp32")
    assert d.N == sum(len(p) for p, _ in paths)
    d.close()
    with pytest.raises(ValueError):
        features.dtw_alignment(A, B[:-1])


def test_dtw_alignment_at_utterance_length():
    """Two ~7 s utterances at a 5 ms hop (1 400 frames, 25 coefficients), the largest files the reference ships."""
    from exemplars_vc_b200 import features
    from oracle import dtw_oracle as o
    rng = np.random.default_rng(32)
    a = np.cumsum(rng.standard_normal((25, 1380)), axis=1)
    warp = np.clip(np.arange(1290) * 1380 // 1290 + rng.integers(-3, 4, 1290), 0, 1379)
    b = a[:, np.sort(warp)] + 0.05 * rng.standard_normal((25, 1290))
    (p, q), = features.dtw_alignment([a], [b])[0]
    C = o.local_cost(a.T, b.T)
    # the accumulated cost along the device path equals the optimum of a vectorised numpy recursion
    D = np.full((1381, 1291), np.inf)
    D[0, 0] = 0.0
    for i in range(1, 1381):
        row = D[i]
        up_diag = np.minimum(D[i - 1, 1:], D[i - 1, :-1])
        for j in range(1, 1291):
            row[j] = C[i - 1, j - 1] + min(up_diag[j - 1], row[j - 1])
    assert (p[0], q[0], p[-1], q[-1]) == (0, 0, 1379, 1289)
    assert np.isclose(C[p, q].sum(), D[-1, -1], rtol=1e-12)
    assert features.dtw_alignment.last_distances[0] == pytest.approx(D[-1, -1] / (1380 + 1290), rel=1e-12)
