"""The oracle against the reference's own outputs (CPU).  Pins the checker before it is trusted:
tests/golden/*.npz were produced by the reference's exact scikit-learn call (oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden_inputs, load_golden, rel_fro
from oracle import nmf_oracle as o

KL_CASES = ["kl_13x32x8_tol1e-2", "kl_13x32x8_tol1e-3", "kl_13x32x8_tol1e-4", "kl_13x32x8_tol0_500",
            "kl_13x32x8_f32", "kl_513x2000x64_tol1e-4", "kl_201x777x37_tol1e-4", "kl_edge_zeros", "kl_edge_f0"]


@pytest.mark.parametrize("name", KL_CASES)
def test_kl_restatement_matches_reference_output(name):
    g = load_golden(name)
    X, A, B = golden_inputs(g)
    W, n_iter, obj = o.kl_mu(X, A, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert n_iter == int(g["n_iter"])
    # same BLAS, same order of operations: identical up to the last bits
    np.testing.assert_allclose(W, g["W"], rtol=1e-9 if X.dtype == np.float64 else 1e-5, atol=0)
    np.testing.assert_allclose(obj, float(g["objective"]), rtol=1e-7 if X.dtype == np.float64 else 1e-3)
    np.testing.assert_allclose(o.convert(W, B), g["Y"], rtol=1e-9 if X.dtype == np.float64 else 1e-5)


def test_survey_known_answer_table():
    """values recorded during the survey from sklearn 1.9.0 (float64, KL, lambda = 0)."""
    X, A, B = o.gen(20190123, 13, 32, 8)
    assert abs(X.sum() - 2.3239120916e+02) < 1e-6 and abs(A.sum() - 4.3037000814e+02) < 1e-6
    W0 = o.initial_activation(X, 32)
    assert abs(W0[0, 0] - 0.26425194283181025) < 1e-15
    assert abs(o.kl_objective(X, W0, A) - 28.552161128782707) < 1e-10
    table = [(1e-2, 150, 40, 9.1014002625e-01, 1.6956426558e+01), (1e-3, 150, 120, 3.7807013311e-01, 1.7169381822e+01),
             (1e-4, 150, 150, 3.1988167812e-01, 1.7198190310e+01), (0.0, 500, 500, 1.4890424243e-01, 1.7277834608e+01)]
    for tol, mi, n_exp, obj_exp, sum_exp in table:
        W, n, obj = o.kl_mu(X, A, tol=tol, max_iter=mi)
        assert n == n_exp
        assert abs(obj - obj_exp) / obj_exp < 1e-7
        assert abs(W.sum() - sum_exp) / sum_exp < 1e-7
    W, n, obj = o.kl_mu(X, A, tol=1e-4, max_iter=150)
    assert abs(W[0, 0] - 8.7677399402e-02) / 8.7677399402e-02 < 1e-7
    assert abs(np.linalg.norm(W) - 2.5967513177e+00) / 2.5967513177e+00 < 1e-7
    assert abs(np.linalg.norm(o.convert(W, B)) - 2.6466200731e+01) / 2.6466200731e+01 < 1e-7


def test_survey_table_513():
    g = load_golden("kl_513x2000x64_tol0_500")
    assert int(g["n_iter"]) == 500
    assert abs(float(g["objective"]) - 8.6895056941e-01) / 8.6895056941e-01 < 1e-7
    assert abs(float(g["sum_W"]) - 1.6001907294e+02) / 1.6001907294e+02 < 1e-7
    assert abs(float(g["norm_Y"]) - 5.7751667606e+02) / 5.7751667606e+02 < 1e-7


@pytest.mark.parametrize("name", ["fro_13x32x8_tol1e-4", "fro_201x777x37_tol1e-4"])
def test_frobenius_restatement(name):
    g = load_golden(name)
    X, A, B = golden_inputs(g)
    W, n_iter, obj = o.frobenius_mu(X, A, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert n_iter == int(g["n_iter"])
    np.testing.assert_allclose(W, g["W"], rtol=1e-9)
    np.testing.assert_allclose(obj, float(g["objective"]), rtol=1e-9)


def test_l1_quirk_q1_and_constant_lambda():
    g = load_golden("kl_l1_sklearn_q1")
    X, A = g["X"], g["A"]
    lam = X.shape[1] * float(g["alpha_W"]) * float(g["l1_ratio"])
    W_acc, n, _ = o.kl_mu(X, A, lam=lam, tol=0.0, max_iter=int(g["max_iter"]), sklearn_l1_accumulate=True)
    np.testing.assert_allclose(W_acc, g["W"], rtol=1e-9)       # sklearn's denominator grows by lam each iteration
    W_const, _, _ = o.kl_mu(X, A, lam=lam, tol=0.0, max_iter=int(g["max_iter"]))
    assert rel_fro(W_const, g["W"]) > 1e-2                      # ... the north-star constant penalty does not
    gc = load_golden("kl_l1_constant")
    np.testing.assert_allclose(W_const, gc["W"], rtol=1e-12)


def test_reference_call_is_the_restatement():
    pytest.importorskip("sklearn")
    X, A, _ = o.gen(77, 29, 50, 11)
    W_ref, n_ref = o.reference_call(X, A, tol=1e-4, max_iter=60)
    W, n, _ = o.kl_mu(X, A, tol=1e-4, max_iter=60)
    assert n == n_ref and np.array_equal(W, W_ref)
    W_ref, n_ref = o.reference_call(X, A, beta_loss="frobenius", tol=1e-4, max_iter=60)
    W, n, _ = o.frobenius_mu(X, A, tol=1e-4, max_iter=60)
    assert n == n_ref and np.array_equal(W, W_ref)


def test_objective_masks_small_x():
    X, A, _ = o.gen(5, 7, 9, 4)
    X[0, 0] = 0.0
    X[1, 2] = 1e-9          # below float32 eps: excluded from X log(X/WH) and from -sum X
    W = o.initial_activation(X, 9)
    WH = W @ A
    m = X > o.EPSILON
    expect = np.sqrt(2 * (np.sum(X[m] * np.log(X[m] / WH[m])) - X[m].sum() + WH.sum()))
    assert abs(o.kl_objective(X, W, A) - expect) < 1e-12


def test_nmf_tool_restatement_decreases_cost():
    rng = np.random.default_rng(3)
    W = rng.random((12, 5)); V = W @ rng.random((5, 7)); H0 = rng.random((5, 7))
    costs = [o.nmf_tool_euclidean_mu(V, W, H0, k)[1] for k in (0, 1, 5, 50)]
    assert all(b <= a + 1e-12 for a, b in zip(costs, costs[1:])) and costs[-1] < 1e-1 * costs[0]


def test_synth_generator_matches_oracle_gen():
    from exemplars_vc_b200 import synth
    X, A, B = o.gen(123, 11, 40, 6, np.float32)
    A2, B2 = synth.dictionaries(123, 11, 40)
    assert np.array_equal(A, A2) and np.array_equal(B, B2)
    A64, _ = synth.dictionaries(123, 11, 40, np.float64)
    X2 = synth.frames(123, A64, 6, np.float32, exact=True)
    assert np.array_equal(X, X2)
    X3 = synth.frames(123, A2, 6)
    assert X3.shape == X.shape and X3.min() > 0


def test_real_speech_fixture_reduced_config1():
    """BASELINE.json configs[0] in reduced form (oracle/make_golden_speech.py): spectra of the reference's own
    wav files, 768 DTW-aligned exemplar pairs, 64 frames of utterance 100162."""
    g = load_golden("speech_sf1_tf1_100162")
    X, A, B = (g[k].astype(np.float64) for k in ("X", "A", "B"))
    W, n_iter, obj = o.kl_mu(X, A, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert n_iter == int(g["n_iter"])
    assert rel_fro(W, g["W"]) < 1e-6            # golden W is stored in float32
    assert abs(obj - float(g["objective"])) / float(g["objective"]) < 1e-9
    assert rel_fro(o.convert(W, B), g["Y"]) < 1e-6
    assert (W < 1e-6 * W.max()).mean() > 0.5     # activations of real speech over exemplars are sparse


def test_griffin_lim_oracle_is_bit_identical_to_the_reference_functions():
    """oracle/griffin_lim_oracle.py against outputs of the reference's own zz_audio_utilities.py functions
    (oracle/make_golden_griffin_lim.py: the unmodified file imported with a numpy stand-in for pylab)."""
    from conftest import load_golden
    from oracle import griffin_lim_oracle as g
    z = load_golden("griffin_lim_400_80")
    fft, hop = int(z["fft_size"]), int(z["hop"])
    S = g.stft_for_reconstruction(z["sig"], fft, hop)
    assert np.array_equal(S.real, z["stft_re"]) and np.array_equal(S.imag, z["stft_im"])
    assert np.array_equal(g.istft_for_reconstruction(S, fft, hop), z["istft"])
    for it in (1, 3, 30):
        x = g.reconstruct_signal_griffin_lim(z["mag"].astype(np.float64), fft, hop, it, z["x0"])
        assert np.array_equal(x, z["x%d" % it])


def test_dtw_oracle_properties_and_known_answer():
    """oracle/dtw_oracle.py (restatement of the un-vendored `dtw` package's recursion, parity unpinned -- see its
    header): endpoints, monotone unit steps, optimality against brute force on a tiny case, the hand-checkable
    known answer of two shifted ramps, and the first-minimum tie-break (diagonal, then up, then left)."""
    import itertools
    from oracle import dtw_oracle as d
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal((7, 3)), rng.standard_normal((5, 3))
    dist, C, D1, (p, q) = d.dtw(x, y)
    assert (p[0], q[0]) == (0, 0) and (p[-1], q[-1]) == (6, 4) and len(p) == len(q)
    steps = set(zip(np.diff(p).tolist(), np.diff(q).tolist()))
    assert steps <= {(1, 1), (1, 0), (0, 1)}
    assert np.isclose(C[p, q].sum(), D1[-1, -1]) and np.isclose(dist, D1[-1, -1] / 12)
    assert np.array_equal(C, np.array([[sum(np.square(a - b)) for b in y] for a in x]))      # the lambda of :226, bit for bit
    # brute force over all monotone paths of a 4 x 3 problem
    x, y = rng.standard_normal((4, 2)), rng.standard_normal((3, 2))
    _, C, D1, _ = d.dtw(x, y)

    def best(i, j):
        if i == 0 and j == 0:
            return C[0, 0]
        cands = []
        if i > 0 and j > 0: cands.append(best(i - 1, j - 1))
        if i > 0: cands.append(best(i - 1, j))
        if j > 0: cands.append(best(i, j - 1))
        return C[i, j] + min(cands)
    assert np.isclose(D1[-1, -1], best(3, 2))
    # y = x with its first frame held twice more: the zero-cost path waits on x[0], then runs along the diagonal
    ramp = np.arange(6, dtype=float)[:, None]
    dist, _, _, (p, q) = d.dtw(ramp, np.concatenate([np.zeros((2, 1)), ramp]))
    assert dist == 0.0 and np.array_equal(p, [0, 0, 0, 1, 2, 3, 4, 5]) and np.array_equal(q, np.arange(8))
    # ties: identical constant sequences -> every step is the diagonal (first minimum)
    _, _, _, (p, q) = d.dtw(np.ones((4, 2)), np.ones((4, 2)))
    assert np.array_equal(p, np.arange(4)) and np.array_equal(q, np.arange(4))
