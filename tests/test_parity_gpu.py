"""Parity of the CUDA path with the reference's outputs (golden vectors made by the reference's own
scikit-learn call, oracle/make_golden.py) -- through the Python operator and the C ABI behind it.

Tolerances (BASELINE.json north_star):
  fp32-accurate modes ("fp32" FFMA and "3xtf32" tensor cores): final H and Y within 1e-3 relative Frobenius
      error of the float64 reference, KL objective within 1e-4 relative, n_iter identical.
  fast mode "tf32" (single-pass TF32): H and Y within 2e-2, objective within 2e-2 (stated here, not claimed
      to be fp32-accurate); n_iter may differ when the stop rule sits on a knife edge, so it is not asserted.
  fast mode "bf16" (bf16 operand copies, fp32 accumulation and fp32 multiplicative update): H and Y within 2e-2,
      objective within 3e-2 at realistic shapes (F >= 201; a numpy emulation of the same rounding gives 1.2e-3 /
      2e-3 / 9e-3); the 13x32x8 toy cases only within 1e-1 / 5e-2 (8-bit mantissas against a tiny residual).
"""
import warnings

import numpy as np
import pytest

from conftest import golden_inputs, load_golden, rel_fro

pytestmark = pytest.mark.gpu

ACCURATE = ["fp32", "3xtf32"]
TOL = {"fp32": (1e-3, 1e-4), "3xtf32": (1e-3, 1e-4), "tf32": (2e-2, 2e-2), "bf16": (2e-2, 3e-2)}
KL_CASES = ["kl_13x32x8_tol1e-2", "kl_13x32x8_tol1e-3", "kl_13x32x8_tol1e-4", "kl_13x32x8_tol0_500",
            "kl_513x2000x64_tol1e-4", "kl_201x777x37_tol1e-4"]


def _solve(mode, X, A, B, **kw):
    from exemplars_vc_b200 import ExemplarDictionary
    with ExemplarDictionary(A, B, mode=mode) as d:
        act = d.solve(X, **kw)
        H = d.to_host(act.H)
        Y = d.to_host(d.convert(act.H))
    return act, H, Y


@pytest.mark.parametrize("mode", ["fp32", "3xtf32", "tf32"])
@pytest.mark.parametrize("name", KL_CASES)
def test_kl_matches_reference_golden(name, mode):
    g = load_golden(name)
    X, A, B = golden_inputs(g)
    act, H, Y = _solve(mode, X, A, B, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    tol_h, tol_obj = TOL[mode]
    if mode in ACCURATE:
        assert act.n_iter == int(g["n_iter"])
    assert H.min() >= 0.0
    assert rel_fro(H, g["W"]) < tol_h, (rel_fro(H, g["W"]), act.n_iter, int(g["n_iter"]))
    assert rel_fro(Y, g["Y"]) < tol_h
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < tol_obj
    assert abs(act.objective_at_init - float(g["objective_at_init"])) / float(g["objective_at_init"]) < tol_obj


@pytest.mark.parametrize("name", ["kl_513x2000x64_tol1e-4", "kl_201x777x37_tol1e-4", "kl_13x32x8_tol1e-2"])
def test_bf16_fast_mode_kl(name):
    """EVC_MODE_BF16: tcgen05 kind::f16 on bf16 copies of A, of the ratio and of H (shadow written by the fused
    update); tolerances stated in the module docstring."""
    g = load_golden(name)
    X, A, B = golden_inputs(g)
    act, H, Y = _solve("bf16", X, A, B, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    tol_h, tol_obj = (1e-1, 5e-2) if name.startswith("kl_13x") else TOL["bf16"]
    assert H.min() >= 0.0 and np.isfinite(H).all()
    assert rel_fro(H, g["W"]) < tol_h, rel_fro(H, g["W"])
    assert rel_fro(Y, g["Y"]) < tol_h, rel_fro(Y, g["Y"])
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < tol_obj
    assert abs(act.objective_at_init - float(g["objective_at_init"])) / float(g["objective_at_init"]) < tol_obj


def test_bf16_fast_mode_frobenius_and_given_init():
    g = load_golden("fro_201x777x37_tol1e-4")
    X, A, B = golden_inputs(g)
    act, H, _ = _solve("bf16", X, A, B, beta_loss="frobenius", tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert rel_fro(H, g["W"]) < 2e-2, rel_fro(H, g["W"])
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < 3e-2
    # reconstruct / objective of a GIVEN H go through the bf16 shadow made on entry
    from exemplars_vc_b200 import ExemplarDictionary
    with ExemplarDictionary(A, B, mode="bf16") as d:
        W = np.ascontiguousarray(g["W"], np.float32)
        WH = d.to_host(d.reconstruct(W))
        assert rel_fro(WH, W.astype(np.float64) @ A.astype(np.float64)) < 1e-2


@pytest.mark.parametrize("mode", ACCURATE)
def test_kl_500_iterations_summary(mode):
    """SURVEY 8(c) table row 513,2000,64, tol=0, 500 iterations (summary scalars only)."""
    g = load_golden("kl_513x2000x64_tol0_500")
    X, A, B = golden_inputs(g)
    act, H, Y = _solve(mode, X, A, B, tol=0.0, max_iter=500)
    assert act.n_iter == 500
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < 1e-4
    assert abs(H.sum(dtype=np.float64) - float(g["sum_W"])) / float(g["sum_W"]) < 1e-4
    assert abs(np.linalg.norm(H.astype(np.float64)) - float(g["norm_W"])) / float(g["norm_W"]) < 1e-3
    assert abs(np.linalg.norm(Y.astype(np.float64)) - float(g["norm_Y"])) / float(g["norm_Y"]) < 1e-3


@pytest.mark.parametrize("mode", ACCURATE)
@pytest.mark.parametrize("name", ["fro_13x32x8_tol1e-4", "fro_201x777x37_tol1e-4"])
def test_frobenius_matches_reference_golden(name, mode):
    """What 04_align_n_nmf.py:210 really runs.  The GPU forms A^T(A H) instead of sklearn's N x N Gram."""
    g = load_golden(name)
    X, A, B = golden_inputs(g)
    act, H, Y = _solve(mode, X, A, B, beta_loss="frobenius", tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert act.n_iter == int(g["n_iter"])
    assert rel_fro(H, g["W"]) < 1e-3
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < 1e-4


@pytest.mark.parametrize("mode", ACCURATE)
def test_edge_zeros_in_frames_and_dead_exemplars(mode):
    """X with exact zeros (masked out of the objective) and all-zero exemplars (den == 0 -> eps)."""
    g = load_golden("kl_edge_zeros")
    act, H, Y = _solve(mode, g["X"], g["A"], g["B"], tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert act.n_iter == int(g["n_iter"])
    assert np.all(H[:, 3] == 0.0) and np.all(H[:, 11] == 0.0) and np.all(g["W"][:, 3] == 0.0)
    assert rel_fro(H, g["W"]) < 1e-3 and rel_fro(Y, g["Y"]) < 1e-3
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < 1e-4


def test_edge_f0_track_single_feature():
    """F = 1 with unvoiced zeros (04_align_n_nmf.py:288): routed to the FFMA kernels."""
    g = load_golden("kl_edge_f0")
    act, H, Y = _solve("fp32", g["X"], g["A"], g["B"], tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    assert act.n_iter == int(g["n_iter"])
    assert rel_fro(Y, g["Y"]) < 1e-3
    assert rel_fro(H, g["W"]) < 1e-3


@pytest.mark.parametrize("mode", ACCURATE)
def test_l1_penalty_constant_and_sklearn_accumulating(mode):
    gq = load_golden("kl_l1_sklearn_q1")
    lam = gq["X"].shape[1] * float(gq["alpha_W"]) * float(gq["l1_ratio"])
    act, H, _ = _solve(mode, gq["X"], gq["A"], gq["B"], tol=0.0, max_iter=int(gq["max_iter"]), lam=0.0, lambda_step=lam)
    assert rel_fro(H, gq["W"]) < 1e-3
    gc = load_golden("kl_l1_constant")
    act, H, _ = _solve(mode, gc["X"], gc["A"], gc["B"], tol=0.0, max_iter=int(gc["max_iter"]), lam=float(gc["lam"]))
    assert rel_fro(H, gc["W"]) < 1e-3
    assert abs(act.objective - float(gc["objective"])) / float(gc["objective"]) < 1e-4


def test_operator_signature_and_conventions():
    """non_negative_factorization(X=X, H=W, init='custom', update_H=False, ...) exactly as 04_align_n_nmf.py:212."""
    from exemplars_vc_b200 import ConvergenceWarning, non_negative_factorization
    g = load_golden("kl_13x32x8_tol1e-4")
    X, A, B = golden_inputs(g)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        _W, _H, n_iter = non_negative_factorization(X=X, H=A, init="custom", update_H=False, n_components=A.shape[0],
                                                    beta_loss="kullback-leibler", solver="mu", tol=1e-4, max_iter=150,
                                                    verbose=0)
    assert _H is A and _W.dtype == X.dtype and _W.shape == (8, 32) and n_iter == 150
    assert any(issubclass(x.category, ConvergenceWarning) for x in w)      # n_iter == max_iter and tol > 0
    assert rel_fro(_W, g["W"]) < 1e-3
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        non_negative_factorization(X=X, W=np.ones((8, 32)), H=A, init="custom", update_H=False, solver="mu",
                                   beta_loss="kullback-leibler", max_iter=5)
    assert any(issubclass(x.category, RuntimeWarning) for x in w)
    with pytest.raises(ValueError, match="Negative values"):
        An = A.copy(); An[2, 1] = -0.5
        non_negative_factorization(X=X, H=An, init="custom", update_H=False, solver="mu", beta_loss="kullback-leibler")
    with pytest.raises(ValueError, match="full of zeros"):
        non_negative_factorization(X=X, H=np.zeros_like(A), init="custom", update_H=False, solver="mu",
                                   beta_loss="kullback-leibler")
    Xf = X.astype(np.float32)
    Wf, _, _ = non_negative_factorization(X=Xf, H=A.astype(np.float32), init="custom", update_H=False, solver="mu",
                                          beta_loss="kullback-leibler", tol=0, max_iter=20)
    assert Wf.dtype == np.float32


def test_script_level_factorize_and_convert():
    """_factorize / factorize / convert of 04_align_n_nmf.py on the STFT branch and the WORLD branch."""
    from exemplars_vc_b200 import align_n_nmf as m
    from oracle import nmf_oracle as o
    X, A, B = o.gen(31, 201, 300, 20)
    files = [slice(0, 100), slice(100, 180), slice(180, 300)]
    src = [{"real": A[s] * np.where(np.arange(201) % 2, -1, 1)} for s in files]      # abs() is taken inside
    tar = [{"real": -B[s]} for s in files]
    m.use_stft = 1
    H, R = m.factorize({"real": -X}, src)
    assert R is None and H["H_stft"].shape == (300, 20)
    W_ref, n, _ = o.kl_mu(X, A, tol=1e-4, max_iter=150)
    assert rel_fro(H["H_stft"].T, W_ref) < 1e-3
    Y = m.convert(H, tar, None)
    assert rel_fro(Y, W_ref @ B) < 1e-3
    # literal reference body: beta forced to frobenius (04_align_n_nmf.py:210)
    m.beta_override = "frobenius"
    try:
        Hf = m._factorize(X, A)
    finally:
        m.beta_override = None
    Wf, _, _ = o.frobenius_mu(X, A, tol=1e-4, max_iter=150)
    assert rel_fro(Hf.T, Wf) < 1e-3
    # WORLD branch: sp, ap (F=33 here) and the F=1 f0 track with unvoiced zeros
    Xs, As, Bs = o.gen(41, 33, 120, 12)
    Xa, Aa, Ba = o.gen(42, 33, 120, 12)
    rng = np.random.default_rng(1)
    f0d = np.where(rng.random(120) < 0.3, 0.0, 100 + 100 * rng.random(120))
    f0t = np.where(rng.random(120) < 0.3, 0.0, 150 + 100 * rng.random(120))
    f0x = np.where(rng.random(12) < 0.3, 0.0, 100 + 100 * rng.random(12))
    m.use_stft = 0
    try:
        H, R = m.factorize({"sp": Xs, "ap": Xa, "f0": f0x},
                           [{"sp": As[:50], "ap": Aa[:50], "f0": f0d[:50]}, {"sp": As[50:], "ap": Aa[50:], "f0": f0d[50:]}])
        assert set(H) == {"H_sp", "H_ap", "H_f0"} and set(R) == {"r_sp", "r_ap", "r_f0"}
        assert rel_fro(H["H_sp"].T, o.kl_mu(Xs, As)[0]) < 1e-3
        assert rel_fro(H["H_f0"].T, o.kl_mu(f0x[:, None], f0d[:, None])[0]) < 1e-3
        out = m.convert(H, [{"sp": Bs, "ap": Ba, "f0": f0t}], R)
        assert out["sp"].shape == (12, 33) and out["f0"].shape == (12,)
    finally:
        m.use_stft = 1


def test_nmf_tool_fixed_dictionary():
    """nmf_tool.nmf.NMF with initW=True: Euclidean MU from the same H0 (nmf_tool/nmf.py:38-40)."""
    from exemplars_vc_b200.nmf_tool.nmf import NMF
    from oracle import nmf_oracle as o
    rng = np.random.default_rng(8)
    W = rng.random((40, 25)).astype(np.float32)
    V = (W @ rng.random((25, 9))).astype(np.float32)
    H0 = rng.random((25, 9)).astype(np.float32)
    model = NMF(max_iter=60, display_step=0, optimizer="mu", mode="fp32")
    W_out, H = model.fit_transform(V, r_components=25, initW=True, givenW=W, H0=H0)
    H_ref, cost = o.nmf_tool_euclidean_mu(V.astype(np.float64), W.astype(np.float64), H0.astype(np.float64), 60)
    assert W_out.shape == (40, 25) and H.shape == (25, 9)
    assert rel_fro(H, H_ref) < 1e-3
    assert rel_fro(model.inverse_transform(W_out, H), W @ H_ref) < 1e-3


@pytest.mark.parametrize("mode", ACCURATE)
def test_batched_utterances_equal_separate_calls(mode):
    """Stacked-T mode with per-utterance H0 and stop rule == one reference call per utterance; includes an
    empty utterance and ragged lengths."""
    from exemplars_vc_b200 import ExemplarDictionary
    from oracle import nmf_oracle as o
    _, A, B = o.gen(51, 40, 150, 1)
    lens = [7, 0, 33, 1, 12]
    utts = [o.gen(60 + i, 40, 150, max(L, 1))[0][:L] for i, L in enumerate(lens)]
    # frames must be mixtures of THIS dictionary: rebuild them from A
    rng = np.random.default_rng(3)
    utts = [(rng.random((L, 150)) * (rng.random((L, 150)) < 0.05)) @ A + 0.01 * rng.random((L, 40)) for L in lens]
    offs = np.concatenate([[0], np.cumsum(lens)]).tolist()
    with ExemplarDictionary(A, B, mode=mode) as d:
        acts = d.solve_batched(np.concatenate(utts, 0), offs, tol=1e-3, max_iter=150, per_utterance_stop=True)
        Hs = [d.to_host(a.H) for a in acts]
    n_iters = []
    for i, L in enumerate(lens):
        if L == 0:
            assert Hs[i].shape == (0, 150)
            continue
        W_ref, n_ref, obj = o.kl_mu(utts[i], A, tol=1e-3, max_iter=150)
        n_iters.append(n_ref)
        assert acts[i].n_iter == n_ref, (i, acts[i].n_iter, n_ref)
        assert rel_fro(Hs[i], W_ref) < 1e-3
        assert abs(acts[i].objective - obj) / obj < 1e-4
    assert len(set(n_iters)) > 1          # the utterances really stopped at different iterations


def test_host_buffer_c_abi_entry_point():
    """evc_factorize_convert_host: the plain-C consumer path (host pointers in, host pointers out)."""
    import ctypes as C
    from exemplars_vc_b200 import ExemplarDictionary, _lib
    g = load_golden("kl_201x777x37_tol1e-4")
    X, A, B = golden_inputs(g)
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    H = np.zeros((37, 777), dtype=np.float32)
    Y = np.zeros((37, 201), dtype=np.float32)
    with ExemplarDictionary(A, B, mode="3xtf32") as d:
        p = _lib.SolveParams()
        _lib.lib().evc_default_params(C.byref(p))
        res = _lib.SolveResult()
        _lib.check(_lib.lib().evc_factorize_convert_host(d._h, X32.ctypes.data, 201, 37, H.ctypes.data, 777,
                                                         Y.ctypes.data, 201, C.byref(p), C.byref(res), None))
    assert res.n_iter == int(g["n_iter"])
    assert rel_fro(H, g["W"]) < 1e-3 and rel_fro(Y, g["Y"]) < 1e-3


@pytest.mark.parametrize("mode", ["fp32", "3xtf32", "tf32", "bf16"])
def test_real_speech_reduced_config1(mode):
    """BASELINE.json configs[0] in reduced form: real spectra (60 dB dynamic range, sparse activations)."""
    g = load_golden("speech_sf1_tf1_100162")
    act, H, Y = _solve(mode, g["X"], g["A"], g["B"], tol=float(g["tol"]), max_iter=int(g["max_iter"]))
    tol_h, tol_obj = TOL[mode]
    if mode in ACCURATE:
        assert act.n_iter == int(g["n_iter"])
    assert rel_fro(H, g["W"]) < tol_h, rel_fro(H, g["W"])
    assert rel_fro(Y, g["Y"]) < tol_h, rel_fro(Y, g["Y"])
    assert abs(act.objective - float(g["objective"])) / float(g["objective"]) < tol_obj


def test_device_dictionary_gather_and_context_stacking():
    """aligned-frame gather (04_align_n_nmf.py:113-124, 230-246) + +-2 frame stacking on the device."""
    from exemplars_vc_b200 import features
    rng = np.random.default_rng(12)
    F = 37
    src = [rng.random((n, F)).astype(np.float32) for n in (11, 5, 23)]
    tar = [rng.random((n, F)).astype(np.float32) for n in (9, 8, 20)]
    sp = [np.sort(rng.integers(0, len(s), size=14)) for s in src]
    tp = [np.sort(rng.integers(0, len(t), size=14)) for t in tar]

    def ref(files, paths, c):
        rows = []
        for m, p in zip(files, paths):
            for k in p:
                rows.append(np.concatenate([m[min(max(k + d, 0), len(m) - 1)] for d in range(-c, c + 1)]))
        return np.asarray(rows)

    for c in (0, 2):
        d = features.build_dictionaries(src, tar, sp, tp, context=c, mode="fp32")
        assert (d.N, d.F) == (42, (2 * c + 1) * F)
        A_ref, B_ref = ref(src, sp, c), ref(tar, tp, c)
        H = np.eye(42, dtype=np.float32)
        assert np.array_equal(d.to_host(d.reconstruct(H)), A_ref)      # identity activations read the dictionary back
        assert np.array_equal(d.to_host(d.convert(H)), B_ref)
        d.close()
    X = rng.random((6, F)).astype(np.float32)
    Xs = features.stack_frames(X, 2).cpu().numpy()
    assert np.array_equal(Xs, ref([X], [np.arange(6)], 2))
    # reference-style per-file dicts with a key, |real(stft)| taken as in 04_align_n_nmf.py:323
    d = features.build_dictionaries([{"real": -s} for s in src], [{"real": t} for t in tar], sp, tp, key="real", mode="fp32")
    assert np.array_equal(d.to_host(d.reconstruct(np.eye(42, dtype=np.float32))), ref(src, sp, 0))
    d.close()
