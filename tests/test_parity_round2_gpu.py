"""Round-2 parity cases: the headline shape for its full 500 iterations against the
float64 oracle, BASELINE configs[0] at its real size from the reference's own wav files, run-to-run bit identity
at the full shape (the race evidence compute-sanitizer cannot give on this pool), and EXECUTED tests of the
script-level entry points that round 1 only checked by signature (05_conversion.py:94-107,
04_align_n_nmf_pytorch.py:213-327, nmf_tool in the tensor-core mode, the content-keyed cache).

Tolerances.  fp32-accurate mode ("3xtf32": three-term bf16 split products, fp32 accumulate): H, Y < 1e-3 relative
Frobenius, objective < 1e-4 relative -- the north star's.  Fast modes AT THE HEADLINE SHAPE after 500 iterations
(stated here, measured values in profiles/accuracy_modes_r2.log): "tf32" H, Y < 3e-2, objective < 3e-2;
"bf16" H, Y < 1.5e-1, objective < 3e-1 (the synthetic frames are an almost exact model, so the converged residual
is tiny against 8-bit mantissas: bf16 is a preview mode, not a drop-in for the accurate one).
"""
import os
import pickle
import warnings

import numpy as np
import pytest

from conftest import load_golden, rel_fro

pytestmark = pytest.mark.gpu

HEADLINE_TOL = {"3xtf32": (1e-3, 1e-4), "tf32": (3e-2, 3e-2), "bf16": (1.5e-1, 3e-1)}


@pytest.fixture(scope="module")
def headline():
    """F=513, N=20000, T=1000 (BASELINE configs[1], bench.py's seed) and the float64 oracle on 32 evenly spaced frames
    for 500 iterations, started from the H0 value of the FULL problem (frames are independent given H0)."""
    from exemplars_vc_b200 import synth
    from oracle import nmf_oracle as o
    wl = synth.CONFIGS["single_utterance_20k"]
    seed = synth.BASE_SEED + 1
    A, B = synth.dictionaries(seed, wl.F, wl.N)
    X = synth.frames(seed, A, wl.T)
    idx = np.linspace(0, wl.T - 1, 32).astype(int)
    w0 = np.float32(np.sqrt(X.mean(dtype=np.float64) / wl.N))      # sklearn _nmf.py:1225-1226 on the whole utterance
    A64, X64 = A.astype(np.float64), X[idx].astype(np.float64)
    W_ref, n_it, obj_ref = o.kl_mu(X64, A64, tol=0.0, max_iter=wl.iterations, W0=np.full((32, wl.N), float(w0)))
    return dict(A=A, B=B, X=X, idx=idx, W_ref=W_ref, obj_ref=obj_ref, Y_ref=W_ref @ B.astype(np.float64),
                iterations=wl.iterations)


@pytest.mark.parametrize("mode", ["3xtf32", "tf32", "bf16"])
def test_headline_shape_500_iterations_vs_oracle(headline, mode):
    from exemplars_vc_b200 import ExemplarDictionary
    h = headline
    with ExemplarDictionary(h["A"], h["B"], mode=mode) as d:
        act = d.solve(h["X"], tol=0.0, max_iter=h["iterations"])
        assert act.n_iter == h["iterations"]
        Hs = act.H[h["idx"].tolist()].contiguous()
        H = d.to_host(Hs)
        Y = d.to_host(d.convert(Hs))
        obj = d.objective(h["X"][h["idx"]], Hs)
    eh, ey = rel_fro(H, h["W_ref"]), rel_fro(Y, h["Y_ref"])
    eo = abs(obj - h["obj_ref"]) / h["obj_ref"]
    print(f"headline[{mode}] 500 iterations, 32 frames: relF(H)={eh:.2e} relF(Y)={ey:.2e} rel(obj)={eo:.2e}")
    tol_h, tol_o = HEADLINE_TOL[mode]
    assert eh < tol_h and ey < tol_h and eo < tol_o, (mode, eh, ey, eo)
    assert np.isfinite(H).all() and H.min() >= 0.0


@pytest.mark.parametrize("mode", ["3xtf32", "fp32"])
def test_config1_full_size_real_speech(mode):
    """BASELINE configs[0]: SF1 -> TF1, utterance 100162, dictionary from all 8 parallel pairs the reference ships
    (6821 DTW-aligned exemplar pairs), the whole utterance (675 frames at this framing), reference defaults
    (max_iter 150, tol 1e-4).  Golden = the reference's own scikit-learn call (oracle/make_golden_speech.py --full)."""
    from exemplars_vc_b200 import ExemplarDictionary
    g = load_golden("speech_sf1_tf1_100162_full")
    X, A, B = (g[k].astype(np.float32) for k in ("X16", "A16", "B16"))
    with ExemplarDictionary(A, B, mode=mode) as d:
        act = d.solve(X, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
        H = d.to_host(act.H)
        Y = d.to_host(d.convert(act.H))
    assert act.n_iter == int(g["n_iter"])
    eh, ey = rel_fro(H[g["frame_idx"]], g["W_sub"]), rel_fro(Y, g["Y"])
    eo = abs(act.objective - float(g["objective"])) / float(g["objective"])
    e0 = abs(act.objective_at_init - float(g["objective_at_init"])) / float(g["objective_at_init"])
    print(f"config1-full[{mode}]: n_iter={act.n_iter} relF(H)={eh:.2e} relF(Y)={ey:.2e} rel(obj)={eo:.2e}")
    assert eh < 1e-3 and ey < 1e-3 and eo < 1e-4 and e0 < 1e-4, (eh, ey, eo, e0)


@pytest.mark.parametrize("mode", ["3xtf32", "tf32", "bf16"])
def test_run_to_run_bit_identity_full_shape(headline, mode):
    """20 solves x 50 iterations at the full shape must give the SAME BITS every time: every reduction in the path has
    a fixed order (split-K partials summed in split order, per-warp leftover partials in row order), so any difference
    is a race (a ring slot reused early, an accumulator drained late, a barrier phase slipped)."""
    import torch
    from exemplars_vc_b200 import ExemplarDictionary
    h = headline
    with ExemplarDictionary(h["A"], h["B"], mode=mode) as d:
        x = torch.from_numpy(h["X"]).cuda()
        first = d.solve(x, tol=0.0, max_iter=50)
        H0, Y0, obj0 = first.H.clone(), d.convert(first.H).clone(), first.objective
        for rep in range(19):
            act = d.solve(x, tol=0.0, max_iter=50)
            assert torch.equal(act.H, H0), f"run {rep + 1}: H differs in {(act.H != H0).sum().item()} entries"
            assert torch.equal(d.convert(act.H), Y0)
            assert act.objective == obj0


def test_conversion_decompose_frame_runs_and_matches():
    """05_conversion.py:94-107 executed: one 513-bin frame against the stacked dictionary (real speech)."""
    from exemplars_vc_b200 import conversion
    from oracle import nmf_oracle as o
    g = load_golden("speech_sf1_tf1_100162")
    W_dict, frame = g["A"], g["X"][5]
    files = [{"sp": W_dict[:300]}, {"sp": W_dict[300:]}, {"sp": W_dict[:7]}]
    stacked = conversion.stack_dictionary(files, drop_last=1)          # all but the last file, :94-98
    assert np.array_equal(stacked, W_dict)
    h = conversion.decompose_frame(frame, stacked, max_iter=200, tol=1e-4)
    W_ref, n_ref, _ = o.kl_mu(frame[None].astype(np.float64), W_dict.astype(np.float64), tol=1e-4, max_iter=200)
    assert h.shape == (W_dict.shape[0],) and h.dtype == W_dict.dtype
    assert rel_fro(h, W_ref[0]) < 1e-3


def test_pytorch_variant_factorize_and_convert_execute():
    """04_align_n_nmf_pytorch.py:213-327 executed on WORLD-style features (sp, ap, f0): two-argument convert, no
    residual, 200 iterations; the reference's solver='cd' is replaced by the multiplicative update (warned once)."""
    from exemplars_vc_b200 import align_n_nmf_pytorch as m
    from oracle import nmf_oracle as o
    rng = np.random.default_rng(21)
    N, F, T = 300, 129, 14
    As, Aa = rng.random((N, F)) ** 2 + 1e-3, rng.random((N, F)) ** 2 + 1e-3
    Bs, Ba = rng.random((N, F)) ** 2 + 1e-3, rng.random((N, F)) ** 2 + 1e-3
    f0d = np.where(rng.random(N) < 0.3, 0.0, 100 + 100 * rng.random(N))
    f0t = np.where(f0d > 0, f0d * 1.2, 0.0)
    Hs = rng.random((T, N)) * (rng.random((T, N)) < 0.03)
    Xs, Xa = Hs @ As + 0.01 * rng.random((T, F)), Hs @ Aa + 0.01 * rng.random((T, F))
    f0x = np.where(rng.random(T) < 0.3, 0.0, 150 + 50 * rng.random(T))
    src = [{"sp": As[:100], "ap": Aa[:100], "f0": f0d[:100]}, {"sp": As[100:], "ap": Aa[100:], "f0": f0d[100:]}]
    tar = [{"sp": Bs, "ap": Ba, "f0": f0t}]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H = m.factorize({"sp": Xs, "ap": Xa, "f0": f0x}, src)
        out = m.convert(H, tar)
    assert set(H) == {"H_sp", "H_ap", "H_f0"} and H["H_sp"].shape == (N, T)
    W_sp, _, _ = o.kl_mu(Xs, As, tol=1e-4, max_iter=200)
    W_f0, _, _ = o.kl_mu(f0x[:, None], f0d[:, None], tol=1e-4, max_iter=200)
    assert rel_fro(H["H_sp"].T, W_sp) < 1e-3 and rel_fro(H["H_f0"].T, W_f0) < 1e-3
    assert rel_fro(out["sp"], W_sp @ Bs) < 1e-3 and out["f0"].shape == (T,)
    assert rel_fro(out["f0"], (W_f0 @ f0t[:, None])[:, 0]) < 1e-3
    # the |stft| branch of this variant (no abs on the utterance side, :270)
    m.use_stft = 1
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Hst = m.factorize({"real": Xs}, [{"stft": As}])
        assert rel_fro(Hst["H_stft"].T, W_sp) < 1e-3
    finally:
        m.use_stft = 0


def test_script_literal_frobenius_through_factorize():
    """What the reference script REALLY runs: the body overwrites beta_loss with 'frobenius' (04_align_n_nmf.py:210).
    beta_override='frobenius' reproduces it; golden from the reference's own call."""
    from exemplars_vc_b200 import align_n_nmf as m
    from conftest import golden_inputs
    g = load_golden("fro_201x777x37_tol1e-4")
    X, A, _ = golden_inputs(g)
    old = (m.beta_override, m.max_iter)
    m.beta_override, m.max_iter = "frobenius", int(g["max_iter"])
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            H = m._factorize(X, A, tol=float(g["tol"]))
    finally:
        m.beta_override, m.max_iter = old
    assert H.shape == (777, 37) and rel_fro(H.T, g["W"]) < 1e-3


def test_nmf_tool_in_the_tensor_core_mode():
    """nmf_tool.nmf.NMF(initW=True) in the fp32-accurate tensor-core mode at a shape that fills MMA tiles."""
    from exemplars_vc_b200.nmf_tool.nmf import NMF
    from oracle import nmf_oracle as o
    rng = np.random.default_rng(9)
    m_, r, n = 257, 640, 70
    W = (rng.random((m_, r)) ** 2 + 1e-3).astype(np.float32)
    V = (W @ (rng.random((r, n)) * (rng.random((r, n)) < 0.05)) + 0.01).astype(np.float32)
    H0 = rng.random((r, n)).astype(np.float32)
    model = NMF(max_iter=40, display_step=0, optimizer="mu", mode="3xtf32")
    W_out, H = model.fit_transform(V, r_components=r, initW=True, givenW=W, H0=H0)
    H_ref, _ = o.nmf_tool_euclidean_mu(V.astype(np.float64), W.astype(np.float64), H0.astype(np.float64), 40)
    assert rel_fro(H, H_ref) < 1e-3
    assert rel_fro(model.inverse_transform(W_out, H), W.astype(np.float64) @ H_ref) < 1e-3


def test_content_keyed_cache_is_used_and_never_stale(tmp_path):
    """the reference's H_test_<feat>_<nfiles>.pkl cache is keyed by feature type and file count only
    (04_align_n_nmf.py:251-255); here the key is the content, so a second identical call hits and a different
    utterance does not."""
    from exemplars_vc_b200 import align_n_nmf as m
    rng = np.random.default_rng(33)
    A = rng.random((200, 65)) ** 2 + 1e-3
    X1 = (rng.random((9, 200)) * (rng.random((9, 200)) < 0.05)) @ A + 0.01
    X2 = (rng.random((9, 200)) * (rng.random((9, 200)) < 0.05)) @ A + 0.01
    old = (m.cache_dir, m.use_stft)
    m.cache_dir, m.use_stft = str(tmp_path), 1
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            H1, R1 = m.factorize({"real": X1}, [{"real": A}])
            files = sorted(os.listdir(tmp_path))
            assert len(files) == 1 and R1 is None
            with open(os.path.join(tmp_path, files[0]), "rb") as f:
                assert np.array_equal(pickle.load(f)["H_stft"], H1["H_stft"])
            # poison the stored result: a hit must return it (the cache is really read) ...
            with open(os.path.join(tmp_path, files[0]), "wb") as f:
                pickle.dump({"H_stft": np.full_like(H1["H_stft"], 7.0)}, f)
            H1b, _ = m.factorize({"real": X1}, [{"real": A}])
            assert np.all(H1b["H_stft"] == 7.0)
            # ... and another utterance of the same shape must NOT (the reference's key would have hit)
            H2, _ = m.factorize({"real": X2}, [{"real": A}])
            assert len(os.listdir(tmp_path)) == 2 and not np.all(H2["H_stft"] == 7.0)
    finally:
        m.cache_dir, m.use_stft = old


def test_in_kernel_split_k_sum_equals_the_separate_reduction_pass():
    """The split-K sum + ratio done by contraction 1's own CTAs (the first GEMM's epilogue, sklearn _nmf.py:554-571)
    gives bit for bit the results of the separate reduce_partials_kernel launch (EVC_NO_FUSED_REDUCE=1), at six shapes
    (split-K with and without leftover rows, several row groups; at T = 19000 K is not split and the ratio leaves
    straight from TMEM, the default there) in all three tensor-core modes.  EVC_FUSED_REDUCE=1 turns the in-kernel sum
    on for the split-K shapes too (it is not the default: measured slower than the separate pass).  The switch is read once per process, so each variant runs in its own interpreter."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "manual", "fused_reduce_ab.py")
    outs = []
    for extra in ({"EVC_FUSED_REDUCE": "1"}, {"EVC_NO_FUSED_REDUCE": "1"}):
        env = {k: v for k, v in os.environ.items() if k not in ("EVC_FUSED_REDUCE", "EVC_NO_FUSED_REDUCE")}
        env.update(extra)
        r = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if " H " in l])
    assert len(outs[0]) == 18 and outs[0] == outs[1], "\n".join(f"{a}\n{b}" for a, b in zip(*outs) if a != b)


ODD_SHAPES = [(64, 128, 1), (65, 129, 33), (136, 1000, 257), (200, 333, 100), (513, 700, 300), (513, 4100, 513),
              (641, 2049, 70), (1025, 520, 260), (257, 200, 19200), (2565, 1300, 40), (520, 130, 5)]


@pytest.mark.parametrize("mode", ["3xtf32", "tf32", "bf16"])
def test_odd_shapes_against_the_exact_fp32_kernels(mode):
    """Ragged shapes through the tensor-core kernels against the exact-fp32 CUDA-core kernels of the same library
    (an independent implementation of the same update, sklearn _nmf.py:521-626): leftover dictionary rows 0, 1, 5 and
    8 (F = 200 / 513, 641, 1025, 257 / 2565 / 136, 520), several dictionary-row groups, exemplar and frame counts
    that fill no tile, one frame, K split and not split (T = 19200: the ratio leaves straight from TMEM), narrow tail
    items.  Five iterations from the sklearn initialisation; H and Y within the mode's tolerance."""
    from exemplars_vc_b200 import ExemplarDictionary, synth
    tol = {"3xtf32": 1e-3, "tf32": 2e-2, "bf16": 3e-2}[mode]
    for F, N, T in ODD_SHAPES:
        A, B = synth.dictionaries(1000 + F + N, F, N)
        X = synth.frames(2000 + T, A, T)
        res = {}
        for m in ("fp32", mode):
            with ExemplarDictionary(A, B, mode=m) as d:
                act = d.solve(X, tol=0.0, max_iter=5)
                res[m] = (d.to_host(act.H), d.to_host(d.convert(act.H)), act.objective)
        eh, ey = rel_fro(res[mode][0], res["fp32"][0]), rel_fro(res[mode][1], res["fp32"][1])
        eo = abs(res[mode][2] - res["fp32"][2]) / max(res["fp32"][2], 1e-30)
        assert eh < tol and ey < tol and eo < tol, (mode, (F, N, T), eh, ey, eo)
