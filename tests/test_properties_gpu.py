"""Size-independent properties at BASELINE.json's full single-GPU size (F=513, N=20000, T=1000), where the
oracle would take minutes: they hold for the exact algorithm, so they check tiling, split-K, padding and
the fused epilogues without a reference value.  A short oracle cross-check at the full dictionary size
(few frames, few iterations) anchors them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

F, N, T = 513, 20000, 1000


@pytest.fixture(scope="module")
def problem():
    from exemplars_vc_b200 import synth
    A, B = synth.dictionaries(synth.BASE_SEED + 1, F, N)
    X = synth.frames(synth.BASE_SEED + 1, A, T)
    return X, A, B


@pytest.fixture(scope="module", params=["3xtf32", "tf32", "fp32"])
def dictionary(request, problem):
    from exemplars_vc_b200 import ExemplarDictionary
    X, A, B = problem
    d = ExemplarDictionary(A, B, mode=request.param)
    yield d
    d.close()


def test_objective_is_monotone_and_h_nonnegative(dictionary, problem):
    X, A, B = problem
    if dictionary.mode == "fp32":
        X = X[:128]
    objs, H = [], None
    for k in (1, 2, 3):
        act = dictionary.solve(X, tol=0.0, max_iter=10 * k)
        objs.append(act.objective)
        H = act.H
    assert objs[0] > objs[1] > objs[2] > 0
    assert act.objective_at_init > objs[0]
    assert float(H.min()) >= 0.0 and bool(torch.isfinite(H).all())


def test_frames_are_independent(dictionary, problem):
    """Column t of H depends only on column t of X: solving a slice of the frames from the same H0 gives the
    same rows (the tile / split-K decomposition changes, the answer must not)."""
    X, A, B = problem
    n = 128 if dictionary.mode == "fp32" else T
    H0 = np.full((n, N), 0.01, dtype=np.float32)
    full = dictionary.solve(X[:n], tol=0.0, max_iter=5, H0=H0).H
    lo, hi = 37, 101
    part = dictionary.solve(X[lo:hi], tol=0.0, max_iter=5, H0=H0[lo:hi]).H
    rel = float(torch.linalg.norm(part - full[lo:hi]) / torch.linalg.norm(part))
    # different T -> different split-K plan -> different accumulation chains in TMEM (accumulation)
    assert rel < {"fp32": 1e-5, "3xtf32": 1e-4, "tf32": 1e-2}[dictionary.mode], rel


def test_scale_equivariance(dictionary, problem):
    """X -> cX, H0 -> cH0 gives H -> cH exactly for c a power of two (every operation is homogeneous)."""
    X, A, B = problem
    n = 64
    H0 = np.full((n, N), 0.01, dtype=np.float32)
    h1 = dictionary.solve(X[:n], tol=0.0, max_iter=3, H0=H0).H
    h4 = dictionary.solve(4.0 * X[:n], tol=0.0, max_iter=3, H0=4.0 * H0).H
    assert torch.equal(4.0 * h1, h4)


def test_exact_model_is_a_fixed_point(dictionary, problem):
    """If X = H* A exactly, the ratio is 1 and the update factor is A^T1 / A^T1 = 1."""
    X, A, B = problem
    rng = np.random.default_rng(2)
    n = 64
    Hs = (rng.random((n, N)) * (rng.random((n, N)) < 0.001)).astype(np.float32)
    Xs = dictionary.to_host(dictionary.reconstruct(Hs))
    h = dictionary.solve(Xs, tol=0.0, max_iter=2, H0=Hs).H
    rel = float(torch.linalg.norm(h.cpu() - torch.from_numpy(Hs)) / np.linalg.norm(Hs))
    assert rel < (1e-5 if dictionary.mode != "tf32" else 5e-3), rel
    assert dictionary.objective(Xs, Hs) < (1e-2 if dictionary.mode != "tf32" else 1.0)


def test_conversion_is_linear(dictionary, problem):
    X, A, B = problem
    rng = np.random.default_rng(4)
    H1 = rng.random((96, N)).astype(np.float32)
    H2 = rng.random((96, N)).astype(np.float32)
    y1, y2 = dictionary.convert(H1), dictionary.convert(H2)
    y = dictionary.convert(2.0 * H1 + 0.5 * H2)
    rel = float(torch.linalg.norm(y - (2.0 * y1 + 0.5 * y2)) / torch.linalg.norm(y))
    assert rel < (1e-5 if dictionary.mode != "tf32" else 2e-3), rel


def test_products_against_float64(dictionary, problem):
    """Both contraction shapes at full dictionary size against numpy float64 (8 frames keep it cheap on CPU)."""
    X, A, B = problem
    rng = np.random.default_rng(6)
    H = rng.random((8, N)).astype(np.float32)
    y = dictionary.to_host(dictionary.convert(H)).astype(np.float64)
    ref = H.astype(np.float64) @ B.astype(np.float64)
    rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    assert rel < {"fp32": 1e-5, "3xtf32": 1e-4, "tf32": 2e-3}[dictionary.mode], rel
    wh = dictionary.to_host(dictionary.reconstruct(H)).astype(np.float64)
    ref = H.astype(np.float64) @ A.astype(np.float64)
    rel = np.linalg.norm(wh - ref) / np.linalg.norm(ref)
    assert rel < {"fp32": 1e-5, "3xtf32": 1e-4, "tf32": 2e-3}[dictionary.mode], rel


def test_oracle_crosscheck_full_dictionary(dictionary, problem):
    """16 frames x 20 iterations against the oracle at N = 20000 (seconds on CPU)."""
    from oracle import nmf_oracle as o
    X, A, B = problem
    Xs = X[:16]
    W_ref, n, obj = o.kl_mu(Xs.astype(np.float64), A.astype(np.float64), tol=0.0, max_iter=20)
    act = dictionary.solve(Xs, tol=0.0, max_iter=20)
    H = dictionary.to_host(act.H).astype(np.float64)
    rel = np.linalg.norm(H - W_ref) / np.linalg.norm(W_ref)
    tol_h, tol_o = (1e-3, 1e-4) if dictionary.mode != "tf32" else (2e-2, 2e-2)
    assert rel < tol_h, rel
    assert abs(act.objective - obj) / obj < tol_o, (act.objective, obj)
