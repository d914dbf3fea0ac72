"""Generate tests/golden/*.npz from the reference's own operator call.  TEST INFRASTRUCTURE.

Run in the authoring container (``python oracle/make_golden.py``): it executes the
reference's exact call (04_align_n_nmf.py:212-213 -> scikit-learn 1.9.0
``non_negative_factorization``) on seeded inputs and stores the outputs.  The GPU box has
no /root/reference and must not need sklearn for parity, so the vectors are committed.

Every file stores the generator arguments (inputs are regenerated from the seed by
``oracle.nmf_oracle.gen``; edge-case inputs that are not seed-derivable are stored in full)
plus the reference outputs W (= H^T, (T,N)), n_iter, objective sqrt(2 KL), and Y = W B.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nmf_oracle as o  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 20190123


def versions():
    import sklearn
    return np.array([f"sklearn={sklearn.__version__}", f"numpy={np.__version__}"])


def sk_call(X, A, beta="kullback-leibler", tol=1e-4, max_iter=150, alpha_W=0.0, l1_ratio=0.0):
    from sklearn.decomposition import non_negative_factorization
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        W, _H, n_iter = non_negative_factorization(
            X=X, H=A, init="custom", update_H=False, n_components=A.shape[0], beta_loss=beta,
            solver="mu", tol=tol, max_iter=max_iter, alpha_W=alpha_W, l1_ratio=l1_ratio)
    return W, n_iter


def synth_case(name, F, N, T, tol, max_iter, dtype=np.float64, beta="kullback-leibler", keep_W=True):
    X, A, B = o.gen(SEED, F, N, T, dtype)
    W, n_iter = sk_call(X, A, beta, tol, max_iter)
    obj = o.kl_objective(X, W, A) if beta != "frobenius" else o.frobenius_objective(X, W, A)
    Y = o.convert(W, B)
    d = dict(seed=SEED, F=F, N=N, T=T, tol=tol, max_iter=max_iter, beta=beta, dtype=np.dtype(dtype).name,
             n_iter=n_iter, objective=obj, sum_W=W.sum(), W00=W[0, 0], norm_W=np.linalg.norm(W),
             norm_Y=np.linalg.norm(Y), sum_X=X.sum(), sum_A=A.sum(), versions=versions(),
             w0=np.sqrt(X.mean() / N), objective_at_init=o.kl_objective(X, o.initial_activation(X, N), A))
    if keep_W:
        d["W"] = W
        d["Y"] = Y
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: n_iter={n_iter} obj={obj:.10e} sumW={W.sum():.10e}")


def edge_case(name, X, A, B, tol, max_iter, **kw):
    W, n_iter = sk_call(X, A, "kullback-leibler", tol, max_iter, **kw)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, A=A, B=B, tol=tol, max_iter=max_iter,
                        W=W, n_iter=n_iter, objective=o.kl_objective(X, W, A), Y=o.convert(W, B),
                        versions=versions(), **{k: np.float64(v) for k, v in kw.items()})
    print(f"{name}: n_iter={n_iter} obj={o.kl_objective(X, W, A):.10e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    # Known-answer table of SURVEY.md 8(c): toy size, every stop-rule outcome.
    synth_case("kl_13x32x8_tol1e-2", 13, 32, 8, 1e-2, 150)
    synth_case("kl_13x32x8_tol1e-3", 13, 32, 8, 1e-3, 150)
    synth_case("kl_13x32x8_tol1e-4", 13, 32, 8, 1e-4, 150)
    synth_case("kl_13x32x8_tol0_500", 13, 32, 8, 0.0, 500)
    synth_case("kl_13x32x8_f32", 13, 32, 8, 1e-4, 150, dtype=np.float32)
    # 513-bin frames (the reference's WORLD feature width), small dictionary.
    synth_case("kl_513x2000x64_tol1e-4", 513, 2000, 64, 1e-4, 150)
    synth_case("kl_513x2000x64_tol0_500", 513, 2000, 64, 0.0, 500, keep_W=False)
    # ragged sizes: nothing a multiple of a tile
    synth_case("kl_201x777x37_tol1e-4", 201, 777, 37, 1e-4, 150)
    # Frobenius: what 04_align_n_nmf.py:210 actually runs
    synth_case("fro_13x32x8_tol1e-4", 13, 32, 8, 1e-4, 150, beta="frobenius")
    synth_case("fro_201x777x37_tol1e-4", 201, 777, 37, 1e-4, 150, beta="frobenius")

    # Edge cases the domain has (SURVEY 8c rules 3,4; hard part "degenerate F=1").
    rng = np.random.default_rng(SEED + 100)
    X, A, B = o.gen(SEED + 7, 17, 40, 9)
    X[rng.random(X.shape) < 0.25] = 0.0          # exact zeros in X: masked out of the objective
    A[3, :] = 0.0                                 # an all-zero exemplar: den == 0 -> EPSILON
    A[11, :] = 0.0
    edge_case("kl_edge_zeros", X, A, B, 1e-4, 150)
    # f0 track: F = 1, unvoiced zeros in both the frames and the dictionary (04_align_n_nmf.py:288)
    f0_dict = np.where(rng.random(60) < 0.3, 0.0, 100.0 + 150.0 * rng.random(60))[:, None]
    f0_tgt = np.where(rng.random(60) < 0.3, 0.0, 180.0 + 120.0 * rng.random(60))[:, None]
    f0_x = np.where(rng.random(25) < 0.3, 0.0, 100.0 + 150.0 * rng.random(25))[:, None]
    edge_case("kl_edge_f0", f0_x, f0_dict, f0_tgt, 1e-4, 150)
    # lambda > 0 through sklearn: reproduces quirk Q1 (den = A^T 1 + k*lambda at iteration k)
    X, A, B = o.gen(SEED + 9, 13, 32, 8)
    edge_case("kl_l1_sklearn_q1", X, A, B, 0.0, 50, alpha_W=0.01, l1_ratio=1.0)
    # constant-lambda restatement (the north star's formula); no sklearn equivalent
    lam = 13 * 0.01
    W, n_iter, obj = o.kl_mu(X, A, lam=lam, tol=0.0, max_iter=50)
    np.savez_compressed(os.path.join(OUT, "kl_l1_constant.npz"), X=X, A=A, B=B, lam=lam, tol=0.0, max_iter=50,
                        W=W, n_iter=n_iter, objective=obj, Y=o.convert(W, B), versions=versions())
    print("kl_l1_constant:", n_iter, obj)


if __name__ == "__main__":
    main()
