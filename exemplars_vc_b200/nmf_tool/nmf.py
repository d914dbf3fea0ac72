"""Drop-in for ``nmf_tool/nmf.py`` (TensorFlow 1.x graph NMF) on the fixed-dictionary path.

    NMF(max_iter=200, learning_rate=0.01, display_step=10, optimizer='mu', initW=False)
    .fit_transform(X, r_components, initW, givenW) -> (W, H)         nmf_tool/nmf.py:75-80
    .inverse_transform(W, H) -> W @ H                                 nmf_tool/nmf.py:82-84

North-star orientation, as in the reference file: X (m, n) = (features, frames), W (m, r) the dictionary,
H (r, n) the activations.  With ``initW=True`` the dictionary is the constant ``givenW``
(nmf_tool/nmf.py:29-31) and each iteration is the Euclidean multiplicative update
``H <- H * (W^T X) / ((W^T W) H)`` (nmf_tool/nmf.py:38-40), run here as the Frobenius mode of libevc_b200
with ``A^T (A H)`` instead of the r x r Gram.  H0 ~ U(0,1) comes from numpy (``self.seed``), because
TensorFlow's initializer stream is not reproducible outside TensorFlow; pass ``H0`` to pin it.

Not on the path and therefore not implemented (raises): ``initW=False`` (learning W) and the
projected-gradient optimizer ``'pg'`` (which in the reference never assigns its clamp, :50-51).
"""
from __future__ import annotations

import numpy as np

from ..dictionary import ExemplarDictionary


class NMF:
    """Compute Non-negative Matrix Factorization (NMF) activations over a given dictionary."""

    def __init__(self, max_iter=200, learning_rate=0.01, display_step=10, optimizer="mu", initW=False,
                 mode="3xtf32", seed=10):
        self.max_iter = max_iter
        self.learning_rate = learning_rate
        self.display_step = display_step
        self.optimizer = optimizer
        self.mode = mode
        self.seed = seed            # nmf_tool/nmf.py:7 np.random.seed(10)

    def NMF(self, X, r_components, learning_rate, max_iter, display_step, optimizer, initW, givenW, H0=None):
        if optimizer != "mu":
            raise NotImplementedError("optimizer=%r: only the multiplicative update 'mu' is on the path" % optimizer)
        if initW is False:
            raise NotImplementedError("initW=False (learning the dictionary W) is not on the exemplar-VC path")
        X = np.asarray(X, dtype=np.float32)
        m, n = X.shape
        W = np.asarray(givenW, dtype=np.float32).reshape(m, r_components)   # nmf_tool/nmf.py:30
        if H0 is None:
            H0 = np.random.default_rng(self.seed).random((r_components, n), dtype=np.float32)
        H0 = np.asarray(H0, dtype=np.float32)
        with ExemplarDictionary(W.T.copy(), None, mode=self.mode) as d:
            done = 0
            Ht = np.ascontiguousarray(H0.T)                # (n, r): frames are rows for the kernels
            cost = None
            while done < max_iter:
                # run in display_step chunks so the cost print-out of the reference (:69-71) keeps its cadence
                step = min(display_step if display_step > 0 else max_iter, max_iter - done)
                act = d.solve(X.T, beta_loss="frobenius", tol=0.0, max_iter=step, H0=Ht, epsilon=1e-30)
                Ht = act.H
                if display_step > 0:
                    cost = act.objective ** 2              # reduce_sum(square(V - WH)), :34
                    print("|Epoch:", "{:4d}".format(done), " Cost=", "{:.3f}".format(cost),
                          "learning rate: {}".format(learning_rate))
                done += step
            H = d.to_host(Ht).T.copy() if not isinstance(Ht, np.ndarray) else Ht.T.copy()
        return W, H

    def fit_transform(self, X, r_components, initW, givenW, H0=None):
        """Transform input data to W, H matrices which are the non-negative matrices."""
        W, H = self.NMF(X=X, r_components=r_components, learning_rate=self.learning_rate, max_iter=self.max_iter,
                        display_step=self.display_step, optimizer=self.optimizer, initW=initW, givenW=givenW, H0=H0)
        return W, H

    def inverse_transform(self, W, H):
        """Transform data back to its original space."""
        return np.matmul(W, H)
