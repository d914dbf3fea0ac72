"""World-size-2 CPU tests (gloo) of the multi-GPU host logic.  The GPU is replaced by a checker object
backed by the oracle -- test infrastructure only; the product has no CPU path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Act:
    def __init__(self, H, n_iter):
        self.H, self.n_iter = H, n_iter


class OracleDictionary:
    """Stands in for ExemplarDictionary in the gloo tests: same solve_batched / convert / to_host surface."""

    def __init__(self, A, B):
        self.A, self.B = A, B

    def solve_batched(self, X, offs, per_utterance_stop=True, tol=1e-4, max_iter=150, **kw):
        from oracle import nmf_oracle as o
        out = []
        for a, b in zip(offs, offs[1:]):
            W, n, _ = o.kl_mu(X[a:b], self.A, tol=tol, max_iter=max_iter)
            out.append(_Act(W, n))
        return out

    def convert(self, H):
        return H @ self.B

    def to_host(self, y):
        return np.asarray(y)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_utterances(rank, world, port, ret):
    _init(rank, world, port)
    from exemplars_vc_b200 import sharding
    from oracle import nmf_oracle as o
    _, A, B = o.gen(11, 9, 24, 1)
    rng = np.random.default_rng(5)
    utts = [o.gen(100 + i, 9, 24, int(rng.integers(3, 9)))[0] for i in range(7)]
    # inputs must use the shared dictionary: regenerate X_i against A
    utts = [np.abs(u) for u in utts]
    d = OracleDictionary(A, B)
    local = sharding.convert_utterances(d, utts, gather=False, tol=1e-4, max_iter=30)
    plan = sharding.partition_utterances([u.shape[0] for u in utts], world)
    assert sorted(local) == plan[rank]
    full = sharding.convert_utterances(d, utts, gather=True, tol=1e-4, max_iter=30)
    assert sorted(full) == list(range(7))
    for i, u in enumerate(utts):
        W, n, _ = o.kl_mu(u, A, tol=1e-4, max_iter=30)
        np.testing.assert_allclose(full[i][0], W @ B, rtol=1e-12)
        assert full[i][1] == n
    dist.destroy_process_group()


def _worker_exemplar_math(rank, world, port, ret):
    """The decomposition libevc_b200 uses for exemplar sharding: partial A_g H_g summed across ranks,
    ratio formed redundantly, H columns local.  Here with numpy + gloo; must equal the unsharded oracle."""
    _init(rank, world, port)
    from exemplars_vc_b200 import sharding
    from oracle import nmf_oracle as o
    X, A, B = o.gen(21, 17, 300, 6)
    n0, n1 = sharding.exemplar_range(300, rank, world)
    assert (n0, n1) == ((0, 256) if rank == 0 else (256, 300))
    Ag, Bg = A[n0:n1], B[n0:n1]
    W = np.full((6, n1 - n0), np.sqrt(X.mean() / 300))           # H0 uses n_total, not the shard size
    colsum = Ag.sum(1)
    for _ in range(25):
        part = torch.from_numpy(W @ Ag)
        dist.all_reduce(part)                                      # the per-iteration F x T exchange
        WH = np.maximum(part.numpy(), o.EPSILON)
        W *= ((X / WH) @ Ag.T) / colsum
    y = torch.from_numpy(W @ Bg)
    dist.all_reduce(y)
    W_ref, _, _ = o.kl_mu(X, A, tol=0.0, max_iter=25)
    np.testing.assert_allclose(W, W_ref[:, n0:n1], rtol=1e-10)
    np.testing.assert_allclose(y.numpy(), W_ref @ B, rtol=1e-10)
    uid = sharding.broadcast_unique_id(lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    dist.destroy_process_group()


def _run(fn):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(fn, args=(2, port, None), nprocs=2, join=True)


@pytest.mark.timeout(180)
def test_utterance_sharding_world2():
    _run(_worker_utterances)


@pytest.mark.timeout(180)
def test_exemplar_sharding_math_world2():
    _run(_worker_exemplar_math)
