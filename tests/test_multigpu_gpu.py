"""Multi-GPU paths on real GPUs (run with `gpurun --gpus 2`): skipped when fewer than 2 devices are visible.
Launches torchrun-style workers with NCCL: utterance sharding (no data-path collective) and exemplar
sharding (per-iteration all-reduce of partial A*H issued by libevc_b200 on the solve stream)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from exemplars_vc_b200 import ExemplarDictionary, sharding
    from oracle import nmf_oracle as o

    # ---- exemplar sharding: N = 1000 rows split at a 128-row boundary, 513 bins (leftover row path too)
    X, A, B = o.gen(71, 513, 1000, 40)
    W_ref, n_ref, obj_ref = o.kl_mu(X, A, tol=1e-4, max_iter=40)
    results = {}
    for p2p in (True, False):      # libevc_b200's own all-reduce over NVLink peer memory, then ncclAllReduce
        d = sharding.make_exemplar_sharded(lambda a, b: A[a:b], lambda a, b: B[a:b], 1000, mode=mode, p2p=p2p,
                                           max_frames=64)
        if p2p and d.all_reduce != "p2p" and rank == 0:
            print("peer-memory all-reduce not available here; NCCL used for both passes")
        act = d.solve(X, tol=1e-4, max_iter=40)
        H = d.to_host(act.H).astype(np.float64)
        Y = d.to_host(d.convert(act.H)).astype(np.float64)
        assert act.n_iter == n_ref
        assert np.linalg.norm(H - W_ref[:, d.n_begin:d.n_end]) / np.linalg.norm(W_ref[:, d.n_begin:d.n_end]) < 1e-3
        assert np.linalg.norm(Y - W_ref @ B) / np.linalg.norm(W_ref @ B) < 1e-3   # Y is all-reduced: complete on every rank
        assert abs(act.objective - obj_ref) / obj_ref < 1e-4
        results[d.all_reduce] = (H, Y, act.objective)
        with pytest.raises(ValueError):
            if d.all_reduce == "p2p":
                d.solve(np.concatenate([X, X]), tol=0.0, max_iter=1)      # more frames than the exchange buffer holds
            else:
                raise ValueError("n/a")
        d.close()
    if len(results) == 2:      # same partials, both sums in a fixed order: the two exchanges agree to rounding
        assert np.allclose(results["p2p"][1], results["nccl"][1], rtol=1e-5, atol=0)

    # ---- utterance sharding: every rank holds the whole dictionary, converts its own utterances
    rng = np.random.default_rng(5)
    lens = [int(v) for v in rng.integers(5, 30, size=6)]
    utts = [(rng.random((L, 1000)) * (rng.random((L, 1000)) < 0.01)) @ A + 0.01 * rng.random((L, 513)) for L in lens]
    with ExemplarDictionary(A, B, mode=mode) as full:
        out = sharding.convert_utterances(full, utts, gather=True, tol=1e-3, max_iter=40)
    assert sorted(out) == list(range(6))
    if rank == 0:
        for i, u in enumerate(utts):
            W, n, _ = o.kl_mu(u, A, tol=1e-3, max_iter=40)
            assert out[i][1] == n
            assert np.linalg.norm(out[i][0] - W @ B) / np.linalg.norm(W @ B) < 1e-3
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["3xtf32", "fp32"])
def test_sharded_paths_two_gpus(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, mode), nprocs=2, join=True)


def test_two_devices_in_one_process():
    """the > 48 KB dynamic shared-memory opt-in and the SM count are per-device state; a second GPU used
    from the same process must launch the tensor-core kernels too."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from exemplars_vc_b200 import ExemplarDictionary
    from oracle import nmf_oracle as o
    X, A, B = o.gen(17, 257, 640, 24)
    W_ref, n_ref, obj = o.kl_mu(X, A, tol=1e-4, max_iter=30)
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            with ExemplarDictionary(A, B, mode="3xtf32", device=dev) as d:
                act = d.solve(X, tol=1e-4, max_iter=30)
                H = d.to_host(act.H).astype(np.float64)
        assert act.n_iter == n_ref and np.linalg.norm(H - W_ref) / np.linalg.norm(W_ref) < 1e-3, dev
