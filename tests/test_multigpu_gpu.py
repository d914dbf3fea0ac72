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
    d = sharding.make_exemplar_sharded(lambda a, b: A[a:b], lambda a, b: B[a:b], 1000, mode=mode)
    act = d.solve(X, tol=1e-4, max_iter=40)
    H = d.to_host(act.H).astype(np.float64)
    Y = d.to_host(d.convert(act.H)).astype(np.float64)
    assert act.n_iter == n_ref
    assert np.linalg.norm(H - W_ref[:, d.n_begin:d.n_end]) / np.linalg.norm(W_ref[:, d.n_begin:d.n_end]) < 1e-3
    assert np.linalg.norm(Y - W_ref @ B) / np.linalg.norm(W_ref @ B) < 1e-3      # Y is all-reduced: complete on every rank
    assert abs(act.objective - obj_ref) / obj_ref < 1e-4
    d.close()

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
