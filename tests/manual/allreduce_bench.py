"""Per-call cost of the partial-A*H exchange: libevc_b200's peer-memory kernel vs ncclAllReduce.
torchrun --nproc-per-node N tests/manual/allreduce_bench.py   (tiny dictionary, T = 2000 frames, F = 513: 4.1 MB)"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplars_vc_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, T, n_per = 513, 2000, 256
N = n_per * world
rng = np.random.default_rng(0)
A = (rng.random((N, F)) + 0.1).astype(np.float32)
H = torch.rand((T, n_per), device="cuda")
for p2p in (True, False, True, False):
    d = sharding.make_exemplar_sharded(lambda a, b: A[a:b], lambda a, b: A[a:b], N, mode="3xtf32", p2p=p2p, max_frames=T)
    for _ in range(20):
        d.reconstruct(H)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        d.reconstruct(H)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 300
    d.profile(True)
    for _ in range(50):
        d.reconstruct(H)
    prof = d.profile_read()
    d.profile(False)
    if rank == 0:
        ex = prof["objective_init"]
        print(f"world {world} {d.all_reduce:5s}: {dt * 1e6:7.1f} us per reconstruct call (tiny GEMM + exchange + copy); "
              f"exchange kernel alone {ex[0] / max(ex[1], 1) * 1e3:6.1f} us", flush=True)
    d.close()
dist.destroy_process_group()
