"""Multi-GPU sharding of the activation path.  One process per GPU,
``torch.distributed`` for the plumbing.

* Utterance (T) sharding: frames are independent, so each rank converts its own utterances against a
  full replica of the dictionary.  No collective on the data path; results are gathered at the end only
  if asked.
* Exemplar (N) sharding: rank g holds rows [n_begin, n_end) of A and B and the matching columns of H.
  Each iteration the partial A_g H_g (T,F) is summed across ranks (ncclAllReduce issued by libevc_b200
  on the solve stream); the ratio is formed redundantly, H columns never move.  Y = sum_g B_g H_g takes
  one more all-reduce.

The functions that decide WHO owns WHAT are pure (tested on CPU with gloo, world_size 2); the compute
object is injected so those tests can stand a checker in for the GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def exemplar_range(N: int, rank: int, world: int, align: int = 128):
    """Contiguous, balanced [n_begin, n_end) for `rank`; boundaries aligned to the 128-row MMA tile."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    tiles = (N + align - 1) // align
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, N), min(t1 * align, N)


def partition_utterances(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of utterances to ranks, balanced by frame count.
    Deterministic (ties broken by index) so every rank computes the same plan without talking."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(lengths[i])
    for p in plan:
        p.sort()
    return plan


def _dist():
    import torch.distributed as dist
    return dist


def convert_utterances(dictionary, utterances: Sequence[np.ndarray], *, group=None, gather: bool = False,
                       **solve_kw):
    """Utterance-sharded batch conversion.  `dictionary` is a full replica on this rank's GPU (an
    ExemplarDictionary, or any object with solve_batched / convert / to_host).  Each rank stacks ITS
    utterances along T, solves them in one batched call (per-utterance H0 and stop rule, like separate
    reference calls) and converts.  Returns {utterance index: (Y, n_iter)} for the local utterances, or
    for all of them on every rank when gather=True."""
    dist = _dist()
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lengths = [int(u.shape[0]) for u in utterances]
    mine = partition_utterances(lengths, world)[rank]
    out = {}
    if mine:
        stacked = np.concatenate([utterances[i] for i in mine], axis=0)
        offs = np.concatenate([[0], np.cumsum([lengths[i] for i in mine])]).astype(int).tolist()
        acts = dictionary.solve_batched(stacked, offs, per_utterance_stop=True, **solve_kw)
        for k, i in enumerate(mine):
            y = dictionary.to_host(dictionary.convert(acts[k].H))
            out[i] = (y, acts[k].n_iter)
    if gather and world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, out, group=group)
        out = {}
        for p in parts:
            out.update(p)
    return out


def broadcast_unique_id(make_id: Callable[[], bytes], *, group=None, src: int = 0) -> bytes:
    """Rank `src` creates the communicator id, everyone receives it (any torch.distributed backend)."""
    dist = _dist()
    box = [make_id() if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def make_exemplar_sharded(A_rows: Callable[[int, int], np.ndarray], B_rows: Optional[Callable[[int, int], np.ndarray]],
                          N: int, *, mode: str = "3xtf32", group=None, p2p: bool = True, max_frames: int = 4096):
    """Build this rank's shard of an N-exemplar dictionary.  `A_rows(n0, n1)` / `B_rows(n0, n1)` return the
    rows the rank owns (so a 200k-exemplar dictionary is never materialised whole on one host).  The
    returned ExemplarDictionary all-reduces partial A*H inside solve / convert / objective."""
    import ctypes as C

    from . import _lib
    from .dictionary import ExemplarDictionary

    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # every rank can compute every rank's range: an empty shard makes ALL ranks raise, before any collective is
    # entered (one rank raising alone would leave the others hanging in the broadcast below)
    ranges = [exemplar_range(N, r, world) for r in range(world)]
    empty = [r for r, (a, b) in enumerate(ranges) if b <= a]
    if empty:
        raise ValueError(f"ranks {empty} would own no exemplars (N={N}, world={world}): use fewer ranks")
    n0, n1 = ranges[rank]
    d = ExemplarDictionary(A_rows(n0, n1), B_rows(n0, n1) if B_rows is not None else None, mode=mode)

    def make_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().evc_comm_unique_id(buf))
        return buf.raw

    uid = broadcast_unique_id(make_id, group=group)
    d.attach_comm(uid, rank, world, N)
    d.n_begin, d.n_end = n0, n1
    d.all_reduce = "nccl"
    if p2p and 2 <= world <= 8:
        # our own all-reduce kernel over NVLink peer memory (CUDA IPC between the per-GPU processes of one node);
        # every rank must succeed, otherwise all of them stay on NCCL
        try:
            handle, err = d.p2p_alloc(max_frames), None
        except Exception as e:  # pragma: no cover - depends on the container's IPC permissions
            handle, err = None, repr(e)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        if all(h is not None for h in handles):
            try:
                d.p2p_attach(handles, rank, world)
                ok = True
            except Exception as e:  # pragma: no cover
                ok, err = False, repr(e)
        else:
            ok = False
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=group)
        if all(oks):
            d.all_reduce = "p2p"
        else:
            # not every rank could map its peers: ALL ranks use NCCL (a rank that did attach lets go again)
            if ok:
                d.p2p_detach()
            if rank == 0:
                import warnings
                warnings.warn("peer-memory all-reduce unavailable on ranks %s, using NCCL on all ranks%s"
                              % ([r for r, o_ in enumerate(oks) if not o_], (": " + err) if err else ""))
    return d
