"""Device-resident exemplar dictionary: the host-side object behind the reference's entry points.

The reference rebuilds ``A_sp = np.asarray(A_sp)`` and hands it to scikit-learn on every call
(04_align_n_nmf.py:230-246, 212-213).  Here the source dictionary A (N,F) and the paired target
dictionary B (N,F) are uploaded once, kept in HBM across utterances (re-pitched for TMA, split /
transposed as the arithmetic mode needs, A^T 1 cached) and every solve / convert runs on the GPU
through the C ABI in include/evc.h.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import SolveParams, SolveResult, check


@dataclass
class Activation:
    """Result of one activation solve.  ``H`` is (T, N) on the device: the reference's ``_W``."""
    H: torch.Tensor
    n_iter: int
    converged: bool
    objective: float
    objective_at_init: float
    H_stacked: Optional[torch.Tensor] = None   # solve_batched: the (T_total, N) tensor every utterance's H is a row view of


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _as_device_f32(x, device) -> torch.Tensor:
    """numpy / torch (any float dtype, host or device) -> contiguous fp32 CUDA tensor with an aligned pitch."""
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D array, got shape {tuple(t.shape)}")
    return t.to(device=device, dtype=torch.float32, non_blocking=True)


def _pitched(t: torch.Tensor, mult: int = 4) -> torch.Tensor:
    """Return a view (rows, cols) of a fresh buffer whose row pitch is a multiple of `mult` floats."""
    rows, cols = t.shape
    ld = _round_up(max(cols, 1), mult)
    if t.is_contiguous() and cols == ld and t.data_ptr() % 16 == 0:
        return t
    buf = torch.zeros((rows, ld), dtype=torch.float32, device=t.device)
    buf[:, :cols].copy_(t)
    return buf[:, :cols]


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ExemplarDictionary:
    """Source/target exemplar dictionaries resident on one GPU.

    Parameters
    ----------
    A : (N, F) array, non-negative -- source exemplars, rows are frames (``W`` of ``_factorize``).
    B : (N, F) array or None      -- paired target exemplars (``B_sp`` / ``B_stft`` of ``convert``).
    mode : "3xtf32" (fp32-accurate, default) | "tf32" | "bf16" (fast modes) | "fp32" (CUDA-core FFMA).
    """

    def __init__(self, A, B=None, mode: str = "3xtf32", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("exemplars_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        if mode not in _lib.MODES:
            raise ValueError(f"mode must be one of {sorted(_lib.MODES)}, got {mode!r}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mode = mode
        self._h = C.c_void_p(0)
        self._comm = C.c_void_p(0)
        self._pinned = {}
        L = _lib.lib()
        with torch.cuda.device(self.device):
            a = _as_device_f32(A, self.device)
            b = _as_device_f32(B, self.device) if B is not None else None
            if b is not None and tuple(b.shape) != tuple(a.shape):
                raise ValueError(f"A {tuple(a.shape)} and B {tuple(b.shape)} must have the same shape")
            self.N, self.F = int(a.shape[0]), int(a.shape[1])
            self.n_total = self.N
            a = a.contiguous()
            b = b.contiguous() if b is not None else None
            check(L.evc_dict_create(_ptr(a), self.F, _ptr(b), self.F, self.F, self.N, _lib.MODES[mode],
                                    _stream(self.device), C.byref(self._h)))
            torch.cuda.current_stream(self.device).synchronize()  # a, b may be freed now

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().evc_dict_destroy(self._h)
            self._h = C.c_void_p(0)
        if getattr(self, "_comm", None) is not None and self._comm.value:
            _lib.lib().evc_comm_destroy(self._comm)
            self._comm = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def has_target(self) -> bool:
        f, n, m, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(_lib.lib().evc_dict_info(self._h, C.byref(f), C.byref(n), C.byref(m), C.byref(t)))
        return bool(t.value)

    def colsum(self) -> torch.Tensor:
        """A^T 1 (sklearn's cached ``H_sum``, _nmf.py:588-590)."""
        out = torch.empty(self.N, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.lib().evc_dict_colsum(self._h, _ptr(out), _stream(self.device)))
        return out

    # -- exemplar sharding ---------------------------------------------------------------
    def attach_comm(self, unique_id: bytes, rank: int, world: int, n_total: int):
        """This handle holds rows [n_begin, n_end) of a dictionary of n_total exemplars; partial A*H is
        all-reduced over NCCL inside evc_solve / evc_convert."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            check(L.evc_comm_create(unique_id, rank, world, C.byref(self._comm)))
        check(L.evc_dict_attach_comm(self._h, self._comm, n_total))
        self.n_total = n_total

    def p2p_alloc(self, max_frames: int) -> bytes:
        """Allocate this rank's peer-memory exchange buffer (for up to max_frames frames); returns its CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(_lib.lib().evc_p2p_alloc(self._h, int(max_frames), buf))
        return buf.raw

    def p2p_attach(self, handles: Sequence[bytes], rank: int, world: int):
        """Map the peers' exchange buffers: from now on the partial A*H is summed by libevc_b200's own kernel over
        NVLink peer memory instead of ncclAllReduce."""
        blob = b"".join(handles)
        if len(blob) != 64 * world:
            raise ValueError("expected one 64-byte IPC handle per rank")
        with torch.cuda.device(self.device):
            check(_lib.lib().evc_p2p_attach(self._h, blob, rank, world))

    def p2p_detach(self):
        """Back to ncclAllReduce (every rank must use the same exchange; see sharding.make_exemplar_sharded)."""
        check(_lib.lib().evc_p2p_detach(self._h))

    # -- the hot path ---------------------------------------------------------------------------------
    def _params(self, beta_loss, tol, max_iter, lam, lambda_step, init_given, check_every, epsilon) -> SolveParams:
        p = SolveParams()
        _lib.lib().evc_default_params(C.byref(p))
        if beta_loss in ("kullback-leibler", 1, 1.0):
            p.loss = _lib.LOSS_KL
        elif beta_loss in ("frobenius", 2, 2.0):
            p.loss = _lib.LOSS_FROBENIUS
        else:
            raise NotImplementedError(f"beta_loss={beta_loss!r}: only 'kullback-leibler' and 'frobenius' are on the path")
        p.init = _lib.INIT_GIVEN if init_given else _lib.INIT_SKLEARN
        p.max_iter, p.check_every = int(max_iter), int(check_every)
        p.tol, p.lam, p.lambda_step, p.epsilon = float(tol), float(lam), float(lambda_step), float(epsilon)
        return p

    def _prep_frames(self, X) -> torch.Tensor:
        x = _as_device_f32(X, self.device)
        if x.shape[1] != self.F:
            raise ValueError(f"X has {x.shape[1]} features, the dictionary has {self.F}")
        return _pitched(x)

    def solve(self, X, beta_loss="kullback-leibler", tol=1e-4, max_iter=150, lam=0.0, lambda_step=0.0,
              H0=None, check_every=10, epsilon=0.0) -> Activation:
        """Activations of the frames X (T,F) over the source dictionary.  Replaces the scikit-learn call
        at 04_align_n_nmf.py:212-213.  Returns H as a (T,N) device tensor (the reference's ``_W``)."""
        res = self.solve_batched(X, None, beta_loss, tol, max_iter, lam, lambda_step, H0, check_every, epsilon,
                                 per_utterance_stop=False)
        return res[0]

    def solve_batched(self, X_stacked, t_offsets: Optional[Sequence[int]], beta_loss="kullback-leibler", tol=1e-4,
                      max_iter=150, lam=0.0, lambda_step=0.0, H0=None, check_every=10, epsilon=0.0,
                      per_utterance_stop=True):
        """Stacked-T mode: utterance u owns rows [t_offsets[u], t_offsets[u+1]).  With
        ``per_utterance_stop`` every utterance gets its own H0 and stop decision, exactly like separate
        reference calls.  Returns a list of Activation whose H are row views of one (T,N) tensor."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            x = self._prep_frames(X_stacked)
            T = int(x.shape[0])
            offs = [0, T] if t_offsets is None else [int(v) for v in t_offsets]
            if offs[0] != 0 or offs[-1] != T:
                raise ValueError("t_offsets must start at 0 and end at the number of stacked frames")
            n_utt = len(offs) - 1
            ldH = _round_up(self.N, 32)      # 128-byte row pitch: TMA row segments never straddle an L2 line
            Hbuf = torch.empty((max(T, 1), ldH), dtype=torch.float32, device=self.device)
            H = Hbuf[:T, : self.N]
            if H0 is not None:
                h0 = _as_device_f32(H0, self.device)
                if tuple(h0.shape) != (T, self.N):
                    raise ValueError(f"H0 must have shape {(T, self.N)}, got {tuple(h0.shape)}")
                H.copy_(h0)
            p = self._params(beta_loss, tol, max_iter, lam, lambda_step, H0 is not None, check_every, epsilon)
            res = (SolveResult * n_utt)()
            offs_c = (C.c_int * (n_utt + 1))(*offs)
            check(L.evc_solve_batched(self._h, _ptr(x), x.stride(0), offs_c, n_utt, _ptr(Hbuf), ldH, C.byref(p),
                                      1 if per_utterance_stop else 0, res, _stream(self.device)))
        out = []
        for u in range(n_utt):
            out.append(Activation(H[offs[u]:offs[u + 1]], res[u].n_iter, bool(res[u].converged),
                                  res[u].objective, res[u].objective_at_init))
        for a in out:
            a.H_stacked = H
        if t_offsets is None:
            out[0].H = H
        return out

    def _product(self, H, target: bool) -> torch.Tensor:
        L = _lib.lib()
        with torch.cuda.device(self.device):
            h = _as_device_f32(H, self.device)
            if h.shape[1] != self.N:
                raise ValueError(f"H has {h.shape[1]} columns, the dictionary has {self.N} exemplars")
            ldH = h.stride(0)
            if h.stride(1) != 1 or ldH % 4 or h.data_ptr() % 16 or ldH < self.N:
                h = _pitched(h.contiguous(), 32)
                ldH = h.stride(0)
            T = int(h.shape[0])
            ldY = _round_up(self.F, 4)
            Ybuf = torch.empty((max(T, 1), ldY), dtype=torch.float32, device=self.device)
            fn = L.evc_convert if target else L.evc_reconstruct
            check(fn(self._h, _ptr(h), ldH, T, _ptr(Ybuf), ldY, _stream(self.device)))
        return Ybuf[:T, : self.F]

    def convert(self, H, residual=None) -> torch.Tensor:
        """Y (T,F) = H (T,N) @ B  -- 04_align_n_nmf.py:391 (``np.matmul(H_stft.T, B_stft)``).

        With ``residual`` (T,F) -- the WORLD branch, 04_align_n_nmf.py:363-373 -- the residual compensation is applied
        in the same pass: ``exp(log(H @ B) + log(r))`` with NaNs of ``r`` replaced by 0 first, IEEE semantics of the
        reference's numpy expression (0 where r == 0, NaN where r < 0)."""
        if residual is None:
            return self._product(H, True)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            h = self._prep_activations(H)
            r = _pitched(_as_device_f32(residual, self.device).contiguous())
            T = int(h.shape[0])
            if tuple(r.shape) != (T, self.F):
                raise ValueError(f"residual must have shape {(T, self.F)}, got {tuple(r.shape)}")
            ldY = _round_up(self.F, 4)
            Ybuf = torch.empty((max(T, 1), ldY), dtype=torch.float32, device=self.device)
            check(L.evc_convert_residual(self._h, _ptr(h), h.stride(0), T, _ptr(r), r.stride(0), _ptr(Ybuf), ldY,
                                         _stream(self.device)))
        return Ybuf[:T, : self.F]

    def residual(self, X, H) -> torch.Tensor:
        """R (T,F) = log(H @ A - X): the residual the reference keeps for the WORLD branch (04_align_n_nmf.py:292-294).
        NaN wherever the model undershoots the frame, like ``np.log`` of a negative number."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            x = self._prep_frames(X)
            h = self._prep_activations(H)
            T = int(h.shape[0])
            if int(x.shape[0]) != T:
                raise ValueError("X and H must have the same number of frames")
            ldR = _round_up(self.F, 4)
            Rbuf = torch.empty((max(T, 1), ldR), dtype=torch.float32, device=self.device)
            check(L.evc_residual(self._h, _ptr(x), x.stride(0), T, _ptr(h), h.stride(0), _ptr(Rbuf), ldR,
                                 _stream(self.device)))
        return Rbuf[:T, : self.F]

    def _prep_activations(self, H) -> torch.Tensor:
        h = _as_device_f32(H, self.device)
        if h.shape[1] != self.N:
            raise ValueError(f"H has {h.shape[1]} columns, the dictionary has {self.N} exemplars")
        ldH = h.stride(0)
        if h.stride(1) != 1 or ldH % 4 or h.data_ptr() % 16 or ldH < self.N:
            h = _pitched(h.contiguous(), 32)
        return h

    def reconstruct(self, H) -> torch.Tensor:
        """WH (T,F) = H (T,N) @ A  -- the model of the source frames (04_align_n_nmf.py:292)."""
        return self._product(H, False)

    def objective(self, X, H, beta_loss="kullback-leibler", epsilon=0.0) -> float:
        """sqrt(2 KL(X || H A)) or ||X - H A||_F (sklearn ``_beta_divergence(..., square_root=True)``)."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            x = self._prep_frames(X)
            h = _pitched(_as_device_f32(H, self.device).contiguous(), 32)
            out = C.c_double()
            loss = _lib.LOSS_KL if beta_loss in ("kullback-leibler", 1, 1.0) else _lib.LOSS_FROBENIUS
            check(L.evc_objective(self._h, _ptr(x), x.stride(0), int(x.shape[0]), _ptr(h), h.stride(0), loss,
                                  float(epsilon), C.byref(out), _stream(self.device)))
        return out.value

    # -- diagnostics -----------------------------------------------------------------------------------
    def profile(self, on: bool):
        check(_lib.lib().evc_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{class: (milliseconds, launches)} since the last read; classes as in include/evc.h."""
        names = ("contraction1", "reduce_ratio", "contraction2_update", "objective_init", "exchange")
        ms = (C.c_double * len(names))()
        n = (C.c_int * len(names))()
        check(_lib.lib().evc_profile_read(self._h, ms, n))
        return {names[i]: (ms[i], n[i]) for i in range(len(names))}

    # -- host staging --------------------------------------------------------------------------------
    def to_host(self, t: torch.Tensor, key: Optional[str] = None) -> np.ndarray:
        """Device tensor -> numpy (the D2H leg of the end-to-end path).  With ``key`` the copy lands in a
        cached pinned buffer that the NEXT call with the same key overwrites; without, a fresh array."""
        if key is None:
            if t.numel() * 4 >= (8 << 20):
                # large results (the 57-80 MB activation matrix) land in page-locked memory that the returned array
                # owns: a pageable D2H copy runs at a quarter of the PCIe rate, and when the array comes back as an
                # input (convert(H)) the upload is fast too.  torch's host allocator recycles the block once the
                # caller drops the array.
                buf = torch.empty(tuple(t.shape), dtype=torch.float32, pin_memory=True)
                buf.copy_(t.detach(), non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                return buf.numpy()
            return t.detach().to("cpu", torch.float32).contiguous().numpy()
        shape = tuple(t.shape)
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != shape:
            buf = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            self._pinned[key] = buf
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return buf.numpy()


# ---- resident dictionaries across calls of the numpy-level entry points ------------------------------------------------
class DictionaryCache:
    """Small LRU of device-resident dictionaries for the numpy-in / numpy-out entry points, which otherwise upload and
    prepare the same (N,F) dictionary on every call (the reference does the same with ``np.asarray(A_sp)``,
    04_align_n_nmf.py:230-246).  A hit requires the same array OBJECTS (identity, data pointer, shape, dtype), the same
    mode AND the same 64-bit content checksum -- an in-place edit of a cached array changes the checksum and rebuilds
    the dictionary, so a stale dictionary is never used (checksum of a 41 MB dictionary: ~5 ms, against ~25 ms for the
    pageable upload + operand preparation).  Not thread-safe (a handle must not be used from two threads at once)."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._entries = {}      # key -> (ExemplarDictionary, checksums, strong refs to the arrays)
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _ident(a: Optional[np.ndarray]):
        if a is None:
            return None
        return (id(a), a.__array_interface__["data"][0], a.shape, a.dtype.str, a.strides)

    @staticmethod
    def _checksum(a: Optional[np.ndarray]) -> int:
        if a is None:
            return 0
        flat = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        n8 = flat.size // 8 * 8
        total = int(flat[:n8].view(np.uint64).sum(dtype=np.uint64)) if n8 else 0
        return (total + int(flat[n8:].sum(dtype=np.uint64)) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def get(self, A: np.ndarray, B: Optional[np.ndarray], mode: str, validate=None) -> "ExemplarDictionary":
        """`validate` (optional callable) runs before a dictionary is BUILT -- a hit was validated when it was built and
        its content is unchanged, so the 40 MB finiteness scan of the reference's check_array is not repeated."""
        if not torch.cuda.is_available():
            if validate is not None:
                validate()          # input errors are reported as such even where nothing can run
            raise RuntimeError("exemplars_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        key = (self._ident(A), self._ident(B), mode, torch.cuda.current_device())
        sa = self._checksum(A)
        sums = (sa, sa if B is A else self._checksum(B))
        hit = self._entries.get(key)
        if hit is not None and hit[1] == sums and hit[0]._h.value:
            self._entries[key] = self._entries.pop(key)     # most recently used last
            self.hits += 1
            return hit[0]
        if validate is not None:
            validate()
        if hit is not None:
            self._entries.pop(key)[0].close()
        self.misses += 1
        d = ExemplarDictionary(A, B, mode=mode)
        self._entries[key] = (d, sums, (A, B))               # the arrays stay alive, so their ids cannot be reused
        while len(self._entries) > max(self.capacity, 1):
            self._entries.pop(next(iter(self._entries)))[0].close()
        return d

    def clear(self):
        for d, _s, _r in self._entries.values():
            d.close()
        self._entries.clear()


dictionary_cache = DictionaryCache()
