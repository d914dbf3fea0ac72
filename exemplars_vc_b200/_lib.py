"""ctypes binding of libevc_b200.so (include/evc.h).  There is NO fallback: if the CUDA library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# EVC_LIB_PATH: load another build of the same library (A/B runs of compile-time variants); still no fallback.
LIB_PATH = os.environ.get("EVC_LIB_PATH") or os.path.join(HERE, "libevc_b200.so")

EVC_OK, EVC_ERR_INVALID_ARGUMENT, EVC_ERR_CUDA, EVC_ERR_UNSUPPORTED, EVC_ERR_VALUE, EVC_ERR_COMM = range(6)
MODE_FP32, MODE_3XTF32, MODE_TF32, MODE_BF16 = range(4)
MODES = {"fp32": MODE_FP32, "3xtf32": MODE_3XTF32, "tf32": MODE_TF32, "bf16": MODE_BF16}
LOSS_KL, LOSS_FROBENIUS = 1, 2
INIT_SKLEARN, INIT_GIVEN = 0, 1


class SolveParams(C.Structure):
    _fields_ = [("loss", C.c_int), ("init", C.c_int), ("max_iter", C.c_int), ("check_every", C.c_int),
                ("tol", C.c_float), ("lam", C.c_float), ("lambda_step", C.c_float), ("epsilon", C.c_float)]


class SolveResult(C.Structure):
    _fields_ = [("n_iter", C.c_int), ("converged", C.c_int), ("objective", C.c_double),
                ("objective_at_init", C.c_double)]


class EvcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libevc_b200 status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def lib():
    """Load the shared library once.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    L.evc_version.restype = ip
    L.evc_last_error_string.restype = C.c_char_p
    L.evc_kernel_launch_count.restype = C.c_longlong
    L.evc_last_enqueue_ms.restype = C.c_double
    L.evc_mma_passes_per_product.argtypes = [C.c_int]
    L.evc_mma_passes_per_product.restype = C.c_int
    L.evc_default_params.argtypes = [C.POINTER(SolveParams)]
    L.evc_default_params.restype = None
    L.evc_dict_create.argtypes = [vp, ip, vp, ip, ip, ip, ip, vp, C.POINTER(vp)]
    L.evc_dict_destroy.argtypes = [vp]
    L.evc_dict_info.argtypes = [vp, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
    L.evc_dict_colsum.argtypes = [vp, vp, vp]
    L.evc_solve.argtypes = [vp, vp, ip, ip, vp, ip, C.POINTER(SolveParams), C.POINTER(SolveResult), vp]
    L.evc_solve_batched.argtypes = [vp, vp, ip, C.POINTER(ip), ip, vp, ip, C.POINTER(SolveParams), ip,
                                    C.POINTER(SolveResult), vp]
    L.evc_convert.argtypes = [vp, vp, ip, ip, vp, ip, vp]
    L.evc_reconstruct.argtypes = [vp, vp, ip, ip, vp, ip, vp]
    L.evc_objective.argtypes = [vp, vp, ip, ip, vp, ip, ip, fp, C.POINTER(C.c_double), vp]
    L.evc_factorize_convert_host.argtypes = [vp, vp, ip, ip, vp, ip, vp, ip, C.POINTER(SolveParams),
                                             C.POINTER(SolveResult), vp]
    L.evc_residual.argtypes = [vp, vp, ip, ip, vp, ip, vp, ip, vp]
    L.evc_convert_residual.argtypes = [vp, vp, ip, ip, vp, ip, vp, ip, vp]
    L.evc_griffin_lim.argtypes = [vp, ip, ip, ip, ip, ip, vp, vp, vp, vp, vp]
    L.evc_stft.argtypes = [vp, C.c_longlong, ip, ip, vp, vp, vp]
    L.evc_istft.argtypes = [vp, ip, ip, ip, vp, vp, vp]
    L.evc_dtw.argtypes = [vp, vp, vp, vp, ip, ip, ip, vp, vp, vp, vp, vp, vp, vp, vp]
    L.evc_gather_stack.argtypes = [vp, ip, ip, ip, vp, vp, vp, ip, ip, vp, ip, vp]
    L.evc_profile_enable.argtypes = [vp, ip]
    L.evc_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(ip)]
    L.evc_comm_unique_id.argtypes = [C.c_char_p]
    L.evc_comm_create.argtypes = [C.c_char_p, ip, ip, C.POINTER(vp)]
    L.evc_comm_destroy.argtypes = [vp]
    L.evc_dict_attach_comm.argtypes = [vp, vp, ip]
    L.evc_p2p_alloc.argtypes = [vp, ip, C.c_char_p]
    L.evc_p2p_attach.argtypes = [vp, C.c_char_p, ip, ip]
    L.evc_p2p_detach.argtypes = [vp]
    for name in ("evc_dict_create", "evc_dict_destroy", "evc_dict_info", "evc_dict_colsum", "evc_solve",
                 "evc_solve_batched", "evc_convert", "evc_reconstruct", "evc_objective", "evc_factorize_convert_host",
                 "evc_gather_stack", "evc_dtw", "evc_residual", "evc_convert_residual", "evc_griffin_lim", "evc_stft",
                 "evc_istft", "evc_profile_enable", "evc_profile_read", "evc_comm_unique_id", "evc_comm_create", "evc_comm_destroy", "evc_dict_attach_comm", "evc_p2p_alloc",
                 "evc_p2p_attach", "evc_p2p_detach"):
        getattr(L, name).restype = ip
    _lib = L
    return L


def check(status: int):
    """Map a status code to the exception the reference's operator raises in the same situation."""
    if status == EVC_OK:
        return
    msg = lib().evc_last_error_string().decode("utf-8", "replace")
    if status == EVC_ERR_VALUE:
        raise ValueError(msg)                      # sklearn _nmf.py:61-76
    if status == EVC_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if status == EVC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise EvcError(status, msg)


def kernel_launch_count() -> int:
    return int(lib().evc_kernel_launch_count())
