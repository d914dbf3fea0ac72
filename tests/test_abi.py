"""The C-ABI library loads and exports every symbol include/evc.h declares (no GPU, no compute calls)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "evc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(evc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for need in ("evc_dict_create", "evc_dict_destroy", "evc_solve", "evc_solve_batched", "evc_convert",
                 "evc_reconstruct", "evc_objective", "evc_factorize_convert_host", "evc_comm_create",
                 "evc_dict_attach_comm", "evc_last_error_string", "evc_version"):
        assert need in syms


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/evc.h but not exported by {built_lib}"


def test_version_defaults_and_struct_layout(built_lib):
    from exemplars_vc_b200 import _lib
    L = _lib.lib()
    assert L.evc_version() == 100
    p = _lib.SolveParams()
    L.evc_default_params(C.byref(p))
    # the reference's defaults: KL signature default, max_iter=150 (04_align_n_nmf.py:213), check every 10, tol 1e-4
    assert (p.loss, p.init, p.max_iter, p.check_every) == (_lib.LOSS_KL, _lib.INIT_SKLEARN, 150, 10)
    assert abs(p.tol - 1e-4) < 1e-10 and p.lam == 0.0 and p.lambda_step == 0.0
    assert C.sizeof(_lib.SolveParams) == 32 and C.sizeof(_lib.SolveResult) == 24


def test_mma_passes_per_product_is_reported(built_lib):
    """bench.py's executed-TFLOP/s figure comes from the library: 3 bf16 MMAs per product in the fp32-accurate split
    mode (x = x1 + x2 in bf16, three of the four partial products), 1 MMA in the fast modes, 0 for the FFMA mode."""
    from exemplars_vc_b200 import _lib
    L = _lib.lib()
    assert L.evc_mma_passes_per_product(_lib.MODE_3XTF32) == 3
    assert L.evc_mma_passes_per_product(_lib.MODE_TF32) == 1
    assert L.evc_mma_passes_per_product(_lib.MODE_BF16) == 1
    assert L.evc_mma_passes_per_product(_lib.MODE_FP32) == 0


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from exemplars_vc_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p(0)
    assert L.evc_dict_create(None, 4, None, 0, 4, 4, 0, None, C.byref(h)) == _lib.EVC_ERR_INVALID_ARGUMENT
    assert b"bad shape" in L.evc_last_error_string()
    assert L.evc_dict_create(C.c_void_p(16), 4, None, 0, 4, 4, 99, None, C.byref(h)) == _lib.EVC_ERR_INVALID_ARGUMENT
    assert L.evc_dict_destroy(None) == _lib.EVC_OK
    assert L.evc_dict_info(None, None, None, None, None) == _lib.EVC_ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        _lib.check(_lib.EVC_ERR_INVALID_ARGUMENT)


def test_sass_has_blackwell_tensor_and_tma_instructions(built_lib):
    """tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
