"""ctypes binding of include/makb200.h — the same symbols the Julia extension ccalls.

The product path has NO CPU fallback: if the shared library is missing or cannot be loaded
this module raises, and every operator fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmakb200.so")

F64, C128 = 0, 1
QR_COMPACT, QR_FULL = 0, 1
OP_N, OP_T, OP_C = 0, 1, 2
ERR_CUDA, ERR_WORKSPACE, ERR_NOCONV, ERR_NCCL = 1000, 1001, 1002, 1003

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_ip = C.POINTER(C.c_int)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); kept in one table so tests can check every declared symbol
SIGNATURES = {
    "makb200_version": (_i, []),
    "makb200_create": (_i, [C.POINTER(_vp), _i]),
    "makb200_destroy": (_i, [_vp]),
    "makb200_set_stream": (_i, [_vp, _vp]),
    "makb200_last_error": (C.c_char_p, [_vp]),
    "makb200_launch_count": (C.c_ulonglong, []),
    "makb200_kernel_timing": (_i, [_i]),
    "makb200_kernel_time": (_i, [_i, C.POINTER(C.c_double), _ip]),
    "makb200_gemm_flops": (C.c_double, []),
    "makb200_gemm": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i]),
    "makb200_geqrf_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_geqrf": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz]),
    "makb200_orgqr_worksize": (_sz, [_vp, _i, _i, _i, _i]),
    "makb200_orgqr": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz]),
    "makb200_ormqr_worksize": (_sz, [_vp, _i, _i, _i, _i]),
    "makb200_ormqr": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz]),
    "makb200_qr_worksize": (_sz, [_vp, _i, _i, _i, _i]),
    "makb200_qr": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _sz]),
    "makb200_qr_batched_worksize": (_sz, [_vp, _i, _i, _ip, _ip]),
    "makb200_qr_batched": (_i, [_vp, _i, _i, _ip, _ip, _vpp, _ip, _vpp, _ip, _vpp, _ip, _vp, _vp, _sz]),
    "makb200_qr_batched_plan_create": (_i, [_vp, _i, _i, _ip, _ip, _vpp, _ip, _vpp, _ip, _vpp, _ip, _vp, _sz,
                                            C.POINTER(C.c_void_p)]),
    "makb200_qr_batched_plan_run": (_i, [_vp, _vp, _vp]),
    "makb200_qr_batched_plan_destroy": (_i, [_vp]),
    "makb200_hermitian_defect": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "makb200_project_hermitian": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i]),
    "makb200_hermitian_props": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "makb200_tri_init": (_i, [_vp, _i, _i, _i, _i, _vp, _i]),
    "makb200_fro2": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "makb200_gram_defect": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "makb200_eigh_worksize": (_sz, [_vp, _i, _i]),
    "makb200_eigh": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "makb200_stedc_worksize": (_sz, [_vp, _i]),
    "makb200_stedc": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "makb200_polar_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_polar_qdwh": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, C.c_double, _i, _vp, _sz, _ip, _vp]),
    "makb200_svd_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_svd": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, C.c_double, _vp, _sz, _vp]),
    "makb200_svd_leading": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, C.c_double, _vp, _sz, _vp]),
    "makb200_tsqr_local_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_tsqr_local": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "makb200_sbr_chase_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_sbr_chase": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _sz]),
    "makb200_sy2sb_worksize": (_sz, [_vp, _i, _i, _i]),
    "makb200_sy2sb": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz]),
    "makb200_sbr_apply_q2_worksize": (_sz, [_vp, _i, _i, _i, _i, _i]),
    "makb200_sbr_apply_q2": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _sz]),
    "makb200_tsqr_local_ex": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _sz, _vp]),
    "makb200_eigh_batched_worksize": (_sz, [_vp, _i, _i, _ip]),
    "makb200_eigh_batched": (_i, [_vp, _i, _i, _i, _ip, _vpp, _ip, _vpp, _vpp, _ip, _vp, _vp, _sz]),
    "makb200_svd_batched_worksize": (_sz, [_vp, _i, _i, _ip, _ip]),
    "makb200_svd_batched": (_i, [_vp, _i, _i, _i, _ip, _ip, _vpp, _ip, _vpp, _vpp, _ip, _vpp, _ip, _vp, _vp, _sz]),
    "makb200_trunc_select_batched_worksize": (_sz, [_vp, _i]),
    "makb200_trunc_select_batched": (_i, [_vp, _i, _ip, _vpp, _vp, _ip, _vp, _vp, _vp, _sz]),
    "makb200_adjoint": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i]),
    "makb200_gauge_columns": (_i, [_vp, _i, _i, _i, _vp, _i]),
    "makb200_nccl_unique_id": (_i, [_vp]),
    "makb200_comm_create": (_i, [C.POINTER(_vp), _i, _i, _vp]),
    "makb200_comm_destroy": (_i, [_vp]),
    "makb200_tsqr_worksize": (_sz, [_vp, _i, _i, _i, _i]),
    "makb200_tsqr": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Load libmakb200.so (built in-tree by build.py). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"makb200: native library {LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; "
            "g.build()'`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
