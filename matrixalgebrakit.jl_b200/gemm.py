"""``mul!``-style GEMM on the DMMA kernel (replaces cuBLAS gemm behind ``mul!`` on CuArray,
implementations/polar.jl:63,88)."""
import ctypes as C

from . import _core, _lib

_OPS = {"N": _lib.OP_N, "T": _lib.OP_T, "C": _lib.OP_C}


def _scalar(dt, v):
    if dt == _lib.F64:
        return (C.c_double * 1)(float(v))
    v = complex(v)
    return (C.c_double * 2)(v.real, v.imag)


def gemm_(C_out, A, B, alpha=1.0, beta=0.0, opa="N", opb="N"):
    """C = alpha*op(A)*op(B) + beta*C on column-major device matrices."""
    h = _core.Handle.get(C_out.device)
    dt = _core.dtype_code(C_out)
    m, n = C_out.shape
    k = A.shape[1] if opa == "N" else A.shape[0]
    al, be = _scalar(dt, alpha), _scalar(dt, beta)
    rc = h.lib.makb200_gemm(h.h, dt, _OPS[opa], _OPS[opb], m, n, k, C.cast(al, C.c_void_p), _core.ptr(A),
                            _core.ld(A), _core.ptr(B), _core.ld(B), C.cast(be, C.c_void_p), _core.ptr(C_out),
                            _core.ld(C_out))
    h.check(rc, "makb200_gemm")
    return C_out
