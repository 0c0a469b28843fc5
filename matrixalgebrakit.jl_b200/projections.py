"""project_hermitian / project_antihermitian / project_isometric and the matrix-property tests on
B200 — mirrors src/implementations/projections.jl (``check_input`` :9-34, ``initialize_output``
:38-47, implementation :60-75) and src/common/matrixproperties.jl (``isisometric`` :13-17,
``isunitary`` :26-33, ``is_left_isometric`` :53-58, ``ishermitian`` :77-84, ``isantihermitian``
:96-102).  Every operation is ONE kernel launch of libmakb200 (csrc/projections.cuh) plus, for the
tests, one small device->host read of the scalars the decision needs."""
import numpy as np
import torch

from . import _core
from .eigh import default_hermitian_tol


def _check_square(A, B=None):
    if A.dim() != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("square matrix expected")  # LinearAlgebra.checksquare -> DimensionMismatch
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    _core.dtype_code(A)
    if B is not None and B is not A:
        if tuple(B.shape) != tuple(A.shape) or B.dtype != A.dtype or not _core.is_colmajor(B):
            raise ValueError(f"B: {A.shape[0]} x {A.shape[0]} column-major matrix of A's eltype expected")


def _same_storage(A, B):
    return B is A or (B.data_ptr() == A.data_ptr() and B.stride() == A.stride())


def _project_(A, B, anti, blocksize):
    if blocksize is not None and int(blocksize) <= 0:
        raise ValueError("blocksize must be positive")  # kwarg of NativeBlocked; the kernel tiles by 32 regardless
    _check_square(A, B)
    if B is None:
        B = A  # initialize_output(project_hermitian!, A, ::NativeBlocked) = A (projections.jl:38-43)
    n = A.shape[0]
    if n == 0:
        return B
    if B is not A and not _same_storage(A, B):
        lo, hi = A.data_ptr(), A.data_ptr() + A.element_size() * (_core.ld(A) * (n - 1) + n)
        blo, bhi = B.data_ptr(), B.data_ptr() + B.element_size() * (_core.ld(B) * (n - 1) + n)
        if blo < hi and lo < bhi:
            raise ValueError("B must be A itself or must not overlap A")
    h = _core.Handle.get(A.device)
    rc = h.lib.makb200_project_hermitian(h.h, _core.dtype_code(A), int(anti), n, _core.ptr(A), _core.ld(A),
                                         _core.ptr(B), _core.ld(B))
    h.check(rc, "makb200_project_hermitian")
    return B


def project_hermitian_(A, B=None, alg=None, blocksize=None):
    """``project_hermitian!(A, [B], alg)``: B = (A + A^H)/2, in place by default (projections.jl:60-63)."""
    return _project_(A, B, False, blocksize)


def project_antihermitian_(A, B=None, alg=None, blocksize=None):
    """``project_antihermitian!(A, [B], alg)``: B = (A - A^H)/2 (projections.jl:64-67)."""
    return _project_(A, B, True, blocksize)


def _copy_input(A):
    from .qr import copy_input
    return copy_input(A)


def project_hermitian(A, alg=None, **kw):
    return project_hermitian_(_copy_input(A), None, alg, **kw)


def project_antihermitian(A, alg=None, **kw):
    return project_antihermitian_(_copy_input(A), None, alg, **kw)


def project_isometric_(A, W=None, alg=None, **kw):
    """``project_isometric!(A, W, alg)`` = the isometric factor of ``left_polar!`` with a zero-length P
    (projections.jl:69-75).  Destroys A."""
    from .polar import left_polar_
    m, n = A.shape
    if m < n:
        raise ValueError("input matrix needs at least as many rows as columns")  # projections.jl:27-28
    if W is None:
        W = _core.colmajor_empty(m, n, A.dtype, A.device)
    elif tuple(W.shape) != (m, n) or W.dtype != A.dtype:
        raise ValueError(f"W: {m} x {n} matrix of A's eltype expected")
    W, _ = left_polar_(A, (W, None), alg, **kw)
    return W


def project_isometric(A, alg=None, **kw):
    return project_isometric_(_copy_input(A), None, alg, **kw)


def hermitian_props(A, anti=False):
    """(||vanishing part||_F, max|A_ij|, ||A||_F, exact mismatches) from one pass over A on the device."""
    _check_square(A)
    n = A.shape[0]
    if n == 0:
        return 0.0, 0.0, 0.0, 0
    h = _core.Handle.get(A.device)
    out = torch.empty(4, dtype=torch.float64, device=A.device)
    rc = h.lib.makb200_hermitian_props(h.h, _core.dtype_code(A), int(anti), n, _core.ptr(A), _core.ld(A),
                                       _core.ptr(out))
    h.check(rc, "makb200_hermitian_props")
    d2, mx, f2, bad = out.tolist()  # the one device->host read the boolean answer needs
    return float(np.sqrt(d2)), float(mx), float(np.sqrt(f2)), int(bad)


def _is_herm(A, anti, atol, rtol):
    if A.dim() != 2 or A.shape[0] != A.shape[1]:
        return False
    defect, mx, fro, bad = hermitian_props(A, anti)
    if atol == 0 and rtol == 0:
        return bad == 0                                  # ishermitian_exact (matrixproperties.jl:86-88,115-150)
    if atol is None:
        atol = default_hermitian_tol(mx) if A.shape[0] else 0.0
    bound = max(atol, rtol * fro) if rtol > 0 else atol   # strided_ishermitian_approx (:152-158)
    return defect <= bound


def ishermitian(A, atol=0, rtol=0):
    """``ishermitian(A; atol, rtol)`` (matrixproperties.jl:77-84): exact entry-wise test when both
    tolerances are zero, else ||(A - A^H)/2||_F <= max(atol, rtol ||A||_F); ``atol=None`` selects the
    reference's ``default_hermitian_tol`` like ``strided_ishermitian_approx``'s default."""
    return _is_herm(A, False, atol, rtol)


def isantihermitian(A, atol=0, rtol=0):
    """``isantihermitian(A; atol, rtol)`` (matrixproperties.jl:96-102)."""
    return _is_herm(A, True, atol, rtol)


def defaulttol(A):
    """``defaulttol(x) = eps(real(float(one(eltype(x)))))^(2/3)`` (src/common/defaults.jl:10)."""
    return float(np.finfo(np.float64).eps ** (2.0 / 3.0))


def _isometric(A, right, atol, rtol):
    from .gemm import gemm_
    if rtol is None:
        rtol = defaulttol(A)
    if A.dim() != 2 or not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    _core.dtype_code(A)
    m, n = A.shape
    k, inner = (m, n) if right else (n, m)   # P is k x k, contraction over `inner`
    if k == 0:
        return True
    h = _core.Handle.get(A.device)
    P = _core.colmajor_empty(k, k, A.dtype, A.device)
    if inner == 0:
        P.zero_()
    elif right:
        gemm_(P, A, A, opa="N", opb="C")
    else:
        gemm_(P, A, A, opa="C", opb="N")
    out = torch.empty(2, dtype=torch.float64, device=A.device)
    rc = h.lib.makb200_gram_defect(h.h, _core.dtype_code(A), k, _core.ptr(P), _core.ld(P), _core.ptr(out))
    h.check(rc, "makb200_gram_defect")
    p2, d2 = out.tolist()   # the one device->host read the boolean answer needs
    return float(np.sqrt(d2)) <= max(atol, rtol * float(np.sqrt(p2)))


def is_left_isometric(A, atol=0, rtol=None):
    """``is_left_isometric`` (matrixproperties.jl:53-58): P = A^H A on the DMMA GEMM, then
    ||P - I||_F <= max(atol, rtol ||P||_F) from one reduction kernel."""
    return _isometric(A, False, atol, rtol)


def is_right_isometric(A, atol=0, rtol=None):
    """``is_right_isometric(A) = is_left_isometric(A')`` (matrixproperties.jl:67): P = A A^H."""
    return _isometric(A, True, atol, rtol)


def isisometric(A, side="left", atol=0, rtol=None):
    """``isisometric(A; side)`` (matrixproperties.jl:13-17)."""
    if side == "left":
        return is_left_isometric(A, atol, rtol)
    if side == "right":
        return is_right_isometric(A, atol, rtol)
    raise ValueError(f"Invalid isometry side: {side}")


def isunitary(A, atol=0, rtol=None):
    """``isunitary(A::AbstractMatrix)`` (matrixproperties.jl:30-33)."""
    if A.shape[0] != A.shape[1]:
        return False
    return is_left_isometric(A, atol, rtol)


def _tri_init_(A, mode):
    if A.dim() != 2 or not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    m, n = A.shape
    if m == 0 or n == 0:
        return A
    h = _core.Handle.get(A.device)
    rc = h.lib.makb200_tri_init(h.h, _core.dtype_code(A), mode, m, n, _core.ptr(A), _core.ld(A))
    h.check(rc, "makb200_tri_init")
    return A


def one_(A):
    """``one!(A)`` (src/common/initialization.jl:11-16): rectangular identity, one launch."""
    return _tri_init_(A, 0)


def uppertriangular_(A):
    """``uppertriangular!(A)`` (initialization.jl:18-26): zero strictly below the diagonal, one launch."""
    return _tri_init_(A, 1)


def lowertriangular_(A):
    """``lowertriangular!(A)`` (initialization.jl:28-36): zero strictly above the diagonal, one launch."""
    return _tri_init_(A, 2)
