"""eigh_full / eigh_vals / eigh_trunc on B200 — mirrors src/implementations/eigh.jl:
``check_hermitian`` (:11-18), ``check_input`` (:20-30), ``initialize_output`` (:65-70),
``eigh_full_<alg>!(driver, A, DV; fixgauge)`` (:150-156), ``eigh_trunc!`` (:165-169)."""
import ctypes as C

import numpy as np
import torch

from . import _core
from .algorithms import (Algorithm, TruncatedAlgorithm, default_fixgauge, resolve_driver, select_algorithm)
from .truncation import findtruncated, select_truncation, truncation_error_


class DomainError(ValueError):
    """Julia ``DomainError`` (eigh.jl:15-16)."""


def _real_dtype(dt):
    return torch.float64


def hermitian_defect(A):
    """(||(A - A^H)/2||_F, max|A_ij|) computed on the device in one pass."""
    h = _core.Handle.get(A.device)
    out = torch.zeros(2, dtype=torch.float64, device=A.device)
    rc = h.lib.makb200_hermitian_defect(h.h, _core.dtype_code(A), A.shape[0], _core.ptr(A), _core.ld(A),
                                        _core.ptr(out))
    h.check(rc, "makb200_hermitian_defect")
    d2, mx = out.tolist()  # one D2H read, like the reference's `norm` on a CuArray
    return float(np.sqrt(d2)), float(mx)


def default_hermitian_tol(maxabs):
    """``default_hermitian_tol`` = eps(norm(A, Inf))^(3/4) (src/common/defaults.jl:44)."""
    return float(np.spacing(maxabs) ** 0.75)


def check_hermitian(A, atol=None):
    if A.dim() != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("square matrix expected")  # checksquare -> DimensionMismatch
    if A.shape[0] == 0:
        return
    defect, mx = hermitian_defect(A)
    tol = default_hermitian_tol(mx) if atol is None else atol
    if not defect <= tol:
        raise DomainError("Hermitian matrix was expected. Use `project_hermitian` to project onto the "
                          "nearest hermitian matrix.")


def initialize_output(A):
    """(D, V): D is the vector of the Diagonal (real), V n x n (eigh.jl:65-70)."""
    n = A.shape[0]
    return (torch.empty(n, dtype=torch.float64, device=A.device), _core.colmajor_empty(n, n, A.dtype, A.device))


def check_input(A, DV, alg=None, check=True):
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    _core.dtype_code(A)
    if check:
        atol = alg.get("hermitian_tol") if isinstance(alg, Algorithm) else None
        check_hermitian(A, atol)
    D, V = DV
    n = A.shape[0]
    if D.dim() != 1 or D.shape[0] != n or D.dtype != torch.float64:
        raise ValueError("D: real vector of length n expected")
    if V is not None and (tuple(V.shape) != (n, n) or V.dtype != A.dtype or not _core.is_colmajor(V)):
        raise ValueError("V: n x n column-major matrix of A's eltype expected")


def _heevd_(A, D, V, fixgauge):
    h = _core.Handle.get(A.device)
    n = A.shape[0]
    dt = _core.dtype_code(A)
    lw = h.lib.makb200_eigh_worksize(h.h, dt, n)
    work = h.workspace(lw)
    rc = h.lib.makb200_eigh(h.h, dt, int(bool(fixgauge)), n, _core.ptr(A), _core.ld(A), _core.ptr(D), _core.ptr(V),
                            _core.ld(V) if V is not None else 0, _core.ptr(work), work.numel(), C.c_void_p(0))
    h.check(rc, "makb200_eigh")


def _alg_ok(alg):
    # heevd!(::B200) and heevr!(::B200) are the same kernels (tridiagonalisation + tridiagonal D&C): the tag names the
    # contract (ascending values, gauge-fixed vectors), the driver the implementation.  So the BASELINE-named alias
    # LAPACK_MultipleRelativelyRobustRepresentations(driver = B200()) works; heev!/heevj! are not provided and throw.
    if not isinstance(alg, Algorithm) or alg.name not in ("DivideAndConquer", "RobustRepresentations"):
        raise ValueError(f"eigh: algorithm {alg} is not provided by the B200 driver "
                         "(DivideAndConquer and RobustRepresentations are)")
    resolve_driver(alg.get("driver"), None)


def eigh_full_(A, DV=None, alg=None, **kw):
    """``eigh_full!(A, (D,V), alg)`` (eigh.jl:123-127). Destroys A (upper triangle is read)."""
    alg = select_algorithm("eigh_full", A, alg, **kw)
    _alg_ok(alg)
    if DV is None:
        DV = initialize_output(A)
    check_input(A, DV, alg)
    fixgauge = alg.get("fixgauge", default_fixgauge())
    _heevd_(A, DV[0], DV[1], fixgauge)
    return DV


def copy_input(A):
    if not A.dtype.is_floating_point and not A.dtype.is_complex:
        A = A.to(torch.float64)
    out = _core.colmajor_empty(A.shape[0], A.shape[1], A.dtype, A.device)
    out.copy_(A)
    return out


def eigh_full(A, alg=None, **kw):
    return eigh_full_(copy_input(A), None, alg, **kw)


def eigh_vals_(A, D=None, alg=None, **kw):
    """``eigh_vals!`` (eigh.jl:157-161)."""
    alg = select_algorithm("eigh_vals", A, alg, **kw)
    _alg_ok(alg)
    n = A.shape[0]
    if D is None:
        D = torch.empty(n, dtype=torch.float64, device=A.device)
    check_input(A, (D, None), alg)
    _heevd_(A, D, None, False)   # V = NULL: job 'N' (zero-length V of the reference, yalapack.jl:1192-1195, 1293-1298)
    return D


def eigh_vals(A, alg=None, **kw):
    return eigh_vals_(copy_input(A), None, alg, **kw)


def eigh_trunc_(A, DV=None, alg=None, trunc=None, **kw):
    """``eigh_trunc!`` (eigh.jl:165-169): full decomposition, then keep the index set."""
    if isinstance(alg, TruncatedAlgorithm):
        if trunc is not None:
            raise ValueError("`trunc` can't be specified when `alg` is a `TruncatedAlgorithm`")
        talg = alg
    else:
        talg = TruncatedAlgorithm(select_algorithm("eigh_full", A, alg, **kw), select_truncation(trunc))
    D, V = eigh_full_(A, DV, talg.alg)
    ind = findtruncated(D, talg.trunc)
    Dt = D[ind].clone()
    Vt = _core.colmajor_empty(V.shape[0], len(ind), V.dtype, V.device)
    Vt.copy_(V[:, ind])
    eps_ = truncation_error_(D, ind)
    return Dt, Vt, eps_


def eigh_trunc(A, alg=None, trunc=None, **kw):
    return eigh_trunc_(copy_input(A), None, alg, trunc, **kw)


class BatchedEighPlan:
    """Argument arrays of a batched ``eigh_full!`` built once; ``run()`` is one C-ABI call
    (one CTA per block, two-sided Jacobi in shared memory; large blocks routed per block)."""

    def __init__(self, As, DVs=None, fixgauge=True, check=False):
        self.As = As
        self.h = _core.Handle.get(As[0].device)
        self.dt = _core.dtype_code(As[0])
        self.DVs = DVs if DVs is not None else [initialize_output(A) for A in As]
        for A, DV in zip(As, self.DVs):
            check_input(A, DV, None, check=check)
        b = self.b = len(As)
        IA, VP = C.c_int * b, C.c_void_p * b
        self.fixgauge = int(bool(fixgauge))
        self.n = IA(*[A.shape[0] for A in As])
        self.lda = IA(*[_core.ld(A) for A in As])
        self.ldv = IA(*[_core.ld(V) for _, V in self.DVs])
        self.Ap = VP(*[A.data_ptr() for A in As])
        self.Wp = VP(*[D.data_ptr() for D, _ in self.DVs])
        self.Vp = VP(*[V.data_ptr() for _, V in self.DVs])
        self.lw = self.h.lib.makb200_eigh_batched_worksize(self.h.h, self.dt, b, self.n)

    def run(self):
        h = _core.Handle.get(self.As[0].device)
        work = h.workspace(self.lw)
        rc = h.lib.makb200_eigh_batched(h.h, self.dt, self.fixgauge, self.b, self.n, self.Ap, self.lda, self.Wp,
                                        self.Vp, self.ldv, C.c_void_p(0), _core.ptr(work), work.numel())
        h.check(rc, "makb200_eigh_batched")
        return self.DVs


def eigh_full_batched_(As, DVs=None, fixgauge=True, check=True):
    """Batched ``eigh_full!`` over a list of Hermitian blocks (per-block semantics of eigh.jl:123-156;
    ``check`` runs the reference's Hermitian pre-check per block)."""
    if len(As) == 0:
        return []
    return BatchedEighPlan(As, DVs, fixgauge, check).run()


def eigh_vals_batched_(As, Ds=None, check=True):
    """Batched ``eigh_vals!`` (eigh.jl:157-161 per block, job 'N'): one C-ABI call with V = NULL; blocks
    beyond the one-CTA kernel take the tridiagonalisation + Sturm K-section path."""
    if len(As) == 0:
        return []
    h = _core.Handle.get(As[0].device)
    dt = _core.dtype_code(As[0])
    b = len(As)
    if Ds is None:
        Ds = [torch.empty(A.shape[0], dtype=torch.float64, device=A.device) for A in As]
    for A, D in zip(As, Ds):
        check_input(A, (D, None), None, check=check)
    IA, VP = C.c_int * b, C.c_void_p * b
    n = IA(*[A.shape[0] for A in As])
    lda = IA(*[_core.ld(A) for A in As])
    Ap, Wp = VP(*[A.data_ptr() for A in As]), VP(*[D.data_ptr() for D in Ds])
    work = h.workspace(h.lib.makb200_eigh_batched_worksize(h.h, dt, b, n))
    rc = h.lib.makb200_eigh_batched(h.h, dt, 0, b, n, Ap, lda, Wp, None, None, C.c_void_p(0), _core.ptr(work),
                                    work.numel())
    h.check(rc, "makb200_eigh_batched")
    return Ds
