"""qr_compact / qr_full on B200 — mirrors src/implementations/qr.jl of the reference:
``check_input`` (:9-29), ``initialize_output`` (:63-75), ``qr_householder!(driver, A, Q, R;
positive, pivoted, blocksize)`` (:132-188).  Python has no ``!``: the in-place, preallocated-
output entry points carry a trailing underscore (``qr_compact_(A, (Q, R), alg)``)."""
import ctypes as C

import torch

from . import _core, _lib
from .algorithms import Algorithm, B200, resolve_driver, select_algorithm


def _check_matrix(A, name="A"):
    if not isinstance(A, torch.Tensor) or A.dim() != 2:
        raise TypeError(f"{name}: 2-d device tensor expected")
    if not _core.is_colmajor(A):
        raise ValueError(f"{name}: column-major matrix with unit stride in dim 1 expected (chkstride1)")
    _core.dtype_code(A)


def initialize_output(f, A, alg=None):
    """``initialize_output(qr_full!/qr_compact!, A, alg)`` (qr.jl:63-75)."""
    m, n = A.shape
    k = min(m, n)
    if f == "qr_full":
        return (_core.colmajor_empty(m, m, A.dtype, A.device), _core.colmajor_empty(m, n, A.dtype, A.device))
    return (_core.colmajor_empty(m, k, A.dtype, A.device), _core.colmajor_empty(k, n, A.dtype, A.device))


def check_input(f, A, QR, alg=None):
    """``check_input`` (qr.jl:9-29): DimensionMismatch -> ValueError. A zero-length R means
    "R not requested" (qr.jl:15,26)."""
    _check_matrix(A)
    m, n = A.shape
    k = min(m, n)
    Q, R = QR
    _check_matrix(Q, "Q")
    nq = m if f == "qr_full" else k
    if tuple(Q.shape) != (m, nq):
        raise ValueError(f"Q: size {tuple(Q.shape)} != {(m, nq)}")
    if Q.dtype != A.dtype:
        raise TypeError("Q: eltype mismatch")
    if R is not None and R.numel() > 0:
        _check_matrix(R, "R")
        if tuple(R.shape) != (nq, n):
            raise ValueError(f"R: size {tuple(R.shape)} != {(nq, n)}")
        if R.dtype != A.dtype:
            raise TypeError("R: eltype mismatch")


def qr_householder_(A, Q, R, driver=None, positive=True, pivoted=False, blocksize=0, mode=_lib.QR_COMPACT):
    """``qr_householder!(::B200, A, Q, R; positive, pivoted, blocksize)``: one fused C-ABI call
    (factorization, R extraction and Q formation; the gauge is built into the reflectors)."""
    driver = resolve_driver(driver, A)
    assert isinstance(driver, B200)
    # capability negatives are thrown, not ignored (qr.jl:140-145; tests assert the throw)
    if pivoted:
        raise ValueError(f"{driver} does not provide a pivoted QR decomposition")
    if blocksize not in (0, 1):
        # the kernels choose their own (two-level) blocking; a user block size is not an option
        raise ValueError(f"{driver} does not provide a blocked QR decomposition with user block size")
    if Q.data_ptr() == A.data_ptr() and A.numel() > 0:
        raise ValueError("inplace Q is not supported by the B200 driver")
    m, n = A.shape
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    compute_r = R is not None and R.numel() > 0
    lw = h.lib.makb200_qr_worksize(h.h, dt, mode, m, n)
    work = h.workspace(lw)
    rc = h.lib.makb200_qr(h.h, dt, mode, int(bool(positive)), m, n, _core.ptr(A), _core.ld(A), _core.ptr(Q),
                          _core.ld(Q), _core.ptr(R) if compute_r else C.c_void_p(0),
                          _core.ld(R) if compute_r else 0, _core.ptr(work), work.numel())
    h.check(rc, "makb200_qr")
    return Q, R


def _alg_kwargs(alg):
    if not isinstance(alg, Algorithm) or alg.name != "Householder":
        raise ValueError(f"qr: algorithm {alg} is not available on the B200 driver")
    return {k: v for k, v in alg.kwargs.items()}


def qr_compact_(A, QR=None, alg=None, **kw):
    """``qr_compact!(A, (Q,R), alg)`` (qr.jl:110-113). Destroys A; returns the objects given."""
    alg = select_algorithm("qr_compact", A, alg, **kw)
    if QR is None:
        QR = initialize_output("qr_compact", A, alg)
    check_input("qr_compact", A, QR, alg)
    return qr_householder_(A, QR[0], QR[1], mode=_lib.QR_COMPACT, **_alg_kwargs(alg))


def qr_full_(A, QR=None, alg=None, **kw):
    """``qr_full!(A, (Q,R), alg)`` (qr.jl:106-109)."""
    alg = select_algorithm("qr_full", A, alg, **kw)
    if QR is None:
        QR = initialize_output("qr_full", A, alg)
    check_input("qr_full", A, QR, alg)
    return qr_householder_(A, QR[0], QR[1], mode=_lib.QR_FULL, **_alg_kwargs(alg))


def copy_input(A):
    """``copy_input`` (qr.jl:3-4): fresh column-major float copy; integer input is converted."""
    if not A.dtype.is_floating_point and not A.dtype.is_complex:
        A = A.to(torch.float64)
    out = _core.colmajor_empty(A.shape[0], A.shape[1], A.dtype, A.device)
    out.copy_(A)
    return out


def qr_compact(A, alg=None, **kw):
    """out-of-place ``qr_compact(A; alg, kw...)``: input untouched (algorithms.jl:355-370)."""
    return qr_compact_(copy_input(A), None, alg, **kw)


def qr_full(A, alg=None, **kw):
    return qr_full_(copy_input(A), None, alg, **kw)


class BatchedQRPlan:
    """Argument arrays of a batched ``qr_compact!`` built once (block sizes, leading dimensions and
    device pointers are fixed for a block-sparse tensor across many calls); ``run()`` is a single
    C-ABI call."""

    def __init__(self, As, QRs=None):
        self.As = As
        self.h = _core.Handle.get(As[0].device)
        self.dt = _core.dtype_code(As[0])
        self.QRs = QRs if QRs is not None else [initialize_output("qr_compact", A) for A in As]
        b = self.b = len(As)
        IA, VP = C.c_int * b, C.c_void_p * b
        for A, QR in zip(As, self.QRs):
            check_input("qr_compact", A, QR)
        self.m = IA(*[A.shape[0] for A in As])
        self.n = IA(*[A.shape[1] for A in As])
        self.lda = IA(*[_core.ld(A) for A in As])
        self.ldq = IA(*[_core.ld(Q) for Q, _ in self.QRs])
        self.ldr = IA(*[_core.ld(R) if R is not None and R.numel() else 0 for _, R in self.QRs])
        self.Ap = VP(*[A.data_ptr() for A in As])
        self.Qp = VP(*[Q.data_ptr() for Q, _ in self.QRs])
        self.Rp = VP(*[(R.data_ptr() if R is not None and R.numel() else 0) for _, R in self.QRs])
        self.lw = self.h.lib.makb200_qr_batched_worksize(self.h.h, self.dt, b, self.m, self.n)
        # the plan owns its workspace (descriptors live there between runs)
        self.work = torch.empty(max(int(self.lw), 256), dtype=torch.uint8, device=As[0].device)
        self.plan = C.c_void_p(0)
        rc = self.h.lib.makb200_qr_batched_plan_create(self.h.h, self.dt, b, self.m, self.n, self.Ap, self.lda, self.Qp,
                                                       self.ldq, self.Rp, self.ldr, _core.ptr(self.work),
                                                       self.work.numel(), C.byref(self.plan))
        self.h.check(rc, "makb200_qr_batched_plan_create")

    def run(self):
        rc = self.h.lib.makb200_qr_batched_plan_run(self.h.h, self.plan, C.c_void_p(0))
        self.h.check(rc, "makb200_qr_batched_plan_run")
        return self.QRs

    def __del__(self):
        try:
            if getattr(self, "plan", None) is not None and self.plan.value:
                self.h.lib.makb200_qr_batched_plan_destroy(self.plan)
                self.plan = C.c_void_p(0)
        except Exception:
            pass


def qr_compact_batched_(As, QRs=None):
    """Batched ``qr_compact!`` over a list of blocks (new capability; per-block semantics)."""
    if len(As) == 0:
        return []
    return BatchedQRPlan(As, QRs).run()


# ---- L1 (LAPACK-shaped) shims: geqrf! / ungqr! / unmqr!(::B200, ...) ---------------------------------------
# Call shapes of YALAPACK / YACUSOLVER (yalapack.jl:168-200,550-583,688-735; yacusolver.jl:12-14), as
# ext/MatrixAlgebraKitCUDAExt/MatrixAlgebraKitCUDAExt.jl:32-34 forwards them; consumers: the blocksize = 1 branch of
# qr_householder! (qr.jl:160-175) and qr_null_householder! (qr.jl:236-262).
def geqrf_(A, tau=None):
    """``geqrf!(A, tau)``: A <- (V below the diagonal, R on and above it), tau[k], k = min(m, n)."""
    _check_matrix(A)
    m, n = A.shape
    k = min(m, n)
    if tau is None:
        tau = torch.empty(k, dtype=A.dtype, device=A.device)
    if tau.dim() != 1 or tau.shape[0] != k or tau.dtype != A.dtype:
        raise ValueError(f"tau: vector of length {k} and eltype {A.dtype} expected")
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    work = h.workspace(h.lib.makb200_geqrf_worksize(h.h, dt, m, n))
    rc = h.lib.makb200_geqrf(h.h, dt, m, n, _core.ptr(A), _core.ld(A), _core.ptr(tau), _core.ptr(work), work.numel())
    h.check(rc, "makb200_geqrf")
    return A, tau


def ungqr_(A, tau, Q=None):
    """``ungqr!``: Q (m x ncols, k <= ncols <= m) = first ncols columns of H_1 ... H_k; A, tau from ``geqrf_``."""
    _check_matrix(A)
    m, k = A.shape[0], tau.shape[0]
    if Q is None:
        Q = _core.colmajor_empty(m, k, A.dtype, A.device)
    _check_matrix(Q, "Q")
    if Q.shape[0] != m or not (k <= Q.shape[1] <= m) or Q.dtype != A.dtype:
        raise ValueError(f"Q: {m} x ncols with {k} <= ncols <= {m} expected, got {tuple(Q.shape)}")
    if Q.data_ptr() == A.data_ptr() and Q.numel() > 0:
        raise ValueError("ungqr_: in-place Q is not supported by the B200 driver")
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    work = h.workspace(h.lib.makb200_orgqr_worksize(h.h, dt, m, Q.shape[1], k))
    rc = h.lib.makb200_orgqr(h.h, dt, m, Q.shape[1], k, _core.ptr(A), _core.ld(A), _core.ptr(tau), _core.ptr(Q), _core.ld(Q),
                             _core.ptr(work), work.numel())
    h.check(rc, "makb200_orgqr")
    return Q


def unmqr_(side, trans, A, tau, C):
    """``unmqr!(side, trans, A, tau, C)``: C <- Q C ('L','N') or Q^H C ('L','C'); Q = H_1 ... H_k from ``geqrf_``.
    The right side is not provided by the B200 driver (ValueError, like the capability negatives of qr.jl:140-145)."""
    _check_matrix(A)
    _check_matrix(C, "C")
    if side != "L":
        raise ValueError("unmqr_: the B200 driver provides side = 'L' only")
    if trans not in ("N", "C") and not (trans == "T" and A.dtype == torch.float64):
        raise ValueError(f"unmqr_: trans = {trans!r}")
    m, n = C.shape
    k = tau.shape[0]
    if A.shape[0] != m or A.shape[1] < k or C.dtype != A.dtype:
        raise ValueError("unmqr_: A must be m x (>= k) with the rows and eltype of C")
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    work = h.workspace(h.lib.makb200_ormqr_worksize(h.h, dt, m, n, k))
    rc = h.lib.makb200_ormqr(h.h, dt, 0, _lib.OP_N if trans == "N" else _lib.OP_C, m, n, k, _core.ptr(A), _core.ld(A),
                             _core.ptr(tau), _core.ptr(C), _core.ld(C), _core.ptr(work), work.numel())
    h.check(rc, "makb200_ormqr")
    return C


def qr_null_householder_(A, N, positive=True, pivoted=False, blocksize=0):
    """``qr_null_householder!(driver, A, N)`` (qr.jl:236-262), the reference's recipe: N = [0; I], geqrf!, then
    unmqr!('L','N') - the m x (m-k) basis comes from k reflectors applied to m-k columns, no m x m Q is formed."""
    if blocksize > 1:
        raise ValueError("B200 does not provide a blocked QR decomposition")
    if pivoted:
        raise ValueError("B200 does not provide a pivoted QR decomposition")
    m, n = A.shape
    k = min(m, n)
    if tuple(N.shape) != (m, m - k):
        raise ValueError(f"N: size {tuple(N.shape)} != {(m, m - k)}")
    if m - k == 0:
        return N
    h = _core.Handle.get(A.device)
    # zero!(N); one!(view(N, k+1:m, 1:m-k)) in one launch: rectangular identity on the bottom block, zeros above
    N[:k].zero_()
    bottom = N[k:]
    rc = h.lib.makb200_tri_init(h.h, _core.dtype_code(N), 0, m - k, m - k, _core.ptr(bottom), _core.ld(N))
    h.check(rc, "makb200_tri_init")
    A, tau = geqrf_(A)
    return unmqr_("L", "N", A, tau, N)
