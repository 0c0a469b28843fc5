"""direct callers of the hot path.

* ``qr_null_``  — ``qr_null!`` (src/implementations/qr.jl:236-295): the last m-k columns of the full Q.
* ``lq_compact_/lq_full_/lq_null_`` through QR of the adjoint — ``lq_via_qr!`` /
  ``lq_null_via_qr!`` (src/implementations/lq.jl:130-131,303-327), which is how every GPU driver of
  the reference provides LQ.
* ``left_orth_/right_orth_/left_null_/right_null_`` — thin routers
  (src/implementations/orthnull.jl:79-117) with ``kind`` in {"qr"/"lq", "polar", "svd"}.
All numerical work stays in libmakb200 (QR, adjoint, SVD, polar, GEMM)."""
from . import _core
from .qr import copy_input, qr_compact_, qr_full_
from .svd import svd_compact_, svd_trunc_no_error_


def adjoint_(B, A):
    """B = A^H on the device (one tiled kernel)."""
    h = _core.Handle.get(A.device)
    m, n = A.shape
    if tuple(B.shape) != (n, m) or B.dtype != A.dtype or not _core.is_colmajor(B) or not _core.is_colmajor(A):
        raise ValueError("adjoint_: B must be the n x m column-major buffer for A^H")
    rc = h.lib.makb200_adjoint(h.h, _core.dtype_code(A), m, n, _core.ptr(A), _core.ld(A), _core.ptr(B), _core.ld(B))
    h.check(rc, "makb200_adjoint")
    return B


def _adj(A):
    return adjoint_(_core.colmajor_empty(A.shape[1], A.shape[0], A.dtype, A.device), A)


def qr_null_(A, N=None, alg=None, **kw):
    """``qr_null!(A, N, alg)``: orthonormal basis of the cokernel of A (m x (m - min(m,n))), by the reference's
    recipe ``qr_null_householder!`` (qr.jl:236-262): N = [0; I], ``geqrf!``, ``unmqr!('L','N')``.  Destroys A."""
    from .algorithms import select_algorithm
    from .qr import _alg_kwargs, qr_null_householder_
    m, n = A.shape
    k = min(m, n)
    if N is not None and tuple(N.shape) != (m, m - k):
        raise ValueError(f"N: size {tuple(N.shape)} != {(m, m - k)}")
    alg = select_algorithm("qr_null", A, alg, **kw)
    kws = _alg_kwargs(alg)
    if N is None:
        N = _core.colmajor_empty(m, m - k, A.dtype, A.device)
    return qr_null_householder_(A, N, positive=kws.get("positive", True), pivoted=kws.get("pivoted", False),
                                blocksize=kws.get("blocksize", 0))


def lq_compact_(A, LQ=None, alg=None, **kw):
    """``lq_compact!`` via ``lq_via_qr!``: A^H = Qt Rt  =>  A = Rt^H Qt^H."""
    m, n = A.shape
    k = min(m, n)
    Qt, Rt = qr_compact_(_adj(A), None, alg, **kw)
    L = LQ[0] if LQ is not None else _core.colmajor_empty(m, k, A.dtype, A.device)
    Q = LQ[1] if LQ is not None else _core.colmajor_empty(k, n, A.dtype, A.device)
    adjoint_(Q, Qt)
    if L is not None and L.numel() > 0:
        adjoint_(L, Rt)
    return L, Q


def lq_full_(A, LQ=None, alg=None, **kw):
    """``lq_full!`` via ``lq_via_qr!`` (Q n x n, L m x n)."""
    m, n = A.shape
    Qt, Rt = qr_full_(_adj(A), None, alg, **kw)
    L = LQ[0] if LQ is not None else _core.colmajor_empty(m, n, A.dtype, A.device)
    Q = LQ[1] if LQ is not None else _core.colmajor_empty(n, n, A.dtype, A.device)
    adjoint_(Q, Qt)
    if L is not None and L.numel() > 0:
        adjoint_(L, Rt)
    return L, Q


def lq_null_(A, Nh=None, alg=None, **kw):
    """``lq_null!`` via ``lq_null_via_qr!``: rows spanning the kernel of A ((n - min(m,n)) x n)."""
    Nt = qr_null_(_adj(A), None, alg, **kw)
    if Nh is None:
        Nh = _core.colmajor_empty(Nt.shape[1], Nt.shape[0], A.dtype, A.device)
    if Nt.numel() > 0:
        adjoint_(Nh, Nt)
    return Nh


def left_orth_(A, VC=None, kind="qr", trunc=None, **kw):
    """``left_orth!``: A = V C with V isometric (orthnull.jl:79-88)."""
    if kind == "qr":
        return qr_compact_(A, VC, **kw)
    if kind == "polar":
        from .polar import left_polar_
        return left_polar_(A, VC, **kw)
    if kind == "svd":
        if trunc is not None:
            U, S, Vh = svd_trunc_no_error_(A, None, None, trunc, **kw)
        else:
            U, S, Vh = svd_compact_(A, None, **kw)
        Cm = _core.colmajor_empty(Vh.shape[0], Vh.shape[1], Vh.dtype, Vh.device)
        Cm.copy_(Vh * S.to(Vh.dtype)[:, None])      # lmul!(S, C)
        return U, Cm
    raise ValueError(f"left_orth: unknown kind {kind}")


def right_orth_(A, CVh=None, kind="lq", trunc=None, **kw):
    """``right_orth!``: A = C Vh with Vh a co-isometry (orthnull.jl:90-99)."""
    if kind == "lq":
        return lq_compact_(A, CVh, **kw)
    if kind == "svd":
        if trunc is not None:
            U, S, Vh = svd_trunc_no_error_(A, None, None, trunc, **kw)
        else:
            U, S, Vh = svd_compact_(A, None, **kw)
        Cm = _core.colmajor_empty(U.shape[0], U.shape[1], U.dtype, U.device)
        Cm.copy_(U * S.to(U.dtype)[None, :])        # rmul!(C, S)
        return Cm, Vh
    raise ValueError(f"right_orth: unknown kind {kind}")


def left_null_(A, N=None, kind="qr", **kw):
    """``left_null!`` (orthnull.jl:103-104)."""
    if kind != "qr":
        raise ValueError("left_null: only the QR route is provided by the B200 driver")
    return qr_null_(A, N, **kw)


def right_null_(A, Nh=None, kind="lq", **kw):
    """``right_null!`` (orthnull.jl:112-113)."""
    if kind != "lq":
        raise ValueError("right_null: only the LQ route is provided by the B200 driver")
    return lq_null_(A, Nh, **kw)


def qr_null(A, **kw):
    return qr_null_(copy_input(A), None, **kw)


def lq_compact(A, **kw):
    return lq_compact_(copy_input(A), None, **kw)


def lq_full(A, **kw):
    return lq_full_(copy_input(A), None, **kw)


def lq_null(A, **kw):
    return lq_null_(copy_input(A), None, **kw)


def left_orth(A, **kw):
    return left_orth_(copy_input(A), None, **kw)


def right_orth(A, **kw):
    return right_orth_(copy_input(A), None, **kw)


def left_null(A, **kw):
    return left_null_(copy_input(A), None, **kw)


def right_null(A, **kw):
    return right_null_(copy_input(A), None, **kw)
