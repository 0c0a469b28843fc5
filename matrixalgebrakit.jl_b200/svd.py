"""svd_compact / svd_vals / svd_trunc on B200 — mirrors src/implementations/svd.jl:
``check_input`` (:23-35), ``initialize_output`` (:81-88), ``svd_compact_<alg>!(driver, A, U, S, Vh;
fixgauge)`` (:196-201), ``svd_vals`` (:214-219), ``svd_trunc!`` / ``svd_trunc_no_error!`` (:226-237).
S is held as the vector of the ``Diagonal`` (``diagview(S)``)."""
import ctypes as C

import torch

from . import _core
from .algorithms import Algorithm, TruncatedAlgorithm, resolve_driver, select_algorithm
from .truncation import TruncationByOrder, findtruncated_svd, select_truncation, truncation_error_


def initialize_output(A):
    m, n = A.shape
    k = min(m, n)
    return (_core.colmajor_empty(m, k, A.dtype, A.device), torch.empty(k, dtype=torch.float64, device=A.device),
            _core.colmajor_empty(k, n, A.dtype, A.device))


def check_input(A, USVh):
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    _core.dtype_code(A)
    m, n = A.shape
    k = min(m, n)
    U, S, Vh = USVh
    if tuple(U.shape) != (m, k) or U.dtype != A.dtype or not _core.is_colmajor(U):
        raise ValueError(f"U: {m} x {k} column-major matrix expected")
    if S.dim() != 1 or S.shape[0] != k or S.dtype != torch.float64:
        raise ValueError("S: real vector (diagview of the Diagonal) of length min(m,n) expected")
    if tuple(Vh.shape) != (k, n) or Vh.dtype != A.dtype or not _core.is_colmajor(Vh):
        raise ValueError(f"Vh: {k} x {n} column-major matrix expected")


def _alg_ok(alg):
    # gesdd!/gesdvd!/gesvdp!(::B200) are one implementation (QDWH polar + Hermitian D&C): the tag names the contract
    # (descending values, gauge-fixed vectors), the driver the implementation, so the BASELINE-named alias
    # LAPACK_DivideAndConquer(driver = B200()) works.  gesvd!/gesvdj! (QRIteration, Jacobi) are not provided and throw.
    if not isinstance(alg, Algorithm) or alg.name not in ("SVDViaPolar", "DivideAndConquer", "SafeDivideAndConquer"):
        raise ValueError(f"svd: algorithm {alg} is not provided by the B200 driver "
                         "(SVDViaPolar, DivideAndConquer and SafeDivideAndConquer are)")
    resolve_driver(alg.get("driver"), None)


def _gesvdp_(A, S, U, Vh, fixgauge):
    h = _core.Handle.get(A.device)
    m, n = A.shape
    dt = _core.dtype_code(A)
    lw = h.lib.makb200_svd_worksize(h.h, dt, m, n)
    work = h.workspace(lw)
    vec = U is not None
    rc = h.lib.makb200_svd(h.h, dt, int(bool(fixgauge)), m, n, _core.ptr(A), _core.ld(A), _core.ptr(S),
                           _core.ptr(U) if vec else C.c_void_p(0), _core.ld(U) if vec else 0,
                           _core.ptr(Vh) if vec else C.c_void_p(0), _core.ld(Vh) if vec else 0, 0.0,
                           _core.ptr(work), work.numel(), C.c_void_p(0))
    h.check(rc, "makb200_svd")


def svd_compact_(A, USVh=None, alg=None, **kw):
    """``svd_compact!(A, (U,S,Vh), alg)`` (svd.jl:169-172,196-201). Destroys A."""
    alg = select_algorithm("svd_compact", A, alg, **kw)
    _alg_ok(alg)
    if USVh is None:
        USVh = initialize_output(A)
    check_input(A, USVh)
    U, S, Vh = USVh
    if A.numel() == 0:  # svd.jl:197: one!(U), zero!(S), one!(Vh)
        U.zero_(); Vh.zero_(); S.zero_()
        if U.numel():
            U.diagonal().fill_(1)
        if Vh.numel():
            Vh.diagonal().fill_(1)
        return U, S, Vh
    _gesvdp_(A, S, U, Vh, alg.get("fixgauge", True))
    return U, S, Vh


def svd_vals_(A, S=None, alg=None, **kw):
    """``svd_vals!`` (svd.jl:214-219): job 'N'."""
    alg = select_algorithm("svd_vals", A, alg, **kw)
    _alg_ok(alg)
    k = min(A.shape)
    if S is None:
        S = torch.empty(k, dtype=torch.float64, device=A.device)
    if A.numel() == 0:
        return S.zero_()
    _core.dtype_code(A)
    _gesvdp_(A, S, None, None, False)
    return S


def _copy_input(A):
    from .qr import copy_input
    return copy_input(A)


def _gauge_columns_(V):
    """columns of V *= conj(sign(first entry of maximal modulus)) (common/gauge.jl:12-14,38-45), one launch."""
    h = _core.Handle.get(V.device)
    rc = h.lib.makb200_gauge_columns(h.h, _core.dtype_code(V), V.shape[0], V.shape[1], _core.ptr(V), _core.ld(V))
    h.check(rc, "makb200_gauge_columns")
    return V


def initialize_output_full(A):
    """``initialize_output(svd_full!, A, alg)`` (svd.jl:74-80): U m x m, S m x n REAL matrix, Vh n x n."""
    m, n = A.shape
    return (_core.colmajor_empty(m, m, A.dtype, A.device), _core.colmajor_zeros(m, n, torch.float64, A.device),
            _core.colmajor_empty(n, n, A.dtype, A.device))


def svd_full_(A, USVh=None, alg=None, **kw):
    """``svd_full!(A, (U,S,Vh), alg)`` (svd.jl:173-176,202-212; LAPACK job 'A').  The leading min(m,n) triplets are
    the compact decomposition; the extra columns of U (m > n) / rows of Vh (m < n) are an orthonormal basis of the
    complement, taken the way ``qr_null!`` takes it (k Householder reflectors applied to [0; I], no m x m QR), each
    with its own gauge (common/gauge.jl:47-67).  ``supports_svd_full(::B200, :svd_polar) = true``.  Destroys A."""
    from .orthnull import adjoint_
    from .qr import qr_null_householder_
    alg = select_algorithm("svd_compact", A, alg, **kw)
    _alg_ok(alg)
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    m, n = A.shape
    k = min(m, n)
    if USVh is None:
        USVh = initialize_output_full(A)
    U, S, Vh = USVh
    if tuple(U.shape) != (m, m) or U.dtype != A.dtype or not _core.is_colmajor(U):
        raise ValueError(f"U: {m} x {m} column-major matrix expected")
    if tuple(S.shape) != (m, n) or S.dtype != torch.float64:
        raise ValueError(f"S: real {m} x {n} matrix expected")
    if tuple(Vh.shape) != (n, n) or Vh.dtype != A.dtype or not _core.is_colmajor(Vh):
        raise ValueError(f"Vh: {n} x {n} column-major matrix expected")
    fixgauge = alg.get("fixgauge", True)
    S.zero_()
    if A.numel() == 0:   # svd.jl:205: one!(U), zero!(S), one!(Vh)
        for M in (U, Vh):
            if M.numel():
                M.zero_()
                M.diagonal().fill_(1)
        return U, S, Vh
    Sd = torch.empty(k, dtype=torch.float64, device=A.device)
    Uc = U[:, :k]                                       # leading columns of the caller's U (ld = m)
    Vc = Vh if m >= n else _core.colmajor_empty(k, n, A.dtype, A.device)   # k x n block: rows of Vh are strided
    _gesvdp_(A, Sd, Uc, Vc, fixgauge)
    if m < n:
        Vh[:k].copy_(Vc)
    torch.diagonal(S).copy_(Sd)
    if m > k:
        N = qr_null_householder_(_clone_cm(Uc), U[:, k:])
        if fixgauge:
            _gauge_columns_(N)
    if n > k:
        # rows spanning the complement of the row space: (qr_null of Vc^H)^H; column gauge before the adjoint makes the
        # entry of maximal modulus of every extra row real positive, as the reference's row rule does
        Vct = adjoint_(_core.colmajor_empty(n, k, A.dtype, A.device), Vc)
        Nt = qr_null_householder_(Vct, _core.colmajor_empty(n, n - k, A.dtype, A.device))
        if fixgauge:
            _gauge_columns_(Nt)
        Nh = adjoint_(_core.colmajor_empty(n - k, n, A.dtype, A.device), Nt)
        Vh[k:].copy_(Nh)
    return U, S, Vh


def _clone_cm(M):
    out = _core.colmajor_empty(M.shape[0], M.shape[1], M.dtype, M.device)
    out.copy_(M)
    return out


def svd_full(A, alg=None, **kw):
    return svd_full_(_copy_input(A), None, alg, **kw)




def svd_compact(A, alg=None, **kw):
    return svd_compact_(_copy_input(A), None, alg, **kw)


def svd_vals(A, alg=None, **kw):
    return svd_vals_(_copy_input(A), None, alg, **kw)


def _select_trunc_alg(A, alg, trunc, kw):
    """``select_algorithm(svd_trunc!, A, alg; trunc, kw...)`` (interface/svd.jl:181-192)."""
    if isinstance(alg, TruncatedAlgorithm):
        if trunc is not None:
            raise ValueError("`trunc` can't be specified when `alg` is a `TruncatedAlgorithm`")
        return alg
    return TruncatedAlgorithm(select_algorithm("svd_compact", A, alg, **kw), select_truncation(trunc))


def _truncate(U, S, Vh, strategy):
    ind = findtruncated_svd(S, strategy)
    Ut = _core.colmajor_empty(U.shape[0], len(ind), U.dtype, U.device)
    Ut.copy_(U[:, ind])
    Vt = _core.colmajor_empty(len(ind), Vh.shape[1], Vh.dtype, Vh.device)
    Vt.copy_(Vh[ind, :])
    return (Ut, S[ind].clone(), Vt), ind


def _leading_rank(A, USVh, talg):
    """Rank r when the kept set is known before the decomposition — ``truncrank(r)`` keeps the r leading
    triplets of the sorted spectrum (truncation.jl:54-58) — and the caller did not hand in full-size
    outputs; else None (values-dependent strategies need the full decomposition first)."""
    s = talg.trunc
    k = min(A.shape)
    if USVh is not None or not isinstance(s, TruncationByOrder) or not s.rev or A.numel() == 0:
        return None
    return s.howmany if 0 < s.howmany < k else None


def _svd_leading_(A, r, alg):
    """Leading-r SVD in one C call (``makb200_svd_leading``): all k singular values, vectors of the r
    leading triplets only (partial back-transformation, m x r x n product for U)."""
    _alg_ok(alg)
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    h = _core.Handle.get(A.device)
    m, n = A.shape
    dt = _core.dtype_code(A)
    S = torch.empty(min(m, n), dtype=torch.float64, device=A.device)
    U = _core.colmajor_empty(m, r, A.dtype, A.device)
    Vh = _core.colmajor_empty(r, n, A.dtype, A.device)
    work = h.workspace(h.lib.makb200_svd_worksize(h.h, dt, m, n))
    rc = h.lib.makb200_svd_leading(h.h, dt, int(bool(alg.get("fixgauge", True))), m, n, r, _core.ptr(A), _core.ld(A),
                                   _core.ptr(S), _core.ptr(U), _core.ld(U), _core.ptr(Vh), _core.ld(Vh), 0.0,
                                   _core.ptr(work), work.numel(), C.c_void_p(0))
    h.check(rc, "makb200_svd_leading")
    return U, S, Vh


def svd_trunc_no_error_(A, USVh=None, alg=None, trunc=None, **kw):
    """``svd_trunc_no_error!`` (svd.jl:226-230): no device->host read of the error."""
    talg = _select_trunc_alg(A, alg, trunc, kw)
    r = _leading_rank(A, USVh, talg)
    if r is not None:
        U, S, Vh = _svd_leading_(A, r, talg.alg)
        return U, S[:r].clone(), Vh
    U, S, Vh = svd_compact_(A, USVh, talg.alg)
    out, _ = _truncate(U, S, Vh, talg.trunc)
    return out


def svd_trunc_(A, USVh=None, alg=None, trunc=None, **kw):
    """``svd_trunc!`` (svd.jl:232-237): full compact SVD, slice, eps = norm of the discarded
    values; clobbers the untruncated S (truncation.jl:171-174).  With ``truncrank(r)`` and no
    caller-provided outputs only the r leading triplets' vectors are formed (same values and error)."""
    talg = _select_trunc_alg(A, alg, trunc, kw)
    r = _leading_rank(A, USVh, talg)
    if r is not None:
        U, S, Vh = _svd_leading_(A, r, talg.alg)
        St = S[:r].clone()
        eps_ = truncation_error_(S, torch.arange(r, device=S.device))
        return U, St, Vh, eps_
    U, S, Vh = svd_compact_(A, USVh, talg.alg)
    out, ind = _truncate(U, S, Vh, talg.trunc)
    eps_ = truncation_error_(S, ind)
    return out + (eps_,)


def svd_trunc(A, alg=None, trunc=None, **kw):
    return svd_trunc_(_copy_input(A), None, alg, trunc, **kw)


def svd_trunc_no_error(A, alg=None, trunc=None, **kw):
    return svd_trunc_no_error_(_copy_input(A), None, alg, trunc, **kw)


class BatchedSVDPlan:
    """Argument arrays of a batched ``svd_compact!`` built once; ``run()`` is one C-ABI call."""

    def __init__(self, As, USVhs=None, fixgauge=True):
        self.As = As
        self.h = _core.Handle.get(As[0].device)
        self.dt = _core.dtype_code(As[0])
        self.outs = USVhs if USVhs is not None else [initialize_output(A) for A in As]
        for A, o in zip(As, self.outs):
            check_input(A, o)
        b = self.b = len(As)
        self.fixgauge = int(bool(fixgauge))
        IA, VP = C.c_int * b, C.c_void_p * b
        self.m = IA(*[A.shape[0] for A in As])
        self.n = IA(*[A.shape[1] for A in As])
        self.lda = IA(*[_core.ld(A) for A in As])
        self.ldu = IA(*[_core.ld(o[0]) for o in self.outs])
        self.ldv = IA(*[_core.ld(o[2]) for o in self.outs])
        self.Ap = VP(*[A.data_ptr() for A in As])
        self.Sp = VP(*[o[1].data_ptr() for o in self.outs])
        self.Up = VP(*[o[0].data_ptr() for o in self.outs])
        self.Vp = VP(*[o[2].data_ptr() for o in self.outs])
        self.lw = self.h.lib.makb200_svd_batched_worksize(self.h.h, self.dt, b, self.m, self.n)

    def run(self):
        h = _core.Handle.get(self.As[0].device)
        work = h.workspace(self.lw)
        rc = h.lib.makb200_svd_batched(h.h, self.dt, self.fixgauge, self.b, self.m, self.n, self.Ap, self.lda,
                                       self.Sp, self.Up, self.ldu, self.Vp, self.ldv, C.c_void_p(0),
                                       _core.ptr(work), work.numel())
        h.check(rc, "makb200_svd_batched")
        return self.outs


def svd_compact_batched_(As, USVhs=None, fixgauge=True):
    """Batched ``svd_compact!`` over a list of blocks (new capability; per-block semantics)."""
    if len(As) == 0:
        return []
    return BatchedSVDPlan(As, USVhs, fixgauge).run()


def svd_vals_batched_(As, Ss=None):
    """Batched ``svd_vals!`` (svd.jl:214-219 per block, job 'N'): one C-ABI call, U = Vh = NULL."""
    if len(As) == 0:
        return []
    h = _core.Handle.get(As[0].device)
    dt = _core.dtype_code(As[0])
    b = len(As)
    for A in As:
        if not _core.is_colmajor(A) or _core.dtype_code(A) != dt:
            raise ValueError("svd_vals_batched_: column-major blocks of one eltype expected")
    if Ss is None:
        Ss = [torch.empty(min(A.shape), dtype=torch.float64, device=A.device) for A in As]
    for A, S in zip(As, Ss):
        if S.dim() != 1 or S.shape[0] != min(A.shape) or S.dtype != torch.float64:
            raise ValueError("S: real vector of length min(m,n) expected")
    IA, VP = C.c_int * b, C.c_void_p * b
    m, n = IA(*[A.shape[0] for A in As]), IA(*[A.shape[1] for A in As])
    lda = IA(*[_core.ld(A) for A in As])
    Ap, Sp = VP(*[A.data_ptr() for A in As]), VP(*[S.data_ptr() for S in Ss])
    work = h.workspace(h.lib.makb200_svd_batched_worksize(h.h, dt, b, m, n))
    rc = h.lib.makb200_svd_batched(h.h, dt, 0, b, m, n, Ap, lda, Sp, None, None, None, None, C.c_void_p(0),
                                   _core.ptr(work), work.numel())
    h.check(rc, "makb200_svd_batched")
    return Ss


def svd_trunc_batched_(As, trunc, USVhs=None, maxranks=None):
    """Batched ``svd_trunc!``: batched compact SVD, then per block the reference's slice and error
    (svd.jl:232-237) -> list of ``(U, S, Vh, eps)``.  Strategies that keep a prefix of the sorted values
    (everything ``TruncationStrategy(; atol, rtol, maxrank, maxerror, minrank)`` builds) are decided for
    ALL blocks by one kernel and one device->host read, and the kept factors are views ``U[:, :r]``,
    ``S[:r]``, ``Vh[:r, :]`` of the compact ones; other strategies take the per-block host search.
    ``maxranks``: optional per-block rank caps (BASELINE config 3 uses ``truncrank(n_i // 2)`` per block)."""
    from .truncation import device_spec, trunc_and, trunc_select_batched_, truncrank
    strategy = select_truncation(trunc)
    outs = svd_compact_batched_(As, USVhs)
    spec = device_spec(strategy)
    res = []
    if spec is not None:
        ranks, eps = trunc_select_batched_([S for _, S, _ in outs], spec, maxranks)
        for (U, S, Vh), r, e in zip(outs, ranks, eps):
            res.append((U[:, :r], S[:r], Vh[:r, :], e))
        return res
    for i, (U, S, Vh) in enumerate(outs):
        st = strategy if maxranks is None else trunc_and(strategy, truncrank(int(maxranks[i])))
        (Ut, St, Vt), ind = _truncate(U, S, Vh, st)
        res.append((Ut, St, Vt, truncation_error_(S, ind)))
    return res
