"""CPU oracle for the MatrixAlgebraKit.jl dense-factorization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``matrixalgebrakit.jl_b200/`` may import this
module: it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

What it is: a restatement of the reference's algorithm for this path.  The reference
(``/root/reference`` = MatrixAlgebraKit.jl v0.6.9) is Julia and cannot run here (no
``julia`` in the image).  Its arithmetic lives in a third-party dependency that is NOT in
the reference tree: LAPACK/BLAS behind ``libblastrampoline`` (``src/yalapack.jl:15-16``;
the LAPACK build is whatever the user's Julia ships, i.e. not pinned,
``Project.toml:6-7,38,48``).  We therefore replay **the same LAPACK routine sequence with
the same job flags** through ``scipy.linalg.lapack`` (OpenBLAS 0.3.31.dev, LP64) and
restate the Julia-side post-processing (R extraction, gauge fixing, truncation) in numpy.

Pinning: the reference's own tests contain no stored LAPACK output vectors ("parity
unpinned" at the vector level).  What the reference does pin — the doctest
KAT ``eigh_full([2 1 0;1 3 1;0 1 4]) -> [3-sqrt3, 3, 3+sqrt3]``
(``docs/src/user_interface/truncations.md:19-21``), the fixed-spectrum fixtures
``diag(0.9,0.3,0.1,0.01)`` (``test/testsuite/decompositions/svd.jl:198-254``,
``eigh.jl:128,167``), the truncation index-set literals (``test/common/truncate.jl``),
gauge conventions and the residual/orthogonality properties — is checked against this
oracle in ``tests/test_oracle.py``.

Every function cites the reference file:line it follows.  Matrices are numpy arrays in
Fortran (column-major) order, like Julia's.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack as _lp

EPS = np.finfo(np.float64).eps


def _pfx(a):
    return "z" if np.iscomplexobj(a) else "d"


def _f(a):
    dt = np.complex128 if np.iscomplexobj(a) else np.float64
    return np.asfortranarray(a, dtype=dt)


# --------------------------------------------------------------------------------------
# common/safemethods.jl:10-11, common/gauge.jl:12-77
# --------------------------------------------------------------------------------------
def sign_safe(x):
    """``sign_safe`` (src/common/safemethods.jl:10-11): +1 at zero, x/|x| otherwise."""
    x = np.asarray(x)
    if np.iscomplexobj(x):
        ax = np.abs(x)
        return np.where(ax == 0, 1.0 + 0j, x / np.where(ax == 0, 1.0, ax))
    return np.where(x < 0, -1.0, 1.0)


def _sign(x):
    """Julia ``sign`` (0 at zero)."""
    x = np.asarray(x)
    if np.iscomplexobj(x):
        ax = np.abs(x)
        return np.where(ax == 0, 0.0 + 0j, x / np.where(ax == 0, 1.0, ax))
    return np.sign(x)


def _argmaxabs_cols(U):
    """``_argmaxabs`` per column (src/common/gauge.jl:12-14): first entry of maximal
    modulus (strict ``<`` so the first maximum wins); abs for real, abs2 for complex."""
    if U.shape[0] == 0:
        return np.zeros(U.shape[1], dtype=U.dtype)
    mod = (U.real ** 2 + U.imag ** 2) if np.iscomplexobj(U) else np.abs(U)
    idx = np.argmax(mod, axis=0)  # numpy argmax returns the first maximum
    return U[idx, np.arange(U.shape[1])]


def gaugefix_qr(Q, R, Rd):
    """``gaugefix!(qr_householder!, Q, R, Rd)`` (src/common/gauge.jl:16-25)."""
    s = sign_safe(Rd)
    k = len(Rd)
    Q[:, :k] *= s[None, :]
    if R is not None:
        R[:k, :] *= np.conj(s)[:, None]
    return Q, R


def gaugefix_eigh(V):
    """``gaugefix!(eigh_full!, V)`` (src/common/gauge.jl:38-45)."""
    s = _sign(_argmaxabs_cols(V))
    V *= np.conj(s)[None, :]
    return V


def gaugefix_svd(U, Vh):
    """``gaugefix!(svd_compact!, U, Vh)`` (src/common/gauge.jl:69-77)."""
    s = _sign(_argmaxabs_cols(U))
    U *= np.conj(s)[None, :]
    Vh *= s[:, None]
    return U, Vh


# --------------------------------------------------------------------------------------
# QR: src/implementations/qr.jl:132-188 (LAPACK branch), yalapack.jl:88-89,303-333,882-928
# --------------------------------------------------------------------------------------
def qr_householder(A, mode="compact", positive=True, blocksize=0, compute_r=True):
    """``qr_householder!(LAPACK(), A, Q, R; positive, blocksize)``.

    blocksize=0 -> ``default_qr_blocksize = min(m,n,36)`` (yalapack.jl:88-89) and the
    ``geqrt`` + ``gemqrt('L','N',A,T,one!(Q))`` branch (qr.jl:155-163); blocksize=1 ->
    ``geqrf`` + ``unmqr`` on the identity (qr.jl:165-174).  Then
    ``R = uppertriangular!(A[axes(R)...])`` and the gauge (qr.jl:177-185)."""
    A = _f(A).copy(order="F")
    m, n = A.shape
    k = min(m, n)
    p = _pfx(A)
    nq = k if mode == "compact" else m
    Q = np.asfortranarray(np.eye(m, nq, dtype=A.dtype))
    if k == 0:
        R = np.zeros((nq, n), dtype=A.dtype, order="F")
        return Q, (R if compute_r else None)
    nb = blocksize if blocksize > 0 else min(k, 36)
    if nb > 1:
        nb = min(k, nb)
        V, T, info = getattr(_lp, p + "geqrt")(nb, A)
        assert info == 0
        Q, info = getattr(_lp, p + "gemqrt")(V[:, :k], T, Q, side="L", trans="N")
        assert info == 0
    else:
        V, tau, _, info = getattr(_lp, p + "geqrf")(A)
        assert info == 0
        mq = getattr(_lp, "dormqr" if p == "d" else "zunmqr")
        lw = max(1, 64 * nq)
        Q, _, info = mq("L", "N", V[:, :k], tau, Q, lw)
        assert info == 0
    Rd = np.diagonal(V)[:k].copy()
    R = None
    if compute_r:
        R = np.asfortranarray(np.triu(V[:nq, :]))
        if nq > V.shape[0]:
            raise AssertionError
    if positive:
        gaugefix_qr(Q, R, Rd)
    return Q, R


def qr_compact(A, **kw):
    """``qr_compact!`` (src/implementations/qr.jl:110-113)."""
    return qr_householder(A, mode="compact", **kw)


def qr_full(A, **kw):
    """``qr_full!`` (src/implementations/qr.jl:106-109)."""
    return qr_householder(A, mode="full", **kw)


# --------------------------------------------------------------------------------------
# SVD: src/implementations/svd.jl:196-237, yalapack.jl:2066-2202
# --------------------------------------------------------------------------------------
def svd_compact(A, alg="DivideAndConquer", fixgauge=True):
    """``svd_compact!`` with ``DivideAndConquer`` (gesdd, jobz='S', svd.jl:196-201),
    ``SafeDivideAndConquer`` (gesdvd!: gesdd, on info>0 gesvd, yalapack.jl:2173-2202) or
    ``QRIteration`` (gesvd)."""
    A = _f(A)
    m, n = A.shape
    k = min(m, n)
    if A.size == 0:  # svd.jl:197 -> one!(U), zero!(S), one!(Vh)
        return (np.asfortranarray(np.eye(m, k, dtype=A.dtype)), np.zeros(k),
                np.asfortranarray(np.eye(k, n, dtype=A.dtype)))
    p = _pfx(A)
    if alg in ("DivideAndConquer", "SafeDivideAndConquer"):
        U, S, Vh, info = getattr(_lp, p + "gesdd")(A, compute_uv=1, full_matrices=0)
        if info > 0 and alg == "SafeDivideAndConquer":
            U, S, Vh, info = getattr(_lp, p + "gesvd")(A, compute_uv=1, full_matrices=0)
    elif alg == "QRIteration":
        U, S, Vh, info = getattr(_lp, p + "gesvd")(A, compute_uv=1, full_matrices=0)
    else:
        raise ValueError(alg)
    assert info == 0, info
    U = np.asfortranarray(U)
    Vh = np.asfortranarray(Vh)
    if fixgauge:
        gaugefix_svd(U, Vh)
    return U, S, Vh


def gaugefix_svd_full(U, Vh):
    """``gaugefix!(svd_full!, U, Vh)`` (src/common/gauge.jl:47-67): leading min(m,n) triplets as in the compact
    gauge; extra columns of U and extra rows of Vh are each scaled by conj(sign(own entry of maximal modulus))."""
    m, n = U.shape[1], Vh.shape[0]
    k = min(m, n)
    s = _sign(_argmaxabs_cols(U[:, :k]))
    U[:, :k] *= np.conj(s)[None, :]
    Vh[:k, :] *= s[:, None]
    if m > k:
        s2 = _sign(_argmaxabs_cols(U[:, k:]))
        U[:, k:] *= np.conj(s2)[None, :]
    if n > k:
        s3 = _sign(_argmaxabs_cols(Vh[k:, :].T))
        Vh[k:, :] *= np.conj(s3)[:, None]
    return U, Vh


def svd_full(A, alg="DivideAndConquer", fixgauge=True):
    """``svd_full!`` (svd.jl:202-212): gesdd with jobz='A'; S is the m x n matrix with the values on its diagonal."""
    A = _f(A)
    m, n = A.shape
    k = min(m, n)
    if A.size == 0:
        return (np.asfortranarray(np.eye(m, dtype=A.dtype)), np.zeros((m, n)), np.asfortranarray(np.eye(n, dtype=A.dtype)))
    p = _pfx(A)
    drv = "gesdd" if alg in ("DivideAndConquer", "SafeDivideAndConquer") else "gesvd"
    U, sv, Vh, info = getattr(_lp, p + drv)(A, compute_uv=1, full_matrices=1)
    assert info == 0, info
    U, Vh = np.asfortranarray(U), np.asfortranarray(Vh)
    S = np.zeros((m, n))
    S[np.arange(k), np.arange(k)] = sv
    if fixgauge:
        gaugefix_svd_full(U, Vh)
    return U, S, Vh


def svd_vals(A):
    """``svd_vals!`` (svd.jl:214-219): gesdd with jobz='N'."""
    A = _f(A)
    if A.size == 0:
        return np.zeros(min(A.shape))
    _, S, _, info = getattr(_lp, _pfx(A) + "gesdd")(A, compute_uv=0)
    assert info == 0
    return S


# -- truncation strategies: src/interface/truncation.jl:96-275, implementations/truncation.jl
def truncrank(howmany):
    return ("rank", int(howmany))


def trunctol(atol=0.0, rtol=0.0, p=2):
    return ("tol", float(atol), float(rtol), p)


def truncerror(atol=0.0, rtol=0.0, p=2):
    return ("error", float(atol), float(rtol), p)


def notrunc():
    return ("none",)


def trunc_and(*components):
    """``TruncationIntersection`` (``&``; implementations/truncation.jl:104-120)."""
    return ("and",) + tuple(components)


def trunc_or(*components):
    """``TruncationUnion`` (``|``; implementations/truncation.jl:140-155)."""
    return ("or",) + tuple(components)


def truncation_strategy(atol=None, rtol=None, maxrank=None, minrank=None, maxerror=None):
    """``TruncationStrategy(; atol, rtol, maxrank, minrank, maxerror)`` keyword form
    (src/interface/truncation.jl:37-66): (tol & maxrank & maxerror) | minrank."""
    comps = []
    if atol is not None or rtol is not None:
        comps.append(trunctol(atol or 0.0, rtol or 0.0))
    if maxrank is not None:
        comps.append(truncrank(maxrank))
    if maxerror is not None:
        comps.append(truncerror(atol=maxerror))
    s = notrunc() if not comps else (comps[0] if len(comps) == 1 else trunc_and(*comps))
    if minrank is not None:
        s = truncrank(minrank) if not comps else trunc_or(s, truncrank(minrank))
    return s


def _combine(values, strategy, finder):
    sets = [set(int(i) for i in finder(values, c)) for c in strategy[1:]]
    if not sets:
        return np.arange(len(values)) if strategy[0] == "and" else np.arange(0)
    out = set.intersection(*sets) if strategy[0] == "and" else set.union(*sets)
    return np.array(sorted(out), dtype=np.int64)


def _pnorm(v, p):
    v = np.abs(np.asarray(v, dtype=np.float64))
    if p == 2:
        return float(np.sqrt(np.sum(v * v)))
    if np.isinf(p):
        return float(v.max()) if v.size else 0.0
    return float(np.sum(v ** p) ** (1.0 / p))


def findtruncated_svd(values, strategy):
    """``findtruncated_svd`` on descending non-negative values; returns 0-based indices.
    rank: implementations/truncation.jl:54-58; tol: :69-79 (searchsortedlast, keep
    ``sigma >= max(atol, rtol*norm(values,p))``); error: :86-102 (_truncerr_impl)."""
    values = np.asarray(values, dtype=np.float64)
    n = len(values)
    kind = strategy[0]
    if kind in ("and", "or"):
        return _combine(values, strategy, findtruncated_svd)
    if kind == "none":
        return np.arange(n)
    if kind == "rank":
        return np.arange(min(strategy[1], n))
    if kind == "tol":
        _, atol, rtol, p = strategy
        thr = max(atol, rtol * _pnorm(values, p))
        # searchsortedlast(values, thr; by=abs, rev=true): last i with values[i] >= thr
        i = int(np.searchsorted(-values, -thr, side="right"))
        return np.arange(i)
    if kind == "error":
        _, atol, rtol, p = strategy
        vp = np.abs(values) ** p
        Np = float(vp.sum())
        ep = max(atol ** p, rtol ** p * Np)
        if ep >= Np:
            return np.arange(0)
        cs = np.cumsum(vp[::-1])
        first = int(np.argmax(cs >= ep))  # findfirst(>=(ep)) 0-based
        rank = n - first
        return np.arange(rank)
    raise ValueError(strategy)


def truncation_error(values, ind):
    """``truncation_error!`` (implementations/truncation.jl:168-174): zero the kept
    entries, 2-norm of the rest."""
    v = np.array(values, dtype=np.float64, copy=True)
    v[ind] = 0.0
    return float(np.linalg.norm(v))


def svd_trunc(A, trunc, alg="DivideAndConquer"):
    """``svd_trunc!`` (svd.jl:232-237): full compact SVD, slice, eps = norm(discarded)."""
    U, S, Vh = svd_compact(A, alg=alg)
    ind = findtruncated_svd(S, trunc)
    eps_ = truncation_error(S, ind)
    return (np.asfortranarray(U[:, ind]), S[ind].copy(), np.asfortranarray(Vh[ind, :]), eps_)


# --------------------------------------------------------------------------------------
# eigh: src/implementations/eigh.jl:11-18,150-169, yalapack.jl:1164-1362
# --------------------------------------------------------------------------------------
def default_hermitian_tol(A):
    """``default_hermitian_tol`` (src/common/defaults.jl:44): eps(norm(A, Inf))^(3/4)
    where Julia's ``norm(A, Inf)`` of a matrix is the largest |entry|."""
    nrm = float(np.max(np.abs(A))) if A.size else 0.0
    return float(np.spacing(nrm) ** 0.75)


def is_hermitian(A, atol=None):
    """``ishermitian(A; atol)`` approx. branch (src/common/matrixproperties.jl:150-172):
    ``norm(project_antihermitian(A)) <= atol``."""
    A = np.asarray(A)
    if A.shape[0] != A.shape[1]:
        return False
    atol = default_hermitian_tol(A) if atol is None else atol
    return float(np.linalg.norm((A - A.conj().T) / 2)) <= atol


def eigh_full(A, alg="RobustRepresentations", fixgauge=True, check=True):
    """``eigh_full!`` (eigh.jl:123-127,150-156): Hermitian check (DomainError ->
    ValueError here), ``heevr!(A, jobz='V', range='A', uplo='U', abstol=-1)`` for
    RobustRepresentations (yalapack.jl:1164-1279) or ``heevd!`` for DivideAndConquer
    (:1280-1362), then the eigenvector gauge."""
    A = _f(A)
    if check and not is_hermitian(A):
        raise ValueError("Hermitian matrix was expected")
    n = A.shape[0]
    if n == 0:
        return np.zeros(0), np.zeros((0, 0), dtype=A.dtype, order="F")
    p = _pfx(A)
    if alg == "RobustRepresentations":
        fn = getattr(_lp, "dsyevr" if p == "d" else "zheevr")
        out = fn(A, compute_v=1, range="A", lower=0, abstol=-1.0)
        w, V, info = out[0], out[1], out[-1]
    elif alg == "DivideAndConquer":
        fn = getattr(_lp, "dsyevd" if p == "d" else "zheevd")
        w, V, info = fn(A, compute_v=1, lower=0)
    else:
        raise ValueError(alg)
    assert info == 0, info
    V = np.asfortranarray(V)
    if fixgauge:
        gaugefix_eigh(V)
    return w, V


def eigh_vals(A):
    """``eigh_vals!`` (eigh.jl:157-161)."""
    A = _f(A)
    fn = getattr(_lp, "dsyevr" if _pfx(A) == "d" else "zheevr")
    out = fn(A, compute_v=0, range="A", lower=0, abstol=-1.0)
    return out[0]


def findtruncated(values, strategy):
    """generic ``findtruncated`` (implementations/truncation.jl:48-83) as used by
    ``eigh_trunc!``: rank = sortperm by abs descending, first howmany."""
    values = np.asarray(values, dtype=np.float64)
    kind = strategy[0]
    if kind in ("and", "or"):
        return _combine(values, strategy, findtruncated)
    if kind == "none":
        return np.arange(len(values))
    if kind == "rank":
        order = np.argsort(-np.abs(values), kind="stable")
        return order[: min(strategy[1], len(values))]
    if kind == "tol":
        _, atol, rtol, p = strategy
        thr = max(atol, rtol * _pnorm(values, p))
        return np.nonzero(np.abs(values) >= thr)[0]
    if kind == "error":
        _, atol, rtol, p = strategy
        order = np.argsort(-np.abs(values), kind="stable")
        sub = findtruncated_svd(np.abs(values)[order], strategy)
        return order[sub]
    raise ValueError(strategy)


def eigh_trunc(A, trunc, alg="RobustRepresentations"):
    """``eigh_trunc!`` (eigh.jl:165-169)."""
    w, V = eigh_full(A, alg=alg)
    ind = findtruncated(w, trunc)
    return w[ind].copy(), np.asfortranarray(V[:, ind]), truncation_error(w, ind)


# --------------------------------------------------------------------------------------
# polar: src/implementations/polar.jl:59-70 (PolarViaSVD), :99-166 (PolarNewton)
# --------------------------------------------------------------------------------------
def left_polar(A, compute_p=True, alg="PolarViaSVD"):
    """``left_polar!`` default ``PolarViaSVD(SafeDivideAndConquer)``: W = U*Vh,
    P = (sqrt(S) Vh)^H (sqrt(S) Vh) (polar.jl:59-70,87-95)."""
    A = _f(A)
    m, n = A.shape
    if m < n:
        raise ValueError("`left_polar!` requires m >= n")  # polar.jl:9-10
    if alg == "PolarNewton":
        return _left_polar_newton(A, compute_p)
    U, S, Vh = svd_compact(A, alg="SafeDivideAndConquer")
    W = np.asfortranarray(U @ Vh)
    P = None
    if compute_p:
        B = np.sqrt(S)[:, None] * Vh
        P = np.asfortranarray(B.conj().T @ B)
    return W, P


def _left_polar_newton(A, compute_p, tol=None, maxiter=10):
    """``left_polar_newton!`` (polar.jl:128-166): scaled Newton on R (QR-reduced if m>n)."""
    m, n = A.shape
    tol = EPS ** (2.0 / 3.0) if tol is None else tol
    A0 = A.copy()
    Q = None
    R = A.copy()
    if m > n:
        Q, R = qr_compact(A)
    R = np.array(R)
    for i in range(maxiter):
        Rinvh = np.linalg.inv(R).conj().T
        g = np.sqrt(np.linalg.norm(Rinvh) / np.linalg.norm(R))
        Rn = (g * R + Rinvh / g) / 2
        conv = float(np.max(np.abs(Rinvh / g - Rn)))
        R = Rn
        if conv <= tol:
            break
    W = np.asfortranarray(Q @ R if Q is not None else R)
    P = None
    if compute_p:
        P = W.conj().T @ A0
        P = np.asfortranarray((P + P.conj().T) / 2)
    return W, P


# --------------------------------------------------------------------------------------
# synthetic inputs: i.i.d. N(0,1), numpy PCG64(seed); complex = (x+iy)/sqrt2
# --------------------------------------------------------------------------------------
def randn_matrix(m, n, dtype="f64", seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    if dtype in ("f64", np.float64):
        return np.asfortranarray(rng.standard_normal((n, m)).T)
    x = rng.standard_normal((n, m)).T
    y = rng.standard_normal((n, m)).T
    return np.asfortranarray((x + 1j * y) / np.sqrt(2.0))


def rand_hermitian(n, dtype="f64", seed=0):
    """``project_hermitian!(randn)`` as in test/testsuite/decompositions/eigh.jl:29."""
    G = randn_matrix(n, n, dtype, seed)
    return np.asfortranarray((G + G.conj().T) / 2)


# --------------------------------------------------------------------------------------
# parity metrics
# --------------------------------------------------------------------------------------
def rel_resid(A, *factors):
    P = factors[0]
    for F in factors[1:]:
        P = P @ F
    return float(np.linalg.norm(A - P) / max(np.linalg.norm(A), np.finfo(float).tiny))


def orth_err(Q, side="left"):
    if side == "left":
        G = Q.conj().T @ Q
    else:
        G = Q @ Q.conj().T
    return float(np.linalg.norm(G - np.eye(G.shape[0])))


def tol_for(m, n=None):
    """north_star tolerance 10*n*eps with n = max(m, n)."""
    return 10.0 * max(m, n if n is not None else m, 1) * EPS
