"""left_polar on B200 — mirrors src/implementations/polar.jl: ``check_input`` (:6-17),
``initialize_output`` (:33-38), ``left_polar!(A, (W,P), alg)`` (:59-70,99-111).  The B200
algorithm is ``B200_QDWH`` (new); ``PolarViaSVD`` is provided on top of the B200 SVD."""
import ctypes as C

import torch

from . import _core
from .algorithms import Algorithm, select_algorithm


def initialize_output(A):
    m, n = A.shape
    return (_core.colmajor_empty(m, n, A.dtype, A.device), _core.colmajor_empty(n, n, A.dtype, A.device))


def check_input(A, WP):
    m, n = A.shape
    if m < n:
        raise ValueError("`left_polar!` requires a matrix A with at least as many rows as columns")  # polar.jl:9-10
    if not _core.is_colmajor(A):
        raise ValueError("A: column-major matrix expected")
    _core.dtype_code(A)
    W, P = WP
    if tuple(W.shape) != (m, n) or W.dtype != A.dtype or not _core.is_colmajor(W):
        raise ValueError(f"W: {m} x {n} column-major matrix expected")
    if P is not None and P.numel() > 0:
        if tuple(P.shape) != (n, n) or P.dtype != A.dtype or not _core.is_colmajor(P):
            raise ValueError(f"P: {n} x {n} column-major matrix (or an empty one) expected")


def left_polar_(A, WP=None, alg=None, **kw):
    """``left_polar!(A, (W,P), alg)``; a zero-length P means "skip P" (polar.jl:14,64,102)."""
    alg = select_algorithm("left_polar", A, alg, **kw)
    if WP is None:
        WP = initialize_output(A)
    check_input(A, WP)
    W, P = WP
    if isinstance(alg, Algorithm) and alg.name == "PolarViaSVD":
        return _left_polar_via_svd_(A, W, P, alg)
    if not isinstance(alg, Algorithm) or alg.name != "QDWH":
        raise ValueError(f"left_polar: algorithm {alg} is not provided by the B200 driver")
    if alg.get("tol") is not None:
        # the QDWH schedule (a, b, c per step) is fixed a priori from the lower bound l0 and always runs to |1 - l| <= 1e-15:
        # there is no iteration tolerance to set.  Say so instead of silently ignoring the keyword.
        raise ValueError("B200_QDWH runs an a-priori schedule to full precision: `tol` is not an option (use `maxiter` to cap the steps)")
    m, n = A.shape
    if m == 0 or n == 0:
        return W, P
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    want_p = P is not None and P.numel() > 0
    # QDWH converges to a PARTIAL isometry when A is singular (zero singular values stay zero); the reference's
    # PolarViaSVD returns an isometric W for any input.  Keep a copy of A (the C call destroys it) so that the
    # rare singular case can take the PolarViaSVD recipe on the rank-robust B200 SVD.
    A_keep = _core.colmajor_empty(m, n, A.dtype, A.device)
    A_keep.copy_(A)
    lw = h.lib.makb200_polar_worksize(h.h, dt, m, n)
    work = h.workspace(lw)
    iters = C.c_int(0)
    rc = h.lib.makb200_polar_qdwh(h.h, dt, m, n, _core.ptr(A), _core.ld(A), _core.ptr(W), _core.ld(W),
                                  _core.ptr(P) if want_p else C.c_void_p(0), _core.ld(P) if want_p else 0,
                                  0.0,   # l0 <= 0: the library picks the lower bound (estimate for n >= 1024, eps otherwise)
                                  int(alg.get("maxiter") or 0), _core.ptr(work), work.numel(), C.byref(iters),
                                  C.c_void_p(0))
    h.check(rc, "makb200_polar_qdwh")
    fro2 = torch.empty(1, dtype=torch.float64, device=A.device)
    rc = h.lib.makb200_fro2(h.h, dt, m, n, _core.ptr(W), _core.ld(W), _core.ptr(fro2))
    h.check(rc, "makb200_fro2")
    w2 = float(fro2.item())            # ||W||_F^2 = n for an isometry, = rank(A) for a partial isometry
    if not (abs(w2 - n) <= 1e-9 * n):  # also catches directions QDWH has not fully converged on (and NaN)
        from .algorithms import SVDViaPolar
        return _left_polar_via_svd_(A_keep, W, P, Algorithm("PolarViaSVD", {"svd_alg": SVDViaPolar()}))
    return W, P


def _left_polar_via_svd_(A, W, P, alg):
    """``PolarViaSVD`` recipe (polar.jl:59-70) on the B200 SVD and DMMA GEMM."""
    from .gemm import gemm_
    from .svd import svd_compact_
    U, S, Vh = svd_compact_(A, None, alg.get("svd_alg"))
    gemm_(W, U, Vh)
    if P is not None and P.numel() > 0:
        B = _core.colmajor_empty(Vh.shape[0], Vh.shape[1], Vh.dtype, Vh.device)
        B.copy_(Vh * torch.sqrt(S).to(Vh.dtype)[:, None])
        gemm_(P, B, B, opa="C", opb="N")
    return W, P


def left_polar(A, alg=None, **kw):
    from .qr import copy_input
    return left_polar_(copy_input(A), None, alg, **kw)
