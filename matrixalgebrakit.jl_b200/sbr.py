"""EXPERIMENTAL (round 1): second stage of a two-stage Hermitian tridiagonalisation.

``sbr_chase_(A, b)`` reduces the Hermitian band matrix held in the lower band (width ``b``) of ``A``
to a real symmetric tridiagonal ``T = Q2^H B Q2`` by bulge chasing on the GPU (``csrc/sbr.cu``) and
returns ``(d, e, V2, tau2)`` with the chase reflectors in the layout of ``csrc/sbr_core.h``.
``eigh_full!`` does not use it yet."""
import torch

from . import _core


def sbr_chase_(A, b):
    if not _core.is_colmajor(A) or A.shape[0] != A.shape[1]:
        raise ValueError("A: square column-major matrix expected")
    if not (1 <= int(b) <= 64):
        raise ValueError("bandwidth b must be in 1..64")
    n = A.shape[0]
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    ldt = (n + b - 1) // b + 1
    d = torch.empty(n, dtype=torch.float64, device=A.device)
    e = torch.empty(max(n - 1, 1), dtype=torch.float64, device=A.device)
    V2 = _core.colmajor_empty(max(n, 1), max(n, 1), A.dtype, A.device)
    tau2 = _core.colmajor_empty(ldt, max(n, 1), A.dtype, A.device)
    lw = h.lib.makb200_sbr_chase_worksize(h.h, dt, n, int(b))
    work = h.workspace(lw)
    rc = h.lib.makb200_sbr_chase(h.h, dt, n, int(b), _core.ptr(A), _core.ld(A), _core.ptr(d), _core.ptr(e),
                                 _core.ptr(V2), _core.ld(V2), _core.ptr(tau2), _core.ld(tau2), _core.ptr(work),
                                 work.numel())
    h.check(rc, "makb200_sbr_chase")
    return d, e[:max(n - 1, 0)], V2, tau2


def sy2sb_(A, b):
    """EXPERIMENTAL first stage: full Hermitian ``A`` (both triangles) -> band (lower bandwidth ``b``) in
    place; returns ``tau1`` (reflectors are left below the band)."""
    n = A.shape[0]
    h = _core.Handle.get(A.device)
    dt = _core.dtype_code(A)
    tau1 = torch.zeros(max(n, 1), dtype=A.dtype, device=A.device)
    lw = h.lib.makb200_sy2sb_worksize(h.h, dt, n, int(b))
    work = h.workspace(lw)
    rc = h.lib.makb200_sy2sb(h.h, dt, n, int(b), _core.ptr(A), _core.ld(A), _core.ptr(tau1), _core.ptr(work), work.numel())
    h.check(rc, "makb200_sy2sb")
    return tau1


def sbr_apply_q2_(V2, tau2, b, Z, g=None):
    """EXPERIMENTAL: ``Z <- Q2 Z`` (diamond-blocked, grouped DMMA GEMMs)."""
    n = V2.shape[0]
    g = int(g or b)
    h = _core.Handle.get(Z.device)
    dt = _core.dtype_code(Z)
    lw = h.lib.makb200_sbr_apply_q2_worksize(h.h, dt, n, int(b), g, Z.shape[1])
    work = h.workspace(lw)
    rc = h.lib.makb200_sbr_apply_q2(h.h, dt, n, int(b), g, _core.ptr(V2), _core.ld(V2), _core.ptr(tau2), _core.ld(tau2),
                                    _core.ptr(Z), _core.ld(Z), Z.shape[1], _core.ptr(work), work.numel())
    h.check(rc, "makb200_sbr_apply_q2")
    return Z
