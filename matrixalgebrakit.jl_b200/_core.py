"""Handle + device-matrix plumbing for the host layer (torch is used only for device memory,
streams and torch.distributed; no torch op touches the numerical path)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_DT = {torch.float64: _lib.F64, torch.complex128: _lib.C128}


class MakError(RuntimeError):
    pass


def dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        # the reference's CUDA ext instantiates F32/C32 too; this backend is
        # Float64/ComplexF64 only and says so instead of falling through to another path
        raise TypeError(f"makb200 supports Float64/ComplexF64 device matrices only, got {t.dtype}")


def colmajor_empty(m, n, dtype, device):
    """m x n column-major (Julia layout) matrix: a (n, m) contiguous buffer viewed transposed."""
    return torch.empty((n, m), dtype=dtype, device=device).t()


def colmajor_zeros(m, n, dtype, device):
    return torch.zeros((n, m), dtype=dtype, device=device).t()


def as_colmajor(A):
    """Return a column-major, unit-stride-in-dim-1 device matrix with A's values (copy if needed)."""
    if A.dim() != 2:
        raise ValueError("matrix expected")
    m, n = A.shape
    if is_colmajor(A):
        return A
    out = colmajor_empty(m, n, A.dtype, A.device)
    out.copy_(A)
    return out


def is_colmajor(A):
    m, n = A.shape
    if m == 0 or n == 0:
        return True
    return (A.stride(0) == 1 or m == 1) and (n == 1 or A.stride(1) >= max(1, m))


def ld(A):
    """lda = stride(A, 2) (yalapack.jl:179)."""
    m, n = A.shape
    if n <= 1:
        return max(1, m)
    return max(A.stride(1), 1)


def to_device(a, device="cuda:0", pinned=None):
    """numpy (any order) -> column-major device tensor. H2D copy on the current stream."""
    a = np.asarray(a)
    dt = np.complex128 if np.iscomplexobj(a) else np.float64
    af = np.asfortranarray(a, dtype=dt)
    t = torch.from_numpy(af.T)  # (n, m) C-contiguous view of the Fortran buffer
    return t.to(device, non_blocking=True).t()


def to_numpy(A):
    """column-major device tensor -> Fortran-ordered numpy array."""
    return A.t().contiguous().cpu().numpy().T


class Handle:
    """One per (process, device): owns the C handle and a grow-only workspace, the way the
    reference borrows CUDA.jl's handle-cached workspace (yacusolver.jl:76-90)."""

    _cache = {}

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MakError("makb200 needs a CUDA device; there is no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        hp = C.c_void_p()
        rc = self.lib.makb200_create(C.byref(hp), idx)
        if rc != 0:
            raise MakError(f"makb200_create failed with code {rc} (needs an sm_100 device)")
        self.h = hp
        self._work = None

    @classmethod
    def get(cls, device):
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx not in cls._cache:
            cls._cache[idx] = Handle(torch.device("cuda", idx))
        h = cls._cache[idx]
        h.lib.makb200_set_stream(h.h, C.c_void_p(torch.cuda.current_stream(h.device).cuda_stream))
        return h

    def workspace(self, nbytes):
        if self._work is None or self._work.numel() < nbytes:
            self._work = None
            self._work = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return self._work

    def check(self, rc, what):
        if rc == 0:
            return
        if rc < 0:
            raise ValueError(f"{what}: argument {-rc} had an illegal value")  # LAPACK info<0 (chkargsok)
        msg = self.lib.makb200_last_error(self.h)
        raise MakError(f"{what}: error {rc} {msg.decode() if msg else ''}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)
