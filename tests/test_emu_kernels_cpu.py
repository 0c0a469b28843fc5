"""CPU logic tests of the emulatable kernel headers (csrc/*.cuh written against the CUDA subset of
tests/cpu_harness/cuda_emu.h): the same device code that nvcc compiles into libmakb200 is compiled
with g++, every CUDA thread is a fiber, and the results are compared with the LAPACK oracle.  Each
case runs under forward, reverse and random fiber orders (a missing barrier is order dependent).
These tests validate kernel LOGIC without a GPU; timing and memory-model behaviour are measured on
the B200 by the -m gpu tests."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import mak_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(float).eps
ORDERS = [(0, 0), (1, 0), (2, 7)]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "emu_kernels_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "emu_kernels_host.cpp")])
    return ctypes.CDLL(out)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ptrs(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data if a is not None else None for a in arrs])


def _ints(v):
    return (ctypes.c_int * len(v))(*v)


def test_emulator_collectives(lib):
    rng = np.random.default_rng(0)
    for order, seed in ORDERS:
        nt = 96
        out = np.zeros(nt * 5)
        Am, Bm, Cm = rng.standard_normal((8, 4)), rng.standard_normal((4, 8)), rng.standard_normal((8, 8))
        C0 = Cm.copy()
        lib.emu_selftest(nt, _vp(out), _vp(Am), _vp(Bm), _vp(Cm), order, ctypes.c_uint64(seed))
        out = out.reshape(nt, 5)
        t = np.arange(nt)
        ws = np.array([(t[w * 32:(w + 1) * 32] + 1).sum() for w in range(3)])
        assert np.array_equal(out[:, 0], ws[t // 32])
        assert np.all(out[:, 1] == (t + 1).sum())
        assert np.all(out[:, 2] == sum(1 << l for l in range(32) if l % 3 == 0))
        dn = np.where(t % 32 == 31, t, t + 1)
        assert np.array_equal(out[:, 3], dn)
        lane = t % 32
        assert np.array_equal(out[:, 4], (lane & 16) | ((lane & 15) ^ 5))
        assert np.allclose(Cm, C0 + Am @ Bm, rtol=0, atol=1e-14)


def _qr_blocks(shapes, dtype, seed):
    As = [np.asfortranarray(O.randn_matrix(m, n, dtype, seed=seed + i)) for i, (m, n) in enumerate(shapes)]
    return As


def _run_bqr(lib, fn_variant, As, order, seed, want_r=True):
    npdt = As[0].dtype
    dt = 1 if np.iscomplexobj(As[0]) else 0
    Ain = [a.copy(order="F") for a in As]
    Qs = [np.zeros((a.shape[0], min(a.shape)), dtype=npdt, order="F") for a in As]
    Rs = [np.zeros((min(a.shape), a.shape[1]), dtype=npdt, order="F") if want_r else None for a in As]
    rc = lib.emu_batched_qr_warp(dt, fn_variant, len(As), _ints([a.shape[0] for a in As]), _ints([a.shape[1] for a in As]),
                                 _ptrs(Ain), _ints([a.shape[0] for a in As]), _ptrs(Qs), _ints([q.shape[0] for q in Qs]),
                                 _ptrs(Rs), _ints([r.shape[0] if r is not None else 0 for r in Rs]),
                                 order, ctypes.c_uint64(seed))
    assert rc == 0
    return Qs, Rs


def _check_qr(As, Qs, Rs):
    for A, Q, R in zip(As, Qs, Rs):
        m, n = A.shape
        tol = O.tol_for(m, n)
        Qo, Ro = O.qr_compact(A.copy())
        assert O.orth_err(Q) <= tol
        if R is not None:
            assert O.rel_resid(A, Q, R) <= tol
            assert np.all(np.tril(R, -1) == 0)
            assert np.all(np.real(np.diag(R)) >= 0) and np.all(np.imag(np.diag(R)) == 0)
            assert np.linalg.norm(R - Ro) <= 200 * tol * np.linalg.norm(Ro)
        assert np.linalg.norm(Q - Qo) <= 200 * tol * np.sqrt(min(m, n))


SHAPES = [(16, 16), (32, 32), (23, 23), (32, 17), (17, 32), (5, 3), (1, 1), (31, 32), (24, 24), (2, 7)]


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("variant", [0, 1])
def test_existing_warp_qr_kernels_in_emulation(lib, dtype, variant):
    """The GPU-verified kernels must also pass here: this pins the emulator itself."""
    As = _qr_blocks(SHAPES, dtype, 100)
    for order, seed in ORDERS:
        Qs, Rs = _run_bqr(lib, variant, As, order, seed)
        _check_qr(As, Qs, Rs)
    Qs, Rs = _run_bqr(lib, variant, As, 0, 0, want_r=False)
    _check_qr(As, Qs, Rs)
