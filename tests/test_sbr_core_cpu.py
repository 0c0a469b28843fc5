"""Second stage of the two-stage tridiagonalisation (csrc/sbr_core.h, round-2 groundwork): the
product header compiled with g++ and driven by tests/cpu_harness/sbr_host.cpp.  Checks the chase
(band -> tridiagonal), the wavefront independence rule (bit-identical results in a randomised
wavefront order), and the diamond-blocked application of Q2 against the one-reflector-at-a-time
reference, for Float64 and ComplexF64."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "sbr_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "sbr_host.cpp")])
    return ctypes.CDLL(out)


def _band(n, b, dtype, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    if dtype == "c128":
        A = A + 1j * rng.standard_normal((n, n))
    A = (A + A.conj().T) / 2
    i, j = np.indices((n, n))
    A[np.abs(i - j) > b] = 0
    return np.asfortranarray(A)


def _chase(lib, A, b, order=0, seed=0):
    n = A.shape[0]
    dt = 1 if np.iscomplexobj(A) else 0
    ldt = (n + b - 1) // b + 1
    d, e = np.zeros(n), np.zeros(max(n - 1, 1))
    V2 = np.zeros((n, n), dtype=A.dtype, order="F")
    tau2 = np.zeros((ldt, n), dtype=A.dtype, order="F")
    vp = ctypes.c_void_p
    bad = lib.sbr_host_chase(dt, n, b, A.ctypes.data_as(vp), n, d.ctypes.data_as(vp), e.ctypes.data_as(vp),
                             V2.ctypes.data_as(vp), tau2.ctypes.data_as(vp), ldt, order, seed)
    return bad, d, e[:n - 1], V2, tau2


def _apply_q2(lib, V2, tau2, b, Z, mode, g=1, seed=0):
    n = V2.shape[0]
    dt = 1 if np.iscomplexobj(V2) else 0
    X = np.asfortranarray(Z.astype(V2.dtype))
    vp = ctypes.c_void_p
    rc = lib.sbr_host_apply_q2(dt, n, b, V2.ctypes.data_as(vp), tau2.ctypes.data_as(vp), tau2.shape[0],
                               X.ctypes.data_as(vp), n, X.shape[1], mode, g, seed)
    assert rc == 0
    return X


CASES = [(1, 1), (2, 1), (3, 2), (5, 8), (17, 4), (33, 32), (64, 16), (100, 7), (200, 16), (130, 64)]


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,b", CASES)
def test_chase_and_q2(lib, n, b, dtype):
    A = _band(n, b, dtype, seed=n * 131 + b)
    bad, d, e, V2, tau2 = _chase(lib, A, b)
    assert bad == 0                       # exactly tridiagonal, real off-diagonal
    tol = 10 * max(n, 2) * EPS
    wref = np.linalg.eigvalsh(A)
    nrm = max(np.abs(wref).max(), 1e-300)
    if n > 1:
        w, Z = eigh_tridiagonal(d, e)
    else:
        w, Z = d.copy(), np.ones((1, 1))
    assert np.max(np.abs(w - wref)) / nrm <= tol
    # randomised wavefront order: the independence rule must make the result bit-identical
    bad2, d2, e2, V22, tau22 = _chase(lib, A, b, order=1, seed=7)
    assert bad2 == 0
    assert np.array_equal(d, d2) and np.array_equal(e, e2) and np.array_equal(V2, V22) and np.array_equal(tau2, tau22)
    # eigenvectors of the band matrix: X = Q2 Z
    X = _apply_q2(lib, V2, tau2, b, Z, mode=0)
    assert np.linalg.norm(A @ X - X * w) / nrm <= tol * np.sqrt(n)
    assert np.linalg.norm(X.conj().T @ X - np.eye(n)) <= tol * np.sqrt(n)
    # diamond blocking, several group sizes, random order inside a diamond wavefront
    for g in sorted({1, 3, max(b // 2, 1), b}):
        Xd = _apply_q2(lib, V2, tau2, b, Z, mode=1, g=g, seed=g)
        assert np.linalg.norm(Xd - X) <= tol * np.sqrt(n)


def test_two_stage_pipeline_numpy_stage1(lib):
    """dense -> band with a plain numpy stage 1 (panel QR + two-sided update), then the product's
    stage 2: A X = X diag(w) with X = Q1 Q2 Z."""
    n, b = 96, 8
    rng = np.random.default_rng(5)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (G + G.conj().T) / 2
    B, Q1 = A.copy(), np.eye(n, dtype=complex)
    for j0 in range(0, n - b - 1, b):
        r0 = j0 + b
        Qp, _ = np.linalg.qr(B[r0:, j0:j0 + b], mode="complete")
        U = np.eye(n, dtype=complex)
        U[r0:, r0:] = Qp
        B = U.conj().T @ B @ U
        Q1 = Q1 @ U
    i, j = np.indices((n, n))
    assert np.abs(B[np.abs(i - j) > b]).max() <= 1e-12
    B[np.abs(i - j) > b] = 0
    B = np.asfortranarray((B + B.conj().T) / 2)
    bad, d, e, V2, tau2 = _chase(lib, B, b)
    assert bad == 0
    w, Z = eigh_tridiagonal(d, e)
    X = Q1 @ _apply_q2(lib, V2, tau2, b, Z, mode=1, g=b)
    tol = 10 * n * EPS * np.sqrt(n)
    assert np.linalg.norm(A @ X - X * w) / np.abs(w).max() <= tol
    assert np.linalg.norm(X.conj().T @ X - np.eye(n)) <= tol
