"""The values-only tridiagonal solver (csrc/sturm_core.h: Sturm counts + K-section) compiled with g++
and run with a plain loop over k, exactly as the CUDA kernel runs one thread per k."""
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
    out = os.path.join(tempfile.mkdtemp(), "sturm_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "sturm_host.cpp")])
    L = ctypes.CDLL(out)
    L.sturm_count_host.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    return L


def _cases():
    rng = np.random.default_rng(0)
    yield "n1", np.array([3.5]), np.zeros(0)
    yield "n2", np.array([1.0, -2.0]), np.array([0.5])
    yield "rand20", rng.standard_normal(20), rng.standard_normal(19)
    yield "rand257", rng.standard_normal(257), rng.standard_normal(256)
    yield "rand1000", rng.standard_normal(1000), rng.standard_normal(999)
    yield "wilkinson", np.abs(np.arange(-100, 101)).astype(float), np.ones(200)
    yield "clustered", np.ones(300), np.full(299, 1e-3)
    yield "diagonal", rng.standard_normal(150), np.zeros(149)
    d = np.tile(np.arange(16.0), 16)
    e = np.ones(255)
    e[15::16] = 1e-9
    yield "glued", d, e
    d = 10.0 ** (-np.arange(200) / 15.0)
    yield "graded", d, 0.1 * d[:-1]
    yield "toeplitz", np.full(500, 2.0), np.full(499, -1.0)
    yield "zero", np.zeros(70), np.zeros(69)
    yield "identity", np.ones(64), np.zeros(63)
    yield "huge", 1e200 * rng.standard_normal(100), 1e200 * rng.standard_normal(99)
    yield "tiny", 1e-200 * rng.standard_normal(100), 1e-200 * rng.standard_normal(99)
    yield "negdef", -np.abs(rng.standard_normal(90)) - 3.0, rng.standard_normal(89)
    # hetrd-like: chi-distributed off-diagonals of a GOE matrix, n = 2048
    n = 2048
    yield "goe2048", rng.standard_normal(n), np.sqrt(rng.chisquare(np.arange(n - 1, 0, -1)) / 2)


CASES = list(_cases())


@pytest.mark.parametrize("name,d,e", CASES, ids=[c[0] for c in CASES])
def test_sturm_eigvals(lib, name, d, e):
    n = len(d)
    d = np.ascontiguousarray(d, dtype=np.float64)
    e = np.ascontiguousarray(np.append(e, 0.0), dtype=np.float64)  # never read past n-1
    e[-1] = np.nan
    w = np.zeros(n)
    vp = ctypes.c_void_p
    assert lib.sturm_host(n, d.ctypes.data_as(vp), e.ctypes.data_as(vp), w.ctypes.data_as(vp)) == 0
    wref = eigh_tridiagonal(d, e[:-1], eigvals_only=True) if n > 1 else d.copy()
    nrm = max(np.abs(wref).max(), 1e-300)
    assert np.all(np.isfinite(w))
    assert np.all(np.diff(w) >= 0), "ascending order without a sort"
    assert np.max(np.abs(w - wref)) / nrm <= 4 * EPS * max(1, np.sqrt(n)), np.max(np.abs(w - wref)) / nrm


def test_sturm_count_matches_sorted_spectrum(lib):
    rng = np.random.default_rng(5)
    n = 300
    d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    wref = eigh_tridiagonal(d, e, eigvals_only=True)
    vp = ctypes.c_void_p
    for x in np.concatenate([[-100.0, 100.0], (wref[:-1] + wref[1:])[::17] / 2]):
        assert lib.sturm_count_host(n, d.ctypes.data_as(vp), e.ctypes.data_as(vp), float(x)) == int(np.sum(wref < x))
