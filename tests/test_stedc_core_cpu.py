"""The product's D&C work-item bodies (csrc/stedc_core.h) compiled with g++ and driven by a plain
loop harness — CPU-only check of the deflation / secular-equation / Loewner logic."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "stedc_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "stedc_host.cpp")])
    return ctypes.CDLL(out)


def _cases():
    rng = np.random.default_rng(0)
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


@pytest.mark.parametrize("name,d,e", list(_cases()), ids=[c[0] for c in _cases()])
def test_stedc_core(lib, name, d, e):
    n = len(d)
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    st = (ctypes.c_int * 3)()
    vp = ctypes.c_void_p
    rc = lib.stedc_host(n, d.ctypes.data_as(vp), e.ctypes.data_as(vp), w.ctypes.data_as(vp), Z.ctypes.data_as(vp), st)
    assert rc == 0
    wref = eigh_tridiagonal(d, e, eigvals_only=True)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    nrm = max(np.abs(wref).max(), 1e-300)
    tol = 10 * n * np.finfo(float).eps
    assert np.max(np.abs(w - wref)) / nrm <= tol
    assert np.linalg.norm(T @ Z - Z * w) / nrm <= tol
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= tol
