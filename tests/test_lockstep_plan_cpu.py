"""Launch plan of the lock-step batched QDWH (csrc/polar_lockstep_plan.h) replayed on the CPU: the product planner
builds the action list and every grouped-GEMM descriptor; tests/cpu_harness/lockstep_host.cpp executes them with naive
loops (NaN above the diagonal wherever the GPU kernel skips tiles).  Checked per block against numpy: W^H W = I,
W P = A, P Hermitian positive semidefinite, W equal to the SVD-based polar factor.  Tolerance 10 n eps scaled by the
conditioning as in tests/test_gpu_svd_polar.py."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "lockstep_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "lockstep_host.cpp")])
    return ctypes.CDLL(out)


def _rand(m, n, cplx, rng):
    a = rng.standard_normal((m, n))
    if cplx:
        a = a + 1j * rng.standard_normal((m, n))
    return np.asfortranarray(a)


def _run(lib, mats, cplx, nb, estimate=1):
    order = sorted(range(len(mats)), key=lambda i: -mats[i].shape[1])
    mats = [mats[i] for i in order]
    count = len(mats)
    dt = np.complex128 if cplx else np.float64
    W = [np.asfortranarray(np.full(a.shape, np.nan, dtype=dt)) for a in mats]
    P = [np.asfortranarray(np.full((a.shape[1], a.shape[1]), np.nan, dtype=dt)) for a in mats]
    m = (ctypes.c_int * count)(*[a.shape[0] for a in mats])
    n = (ctypes.c_int * count)(*[a.shape[1] for a in mats])
    vp = ctypes.c_void_p
    pa = (vp * count)(*[a.ctypes.data for a in mats])
    pw = (vp * count)(*[a.ctypes.data for a in W])
    pp = (vp * count)(*[a.ctypes.data for a in P])
    na, ng, nf, l0 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    rc = lib.lockstep_replay(1 if cplx else 0, count, m, n, pa, pw, pp, nb, estimate, ctypes.byref(na), ctypes.byref(ng),
                             ctypes.byref(nf), ctypes.byref(l0))
    assert rc == 0
    _run.last = (nf.value, l0.value)
    return mats, W, P, na.value, ng.value


def _check(a, w, p, cond_scale=1.0):
    m, n = a.shape
    tol = 10 * max(m, n) * EPS
    assert np.all(np.isfinite(w)) and np.all(np.isfinite(p))
    assert np.linalg.norm(w.conj().T @ w - np.eye(n)) <= tol
    assert np.linalg.norm(w @ p - a) <= tol * np.linalg.norm(a)
    assert np.array_equal(p, p.conj().T)
    assert np.linalg.eigvalsh(p).min() >= -tol * np.linalg.norm(p, 2)
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    assert np.linalg.norm(w - u @ vh) <= 50 * tol * cond_scale


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("nb", [4, 8])
@pytest.mark.parametrize("estimate", [1, 0])
def test_square_ragged_chunk(lib, cplx, nb, estimate):
    rng = np.random.default_rng(3 + nb)
    sizes = [3, 5, 8, 9, 16, 17, 23, 24, 31, 7, 12]
    mats = [_rand(s, s, cplx, rng) for s in sizes]
    mats, W, P, na, ng = _run(lib, mats, cplx, nb, estimate)
    for a, w, p in zip(mats, W, P):
        _check(a, w, p, max(1.0, np.linalg.cond(a) / 100))
    if estimate:
        # every Gaussian block has a usable estimate; the chunk's l0 is a lower bound of every block's sigma_min(X0)
        nf, l0 = _run.last
        assert nf == len(sizes)
        assert all(l0 <= np.linalg.svd(a, compute_uv=False)[-1] / np.linalg.norm(a) for a in mats)
    else:
        assert _run.last[0] == 0
    # the descriptor storage bound used by the workspace query covers the plan
    assert ng <= lib.lockstep_launch_bound(1 if cplx else 0, max(sizes), nb, 0)


@pytest.mark.parametrize("cplx", [False, True])
def test_tall_and_square_mixed(lib, cplx):
    rng = np.random.default_rng(11)
    shapes = [(20, 20), (31, 20), (26, 13), (13, 13), (40, 9), (9, 9), (10, 9)]
    mats = [_rand(m, n, cplx, rng) for m, n in shapes]
    mats, W, P, na, ng = _run(lib, mats, cplx, 8)
    for a, w, p in zip(mats, W, P):
        _check(a, w, p, max(1.0, np.linalg.cond(a) / 100))
    assert ng <= lib.lockstep_launch_bound(1 if cplx else 0, 20, 8, 1)


def test_graded_and_rank_deficient(lib):
    rng = np.random.default_rng(5)
    n = 24
    q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
    q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    graded = np.asfortranarray((q1 * 10.0 ** (-10 * np.arange(n) / n)) @ q2)       # kappa = 1e10
    lowrank = np.asfortranarray(rng.standard_normal((n, 3)) @ rng.standard_normal((3, n)))
    tiny = np.asfortranarray(1e-200 * rng.standard_normal((n, n)))
    easy = _rand(n, n, False, rng)
    mats, W, P, _, _ = _run(lib, [graded, lowrank, tiny, easy], False, 8)
    # kappa = 1e10 and rank 3 fall back to the l0 = eps schedule; the tiny-norm Gaussian block and the plain one do not
    assert _run.last[0] == 2
    tol = 10 * n * EPS
    for a, w, p in zip(mats, W, P):
        assert np.all(np.isfinite(w)) and np.all(np.isfinite(p))
        assert np.linalg.norm(w @ p - a) <= tol * np.linalg.norm(a)
        assert np.array_equal(p, p.T)
    # full rank (however ill-conditioned): W is an isometry
    for a, w in zip(mats, W):
        if np.linalg.matrix_rank(a) == n:
            assert np.linalg.norm(w.T @ w - np.eye(n)) <= tol
