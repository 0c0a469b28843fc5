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


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_panel_blocked_warp_qr_kernel(lib, dtype):
    """Round-2 bring-up kernel batched_qr_warp_blk_kernel (4-column panels, lane = row for the panel, two-pass
    block reflector for the trailing columns): same gauge-fixed Q, R as the LAPACK oracle."""
    shapes = SHAPES + [(32, 4), (32, 5), (4, 32), (9, 9), (12, 8), (8, 12), (32, 31), (30, 3), (3, 30), (1, 5), (6, 1),
                       (20, 20), (28, 27)]
    As = _qr_blocks(shapes, dtype, 300)
    rng = np.random.default_rng(1)
    # structured blocks: zero, identity-like, rank deficient, graded columns, already triangular
    As.append(np.zeros((12, 9), dtype=As[0].dtype, order="F"))
    As.append(np.asfortranarray(np.eye(16, 16).astype(As[0].dtype)))
    lowr = rng.standard_normal((20, 3)) @ rng.standard_normal((3, 14))
    As.append(np.asfortranarray(lowr.astype(As[0].dtype)))
    As.append(np.asfortranarray((O.randn_matrix(24, 24, dtype, 9) * 10.0 ** (-np.arange(24) / 2.0))))
    As.append(np.asfortranarray(np.triu(O.randn_matrix(10, 10, dtype, 10))))
    for order, seed in ORDERS:
        Qs, Rs = _run_bqr(lib, 2, As, order, seed)
        for A, Q, R in zip(As, Qs, Rs):
            m, n = A.shape
            tol = O.tol_for(m, n)
            assert O.orth_err(Q) <= tol, (A.shape, O.orth_err(Q))
            assert O.rel_resid(A, Q, R) <= tol or np.linalg.norm(A) == 0
            assert np.all(np.tril(R, -1) == 0)
            assert np.all(np.real(np.diag(R)) >= 0) and np.all(np.imag(np.diag(R)) == 0)
        _check_qr(As[:len(shapes)], Qs[:len(shapes)], Rs[:len(shapes)])      # full-rank blocks: factors equal the oracle's
    Qs, Rs = _run_bqr(lib, 2, As[:len(shapes)], 0, 0, want_r=False)
    _check_qr(As[:len(shapes)], Qs, Rs)


# ---- persistent bulge chasing (csrc/sbr_chase_persistent.cuh): all CTAs co-resident as fibers ----
@pytest.fixture(scope="module")
def sbr_lib():
    out = os.path.join(tempfile.mkdtemp(), "sbr_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "sbr_host.cpp")])
    return ctypes.CDLL(out)


def _band_matrix(n, b, dtype, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    if dtype == "c128":
        A = A + 1j * rng.standard_normal((n, n))
    A = (A + A.conj().T) / 2
    i, j = np.indices((n, n))
    A[np.abs(i - j) > b] = 0
    return np.asfortranarray(A)


def _band_pack(A, b):
    n = A.shape[0]
    AB = np.zeros((2 * b, n), dtype=A.dtype, order="F")
    for j in range(n):
        hi = min(n, j + b + 1)
        AB[:hi - j, j] = A[j:hi, j]
        AB[0, j] = AB[0, j].real
    return AB


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,b,grid", [(33, 4, 3), (40, 8, 5), (19, 3, 24), (50, 16, 2), (9, 4, 1), (2, 4, 4), (70, 33, 3)])
def test_emu_chase_persistent(lib, sbr_lib, dtype, n, b, grid):
    from scipy.linalg import eigh_tridiagonal
    A = _band_matrix(n, b, dtype, seed=n + b)
    dt = 1 if dtype == "c128" else 0
    ldt = (n + b - 1) // b + 1
    # sequential reference (sbr_core.h task bodies in generation order)
    d0, e0 = np.zeros(n), np.zeros(max(n - 1, 1))
    V0 = np.zeros((n, n), dtype=A.dtype, order="F")
    t0 = np.zeros((ldt, n), dtype=A.dtype, order="F")
    bad = sbr_lib.sbr_host_chase(dt, n, b, _vp(A), n, _vp(d0), _vp(e0), _vp(V0), _vp(t0), ldt, 0, 0)
    assert bad == 0
    ev = np.linalg.eigvalsh(A)
    scale = max(np.abs(ev).max(), 1.0)
    for order, seed in ORDERS:
        AB = _band_pack(A, b)
        V2 = np.zeros((n, n), dtype=A.dtype, order="F")
        tau2 = np.zeros((ldt, n), dtype=A.dtype, order="F")
        rc = lib.emu_chase_persistent(dt, n, b, _vp(AB), 2 * b, _vp(V2), n, _vp(tau2), ldt, grid, order,
                                      ctypes.c_uint64(seed))
        assert rc == 0
        assert np.all(AB[2:, :] == 0), "fill below the first subdiagonal"
        assert np.all(AB[:2, :].imag == 0) if dt else True
        d, e = AB[0, :].real.copy(), AB[1, :n - 1].real.copy()
        w = eigh_tridiagonal(d, e, eigvals_only=True) if n > 1 else d
        assert np.abs(w - ev).max() <= 50 * n * EPS * scale
        assert np.abs(d - d0).max() <= 1e-11 * scale and np.abs(e - e0[:n - 1]).max() <= 1e-11 * scale
        assert np.abs(V2 - V0).max() <= 1e-10 and np.abs(tau2 - t0).max() <= 1e-10


# ---- fused Q2 application (csrc/sbr_q2_slab.cuh): DMMA fragments, 2b-row ring, zero skipping ----
@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,b,g,cw,ncols", [(50, 8, 8, 32, 40), (45, 16, 8, 32, 32), (61, 8, 8, 64, 70), (40, 16, 16, 64, 9),
                                             (2, 8, 8, 32, 5), (9, 8, 8, 32, 33)])
def test_emu_q2_slab(lib, sbr_lib, dtype, n, b, g, cw, ncols):
    A = _band_matrix(n, b, dtype, seed=3 * n + b)
    dt = 1 if dtype == "c128" else 0
    ldt = (n + b - 1) // b + 1
    d0, e0 = np.zeros(n), np.zeros(max(n - 1, 1))
    V2 = np.zeros((n, n), dtype=A.dtype, order="F")
    tau2 = np.zeros((ldt, n), dtype=A.dtype, order="F")
    assert sbr_lib.sbr_host_chase(dt, n, b, _vp(A), n, _vp(d0), _vp(e0), _vp(V2), _vp(tau2), ldt, 0, 0) == 0
    Z0 = np.asfortranarray(O.randn_matrix(n, ncols, dtype, seed=n))
    Xref = Z0.copy(order="F")
    assert sbr_lib.sbr_host_apply_q2(dt, n, b, _vp(V2), _vp(tau2), ldt, _vp(Xref), n, ncols, 0, 1, 0) == 0
    for order, seed in ORDERS:
        X = Z0.copy(order="F")
        assert lib.emu_q2_slab(dt, n, b, g, cw, _vp(V2), n, _vp(tau2), ldt, _vp(X), n, ncols, order,
                               ctypes.c_uint64(seed)) == 0
        assert np.abs(X - Xref).max() <= 200 * n * EPS * max(np.abs(Xref).max(), 1.0)
    # Q2 is unitary: the fused result keeps column norms
    assert np.allclose(np.linalg.norm(X, axis=0), np.linalg.norm(Z0, axis=0), rtol=1e-12)


# ---------------------------------------------------------------------------------------------------
# projections.cuh: project_hermitian! / ishermitian / isisometric kernels
# ---------------------------------------------------------------------------------------------------
def _ref_project(A, anti):
    """Entry for entry the reference's arithmetic (implementations/projections.jl:109-139)."""
    return (A - A.conj().T) / 2 if anti else (A + A.conj().T) / 2


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("anti", [0, 1])
@pytest.mark.parametrize("n,pad", [(1, 0), (5, 3), (32, 0), (33, 1), (70, 0), (97, 5)])
def test_emu_project_hermitian(lib, n, pad, anti, dtype):
    A0 = O.randn_matrix(n, n, dtype, seed=n + anti)
    ref = _ref_project(A0, anti)
    dt = 0 if dtype == "f64" else 1
    for order, seed in ORDERS:
        # out of place, strided input
        Abuf = np.zeros((n + pad, n), dtype=A0.dtype, order="F")
        Abuf[:n] = A0
        B = np.full((n, n), np.nan, dtype=A0.dtype, order="F")
        lib.emu_project_herm(dt, anti, n, _vp(Abuf), n + pad, _vp(B), n, order, ctypes.c_uint64(seed))
        assert np.array_equal(B, ref)                      # bit for bit
        assert np.array_equal(Abuf[:n], A0) and not Abuf[n:].any()
        # in place (B === A, the reference's default output)
        lib.emu_project_herm(dt, anti, n, _vp(Abuf), n + pad, _vp(Abuf), n + pad, order, ctypes.c_uint64(seed))
        assert np.array_equal(Abuf[:n], ref) and not Abuf[n:].any()
    d = np.diag(ref)
    assert np.all(d.real == 0) if anti else np.all(d.imag == 0)
    assert np.array_equal(ref, -ref.conj().T if anti else ref.conj().T)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("anti", [0, 1])
@pytest.mark.parametrize("n", [1, 7, 32, 45, 100])
def test_emu_herm_props(lib, n, anti, dtype):
    dt = 0 if dtype == "f64" else 1
    G = O.randn_matrix(n, n, dtype, seed=3 * n + anti)
    H = _ref_project(G, anti)                              # exactly (anti-)Hermitian
    out = np.zeros(4)
    for order, seed in ORDERS:
        lib.emu_herm_props(dt, anti, n, _vp(np.asfortranarray(H)), n, _vp(out), order, ctypes.c_uint64(seed))
        assert out[0] == 0 and out[3] == 0
        assert np.isclose(out[1], np.abs(H).max(), rtol=1e-15) and np.isclose(out[2], np.linalg.norm(H) ** 2, rtol=1e-13)
        lib.emu_herm_props(dt, anti, n, _vp(G), n, _vp(out), order, ctypes.c_uint64(seed))
        van = _ref_project(G, 1 - anti)                    # the part that must vanish
        assert np.isclose(out[0], np.linalg.norm(van) ** 2, rtol=1e-13)
        assert np.isclose(out[1], np.abs(G).max(), rtol=1e-15) and np.isclose(out[2], np.linalg.norm(G) ** 2, rtol=1e-13)
        want = -G.conj().T if anti else G.conj().T
        assert out[3] == np.count_nonzero(np.triu(G != want))
    # one perturbed entry below the diagonal is seen by the exact test, and a strided view is read correctly
    if n > 1:
        Hp = np.zeros((n + 3, n), dtype=H.dtype, order="F")
        Hp[:n] = H
        Hp[n - 1, 0] += 1e-9
        lib.emu_herm_props(dt, anti, n, _vp(Hp), n + 3, _vp(out), 0, ctypes.c_uint64(0))
        assert out[3] == 1 and np.isclose(out[0], 2 * (0.5e-9) ** 2, rtol=1e-6)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,grid", [(1, 1), (20, 1), (65, 3), (130, 8)])
def test_emu_gram_defect(lib, n, grid, dtype):
    dt = 0 if dtype == "f64" else 1
    A = O.randn_matrix(n + 5, n, dtype, seed=n)
    P = np.asfortranarray(A.conj().T @ A)
    out = np.zeros(2)
    for order, seed in ORDERS:
        lib.emu_gram_defect(dt, n, _vp(P), n, _vp(out), grid, order, ctypes.c_uint64(seed))
        assert np.isclose(out[0], np.linalg.norm(P) ** 2, rtol=1e-13)
        assert np.isclose(out[1], np.linalg.norm(P - np.eye(n)) ** 2, rtol=1e-13)
    Q, _ = np.linalg.qr(A)
    P = np.asfortranarray(Q.conj().T @ Q)
    lib.emu_gram_defect(dt, n, _vp(P), n, _vp(out), grid, 0, ctypes.c_uint64(0))
    assert np.sqrt(out[1]) <= 1e-13 * np.sqrt(out[0]) * n


# ---------------------------------------------------------------------------------------------------
# bhetrd.cuh: one-CTA-per-block Hermitian tridiagonalisation (round-2 bring-up kernel)
# ---------------------------------------------------------------------------------------------------
def _q_from_reflectors(Aout, tau):
    """Q = H_0 ... H_{n-2}, H_j = I - tau_j v_j v_j^H, v_j = [0_{j+1}; 1; Aout[j+2:, j]]."""
    n = Aout.shape[0]
    Q = np.eye(n, dtype=Aout.dtype)
    for j in range(n - 2, -1, -1):
        v = np.zeros(n, dtype=Aout.dtype)
        v[j + 1] = 1.0
        v[j + 2:] = Aout[j + 2:, j]
        Q -= tau[j] * np.outer(v, v.conj() @ Q)
    return Q


@pytest.mark.parametrize("mirror", [0, 1])
@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_emu_bhetrd_batch(lib, dtype, mirror):
    dt = 0 if dtype == "f64" else 1
    ns = [1, 2, 3, 17, 33, 64, 70]
    for order, seed in ORDERS[mirror:mirror + 2]:
        As0 = [O.rand_hermitian(n, dtype, seed=50 + n) for n in ns]
        As0[3] = np.asfortranarray(np.diag(np.arange(1.0, 18.0)).astype(As0[3].dtype))   # already tridiagonal: tau = 0 steps
        bd = np.diag(np.arange(1.0, 34.0)).astype(As0[4].dtype)                           # dense 10x10 block + diagonal rest:
        bd[:10, :10] = O.rand_hermitian(10, dtype, seed=9)                                # a tau = 0 step with an update pending
        bd[20:28, 20:28] = O.rand_hermitian(8, dtype, seed=10)                            # and reflectors starting again after it
        As0[4] = np.asfortranarray(bd)
        pads = [0, 1, 0, 3, 0, 5, 0]
        bufs = []
        for a, p in zip(As0, pads):
            b = np.zeros((a.shape[0] + p, a.shape[0]), dtype=a.dtype, order="F")
            b[:a.shape[0]] = a
            if mirror:
                b[np.tril_indices(a.shape[0], -1)] = 777.0     # uplo = 'U': the lower triangle is garbage on entry
            else:
                b[np.triu_indices(a.shape[0], 1)] = 777.0      # the upper triangle must never be read
            bufs.append(b)
        ds = [np.zeros(n) for n in ns]
        es = [np.zeros(max(n - 1, 1)) for n in ns]
        taus = [np.zeros(max(n - 1, 1), dtype=As0[0].dtype) for n in ns]
        rc = lib.emu_bhetrd(dt, len(ns), _ints(ns), _ptrs(bufs), _ints([b.shape[0] for b in bufs]), _ptrs(ds), _ptrs(es),
                            _ptrs(taus), mirror, order, ctypes.c_uint64(seed))
        assert rc == 0
        for a, b, d, e, tau, n in zip(As0, bufs, ds, es, taus, ns):
            assert not b[n:].any()
            if mirror:
                assert np.array_equal(b[:n][np.triu_indices(n, 1)], a[np.triu_indices(n, 1)])   # upper untouched
            else:
                assert np.all(b[np.triu_indices(n, 1)] == 777.0)
            T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
            assert np.all(e[:n - 1] >= 0)                               # non-negative-beta reflectors
            wref = np.linalg.eigvalsh(a)
            scale = max(np.abs(wref).max(), 1.0)
            assert np.max(np.abs(np.linalg.eigvalsh(T) - wref)) <= 10 * n * EPS * scale
            Q = _q_from_reflectors(b[:n], tau)
            assert np.linalg.norm(Q.conj().T @ Q - np.eye(n)) <= 10 * n * EPS
            assert np.linalg.norm(Q.conj().T @ a @ Q - T) <= 10 * n * EPS * np.linalg.norm(a)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,pad,grid", [(1, 1, 0, 1), (7, 5, 2, 1), (5, 7, 0, 2), (40, 40, 3, 3)])
def test_emu_tri_init(lib, m, n, pad, grid, dtype):
    dt = 0 if dtype == "f64" else 1
    A0 = O.randn_matrix(m, n, dtype, seed=m + n)
    want = [np.eye(m, n, dtype=A0.dtype), np.triu(A0), np.tril(A0)]
    for mode in range(3):
        buf = np.full((m + pad, n), 5.0, dtype=A0.dtype, order="F")
        buf[:m] = A0
        lib.emu_tri_init(dt, mode, m, n, _vp(buf), m + pad, grid, 0, ctypes.c_uint64(0))
        assert np.array_equal(buf[:m], want[mode]) and np.all(buf[m:] == 5.0)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,pad,grid", [(1, 1, 0, 1), (70, 33, 3, 2), (33, 70, 0, 5)])
def test_emu_fro2(lib, m, n, pad, grid, dtype):
    dt = 0 if dtype == "f64" else 1
    A0 = O.randn_matrix(m, n, dtype, seed=m + n)
    buf = np.full((m + pad, n), 9.0, dtype=A0.dtype, order="F")
    buf[:m] = A0
    out = np.zeros(1)
    lib.emu_fro2(dt, m, n, _vp(buf), m + pad, _vp(out), grid, 2, ctypes.c_uint64(3))
    assert np.isclose(out[0], np.linalg.norm(A0) ** 2, rtol=1e-13)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,pad", [(1, 0), (31, 2), (32, 0), (33, 1), (100, 0)])
def test_emu_mirror_lower(lib, n, pad, dtype):
    dt = 0 if dtype == "f64" else 1
    G = O.randn_matrix(n, n, dtype, seed=n)
    L = np.tril(G, -1)
    want = L + L.conj().T + np.diag(np.diag(G).real)
    for order, seed in ORDERS:
        buf = np.full((n + pad, n), 3.0, dtype=G.dtype, order="F")
        buf[:n] = G
        lib.emu_mirror_lower(dt, n, _vp(buf), n + pad, order, ctypes.c_uint64(seed))
        assert np.array_equal(buf[:n], want) and np.all(buf[n:] == 3.0)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_emu_col_norm_defect(lib, dtype):
    """The check that triggers the re-orthonormalisation of U for rank-deficient input (polar.cu: svd_tall)."""
    dt = 0 if dtype == "f64" else 1
    Q, _ = np.linalg.qr(O.randn_matrix(70, 19, dtype, seed=1))
    out = np.zeros(1)
    for order, seed in ORDERS:
        U = np.asfortranarray(Q)
        lib.emu_col_norm_defect(dt, 70, 19, _vp(U), 70, _vp(out), order, ctypes.c_uint64(seed))
        assert out[0] <= 1e-14
        U = np.asfortranarray(Q.copy())
        U[:, 11] *= 0.5                                   # a collapsed column: ||u||^2 = 0.25
        lib.emu_col_norm_defect(dt, 70, 19, _vp(U), 70, _vp(out), order, ctypes.c_uint64(seed))
        assert abs(out[0] - 0.75) <= 1e-14
        U[:, 3] = 0.0
        lib.emu_col_norm_defect(dt, 70, 19, _vp(U), 70, _vp(out), order, ctypes.c_uint64(seed))
        assert out[0] == 1.0
        U[5, 7] = np.nan
        lib.emu_col_norm_defect(dt, 70, 19, _vp(U), 70, _vp(out), order, ctypes.c_uint64(seed))
        assert out[0] == 1e300
