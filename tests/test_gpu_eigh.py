"""eigh_full! on B200 vs the LAPACK-replay oracle.  Tolerances (north_star): ||AV - VD||/||A||,
||V^H V - I||_F and max|lambda - lambda_oracle|/max|lambda| <= 10*n*eps; gauge-fixed eigenvectors
are compared with the oracle's scaled by the spectral gap."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


def _eigh(A_np, **kw):
    import makb200
    A = makb200.to_device(A_np)
    D, V = makb200.eigh_full(A, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(makb200.to_numpy(A), A_np)
    return D.cpu().numpy(), makb200.to_numpy(V)


def _check(A, w, V, vec_cmp=True):
    n = A.shape[0]
    tol = O.tol_for(n)
    wo, Vo = O.eigh_full(A)
    nrm = max(np.abs(wo).max(), 1e-300)
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - wo)) / nrm <= tol
    assert np.linalg.norm(A @ V - V * w) / max(np.linalg.norm(A), 1e-300) <= tol
    assert O.orth_err(V) <= tol
    piv = O._argmaxabs_cols(V)
    assert np.all(piv.real > 0) and np.all(np.abs(piv.imag) <= 1e-15)
    if vec_cmp and n > 1:
        gap = np.minimum(np.diff(wo, prepend=-np.inf), np.diff(wo, append=np.inf))
        err = np.linalg.norm(V - Vo, axis=0)
        # perturbation bound: eps*||A||/gap per vector (skip near-ties of the arg-max pivot)
        ok = err <= 100 * n * O.EPS * nrm / gap + 1e-13
        mod = np.abs(Vo)
        top2 = np.sort(mod, axis=0)[-2:]
        tie = (top2[1] - top2[0]) < 1e-8
        assert np.all(ok | tie)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 54, 65, 200, 513, 1100])
def test_eigh_full_vs_oracle(n, dtype):
    A = O.rand_hermitian(n, dtype, seed=123 + n)
    w, V = _eigh(A)
    _check(A, w, V)


def test_eigh_doctest_kat():
    # docs/src/user_interface/truncations.md:19-21
    A = np.array([[2.0, 1, 0], [1, 3, 1], [0, 1, 4]])
    w, V = _eigh(A)
    np.testing.assert_allclose(w, [3 - np.sqrt(3), 3, 3 + np.sqrt(3)], rtol=0, atol=1e-14)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_special_spectra(dtype):
    n = 300
    Q, _ = O.qr_compact(O.randn_matrix(n, n, dtype, seed=3))
    for spec in (np.ones(n), np.repeat(np.arange(10.0), 30), 10.0 ** (-12 * np.arange(n) / n),
                 np.concatenate([np.zeros(150), np.linspace(1, 2, 150)])):
        A = (Q * spec) @ Q.conj().T
        A = (A + A.conj().T) / 2
        w, V = _eigh(A)
        _check(A, w, V, vec_cmp=False)
    for A in (np.zeros((40, 40)), np.eye(70), np.diag(np.arange(50.0)[::-1])):
        w, V = _eigh(A.astype(np.complex128) if dtype == "c128" else A)
        _check(A, w, V, vec_cmp=False)


def test_eigh_only_upper_triangle_is_read_and_tolerance():
    import makb200
    A = O.rand_hermitian(80, "c128", seed=8)
    B = A.copy()
    B[np.tril_indices(80, -1)] += 1e-14 * (1 + 1j)  # within default_hermitian_tol: accepted, lower ignored
    w, V = _eigh(B)
    Au = np.triu(A) + np.triu(A, 1).conj().T
    np.testing.assert_allclose(w, np.linalg.eigvalsh(Au), atol=1e-12)
    G = O.randn_matrix(30, 30, "f64", 4)
    with pytest.raises(makb200.DomainError):       # eigh.jl:11-18
        makb200.eigh_full(makb200.to_device(G))
    with pytest.raises(ValueError):
        makb200.eigh_full(makb200.to_device(G), alg="QRIteration")


def test_eigh_fixgauge_false_vals_trunc_and_outputs():
    import makb200
    A0 = O.rand_hermitian(64, "f64", seed=2)
    A = makb200.to_device(A0)
    D, V = makb200.eigh_full(A, fixgauge=False)
    Vn = makb200.to_numpy(V)
    assert np.linalg.norm(A0 @ Vn - Vn * D.cpu().numpy()) < 1e-12
    # eigh_vals! (values-only path): tests/test_gpu_y_vals.py
    D2 = torch.empty(64, dtype=torch.float64, device=A.device)
    V2 = makb200.colmajor_empty(64, 64, torch.float64, A.device)
    o = makb200.eigh_full_(makb200.to_device(A0), (D2, V2))
    assert o[0] is D2 and o[1] is V2
    # fixed spectrum (test/testsuite/decompositions/eigh.jl:128,167)
    Q, _ = O.qr_compact(O.randn_matrix(4, 4, "f64", 3))
    d = np.array([0.9, 0.3, 0.1, 0.01])
    S = (Q * d) @ Q.T
    S = (S + S.T) / 2
    Dt, Vt, eps = makb200.eigh_trunc(makb200.to_device(S), trunc={"rtol": 0.2, "maxrank": 3})
    np.testing.assert_allclose(np.sort(Dt.cpu().numpy())[::-1], d[:2], rtol=1e-12)
    np.testing.assert_allclose(eps, np.linalg.norm(d[2:]), rtol=1e-10)
    wo, Vo, epso = O.eigh_trunc(S, O.truncation_strategy(rtol=0.2, maxrank=3))
    np.testing.assert_allclose(eps, epso, rtol=1e-10)


@pytest.mark.parametrize("case", ["rand", "wilkinson", "glued", "toeplitz", "zero_e"])
def test_stedc_vs_scipy(case):
    import makb200
    from scipy.linalg import eigh_tridiagonal
    rng = np.random.default_rng(0)
    n = 700
    if case == "rand":
        d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    elif case == "wilkinson":
        n = 601
        d, e = np.abs(np.arange(-300, 301)).astype(float), np.ones(n - 1)
    elif case == "glued":
        n = 640
        d, e = np.tile(np.arange(16.0), n // 16), np.ones(n - 1)
        e[15::16] = 1e-9
    elif case == "toeplitz":
        d, e = np.full(n, 2.0), np.full(n - 1, -1.0)
    else:
        d, e = rng.standard_normal(n), np.zeros(n - 1)
    h = makb200.Handle.get("cuda:0")
    dd, ed = torch.tensor(d, device="cuda"), torch.tensor(e, device="cuda")
    W = torch.empty(n, dtype=torch.float64, device="cuda")
    Z = makb200.colmajor_empty(n, n, torch.float64, "cuda")
    lw = h.lib.makb200_stedc_worksize(h.h, n)
    work = h.workspace(lw)
    rc = h.lib.makb200_stedc(h.h, n, dd.data_ptr(), ed.data_ptr(), W.data_ptr(), Z.data_ptr(), n, work.data_ptr(),
                             work.numel(), ctypes.c_void_p(0))
    assert rc == 0
    torch.cuda.synchronize()
    w, Zn = W.cpu().numpy(), makb200.to_numpy(Z)
    wref = eigh_tridiagonal(d, e, eigvals_only=True)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    nrm = np.abs(wref).max()
    tol = O.tol_for(n)
    assert np.max(np.abs(w - wref)) / nrm <= tol
    assert np.linalg.norm(T @ Zn - Zn * w) / nrm <= tol
    assert O.orth_err(Zn) <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_batched_vs_oracle(dtype):
    """Batched eigh_full! (one CTA per block, two-sided Jacobi): per-block results vs the LAPACK oracle,
    incl. degenerate / zero / identity / +-pair spectra, only the upper triangle read, inputs intact,
    and blocks too large for shared memory routed through the single-matrix path."""
    import makb200
    rng = np.random.default_rng(9)
    ns = [1, 2, 3, 16, 17, 31, 32, 33, 54, 64, 75, 79] + [int(v) for v in rng.integers(16, 79, size=20)]
    ns += [130]  # routed (does not fit one CTA)
    As0 = [O.rand_hermitian(n, dtype, seed=500 + i) for i, n in enumerate(ns)]
    Q, _ = O.qr_compact(O.randn_matrix(40, 40, dtype, seed=3))
    specials = [np.zeros((24, 24)), np.eye(30), (Q * np.repeat([-1.0, 1.0], 20)) @ Q.conj().T,
                (Q * np.repeat(np.arange(4.0), 10)) @ Q.conj().T, (Q * 10.0 ** (-12 * np.arange(40) / 40)) @ Q.conj().T]
    for S in specials:
        S = (S + S.conj().T) / 2
        As0.append(S.astype(np.complex128) if dtype == "c128" else np.ascontiguousarray(S.real))
    As = [makb200.to_device(a) for a in As0]
    # garbage in the strictly lower triangle of one block must be ignored (uplo='U')
    junk = As0[8].copy()
    junk[np.tril_indices(junk.shape[0], -1)] = 7.0
    As[8] = makb200.to_device(junk)
    DVs = makb200.eigh_full_batched_(As, check=False)
    torch.cuda.synchronize()
    for i, (a, (D, V)) in enumerate(zip(As0, DVs)):
        if a.shape[0] <= 79 and i != 8:
            assert np.array_equal(makb200.to_numpy(As[i]), a)   # small blocks are not destroyed
        _check(a, D.cpu().numpy(), makb200.to_numpy(V), vec_cmp=(i < len(ns)))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_batched_pooled_threads(dtype):
    """Blocks beyond the one-CTA limit fan out over the stream pool, one host thread per stream."""
    import makb200
    rng = np.random.default_rng(12)
    ns = [int(v) for v in rng.integers(90, 220, size=20)] + [20, 48]
    As0 = [O.rand_hermitian(n, dtype, seed=800 + i) for i, n in enumerate(ns)]
    DVs = makb200.eigh_full_batched_([makb200.to_device(a) for a in As0], check=False)
    torch.cuda.synchronize()
    for a, (D, V) in zip(As0, DVs):
        _check(a, D.cpu().numpy(), makb200.to_numpy(V), vec_cmp=False)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("lockstep", ["1", "0"])
def test_eigh_batched_lockstep(dtype, lockstep, monkeypatch, capfd):
    """>= 8 blocks beyond the one-CTA limit: after the one-launch tridiagonalisation the eigensolve runs in lock-step
    (ONE batched D&C over the block-diagonal union of the tridiagonals + ONE batched back-transformation,
    csrc/eigh_lockstep.cuh); MAKB200_SVD_LOCKSTEP=0 keeps the per-block chain.  Ragged sizes (different tree depths
    in one pass), odd sizes, and spectra that deflate heavily: identity, zero, +-1 pairs, four clusters, graded."""
    import makb200
    monkeypatch.setenv("MAKB200_SVD_LOCKSTEP", lockstep)
    monkeypatch.setenv("MAKB200_LOCKSTEP_VERBOSE", "1")
    rng = np.random.default_rng(14)
    ns = [int(v) for v in rng.integers(128, 340, size=14)] + [129, 257, 260, 300, 333, 511, 512]
    As0 = [O.rand_hermitian(n, dtype, seed=1300 + i) for i, n in enumerate(ns)]
    nq = 160
    Q, _ = O.qr_compact(O.randn_matrix(nq, nq, dtype, seed=3))
    specials = [np.zeros((150, 150)), np.eye(140), (Q * np.repeat([-1.0, 1.0], nq // 2)) @ Q.conj().T,
                (Q * np.repeat(np.arange(4.0), nq // 4)) @ Q.conj().T, (Q * 10.0 ** (-12 * np.arange(nq) / nq)) @ Q.conj().T,
                1e-100 * O.rand_hermitian(131, dtype, seed=5), 1e100 * O.rand_hermitian(137, dtype, seed=6)]
    for S in specials:
        S = (S + S.conj().T) / 2
        As0.append(np.asfortranarray(S.astype(np.complex128) if dtype == "c128" else np.ascontiguousarray(S.real)))
    DVs = makb200.eigh_full_batched_([makb200.to_device(a) for a in As0], check=False)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert ("solved in lock-step" in err) == (lockstep == "1"), err
    for i, (a, (D, V)) in enumerate(zip(As0, DVs)):
        _check(a, D.cpu().numpy(), makb200.to_numpy(V), vec_cmp=(i < len(ns)))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_persistent_column_kernels_small_n_sweep(dtype, monkeypatch):
    """The persistent TMA column kernels (csrc/trd2.cuh) are the default from n = 1536 up; here the switch is moved to
    n = 2 so that every tile-geometry corner (one tile, partial bands, chunks with no work, odd n -> unaligned lda ->
    round-1 kernels) runs against the oracle.  The geometry itself is replayed on the CPU in tests/test_trd_tiles_cpu.py."""
    monkeypatch.setenv("MAKB200_SYMV_V2_MIN", "2")
    for n in list(range(2, 36, 3)) + [64, 66, 128, 130, 256, 258, 300, 512, 514, 600, 1024, 1026]:
        A = O.rand_hermitian(n, dtype, seed=500 + n)
        w, V = _eigh(A)
        _check(A, w, V, vec_cmp=False)
