"""Values-only paths (eigh_vals! eigh.jl:157-161, svd_vals! svd.jl:214-219; LAPACK job 'N'): after the
tridiagonalisation the eigenvalues come from the Sturm K-section kernel (csrc/sturm_core.h) instead of
the D&C solver + back-transformation.  Checked against the LAPACK oracle and the full decomposition."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 54, 130, 257, 600])
def test_eigh_vals_vs_oracle(n, dtype):
    import makb200
    A0 = O.rand_hermitian(n, dtype, seed=100 + n)
    D = makb200.eigh_vals(makb200.to_device(A0))
    assert D.dtype == torch.float64 and tuple(D.shape) == (n,)
    w = D.cpu().numpy()
    wref = O.eigh_vals(A0)
    assert np.all(np.diff(w) >= 0)                                   # ascending (eigh.jl:157-161)
    assert np.max(np.abs(w - wref)) / np.abs(wref).max() <= 10 * n * EPS


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_vals_matches_full_and_keeps_contract(dtype):
    import makb200
    n = 200
    A0 = O.rand_hermitian(n, dtype, seed=7)
    D, _ = makb200.eigh_full(makb200.to_device(A0))
    Dv = torch.empty(n, dtype=torch.float64, device="cuda:0")
    out = makb200.eigh_vals_(makb200.to_device(A0), Dv)
    assert out is Dv                                                 # same output object
    assert float((Dv - D).abs().max()) <= 10 * n * EPS * float(D.abs().max())
    # strided view (lda != n)
    big = makb200.colmajor_zeros(n + 9, n, makb200.to_device(A0).dtype, "cuda:0")
    big[:n, :] = makb200.to_device(A0)
    Ds = makb200.eigh_vals_(big[:n, :])
    assert float((Ds - D).abs().max()) <= 10 * n * EPS * float(D.abs().max())
    with pytest.raises(makb200.DomainError):
        makb200.eigh_vals(makb200.to_device(O.randn_matrix(n, n, dtype, 3)))


def test_eigh_vals_special_spectra():
    import makb200
    # doctest KAT (docs/src/user_interface/truncations.md:19-21)
    A = np.array([[2.0, 1, 0], [1, 3, 1], [0, 1, 4]], order="F")
    w = makb200.eigh_vals(makb200.to_device(A)).cpu().numpy()
    np.testing.assert_allclose(w, [3 - np.sqrt(3), 3, 3 + np.sqrt(3)], rtol=0, atol=1e-14)
    # fixed spectrum between random unitaries (test/testsuite/decompositions/eigh.jl:128,167)
    Q, _ = O.qr_compact(O.randn_matrix(4, 4, "c128", 3))
    d = np.array([0.01, 0.1, 0.3, 0.9])
    H = (Q * d) @ Q.conj().T
    H = (H + H.conj().T) / 2
    np.testing.assert_allclose(makb200.eigh_vals(makb200.to_device(H)).cpu().numpy(), d, rtol=0, atol=1e-14)
    # zero, identity, diagonal with repeated entries, graded
    n = 70
    for M in (np.zeros((n, n)), np.eye(n), np.diag(np.repeat([3.0, -1.0], n // 2)), np.diag(10.0 ** (-np.arange(n) / 5.0))):
        w = makb200.eigh_vals(makb200.to_device(np.asfortranarray(M))).cpu().numpy()
        ref = np.sort(np.diag(M))
        assert np.all(np.diff(w) >= 0)
        assert np.max(np.abs(w - ref)) <= 10 * n * EPS * max(np.abs(ref).max(), 1e-300)
    # huge / tiny scales do not overflow the Sturm recurrence (the kernel scales by max|T|)
    for sc in (1e100, 1e-100):
        A0 = O.rand_hermitian(90, "f64", seed=11) * sc
        w = makb200.eigh_vals(makb200.to_device(A0)).cpu().numpy()
        wref = O.eigh_vals(A0)
        assert np.all(np.isfinite(w)) and np.max(np.abs(w - wref)) / np.abs(wref).max() <= 10 * 90 * EPS


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (37, 54), (300, 120), (3, 2), (2, 5), (1, 1)])
def test_svd_vals_vs_oracle(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=m * 7 + n)
    S = makb200.svd_vals(makb200.to_device(A0))
    s = S.cpu().numpy()
    sref = O.svd_vals(A0)
    assert s.shape == (min(m, n),) and np.all(s >= 0) and np.all(np.diff(s) <= 0)   # descending, non-negative
    assert np.max(np.abs(s - sref)) / sref[0] <= 10 * max(m, n) * EPS
    _, Sc, _ = makb200.svd_compact(makb200.to_device(A0))
    assert float((Sc - S).abs().max()) <= 10 * max(m, n) * EPS * sref[0]


def test_vals_empty():
    import makb200
    assert makb200.eigh_vals(makb200.colmajor_zeros(0, 0, torch.float64, "cuda:0")).numel() == 0
    assert makb200.svd_vals(makb200.colmajor_zeros(0, 5, torch.float64, "cuda:0")).numel() == 0


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_vals_batched(dtype):
    """Batched eigh_vals! / svd_vals! (V = NULL, U = Vh = NULL through the batched C entry points): small blocks in
    the one-CTA Jacobi kernels, larger ones through the single-matrix values-only path."""
    import makb200
    ns = [1, 8, 24, 40, 64, 90, 150, 260]
    Hs0 = [O.rand_hermitian(n, dtype, seed=40 + n) for n in ns]
    Ds = makb200.eigh_vals_batched_([makb200.to_device(a) for a in Hs0])
    torch.cuda.synchronize()
    for a, D, n in zip(Hs0, Ds, ns):
        wref = O.eigh_vals(a)
        assert np.max(np.abs(D.cpu().numpy() - wref)) / np.abs(wref).max() <= 10 * n * EPS
    with pytest.raises(makb200.DomainError):
        makb200.eigh_vals_batched_([makb200.to_device(O.randn_matrix(20, 20, dtype, 1))])
    shapes = [(1, 1), (16, 16), (17, 9), (9, 17), (54, 37), (64, 64), (100, 100), (180, 120), (120, 180)]
    As0 = [O.randn_matrix(m, k, dtype, seed=60 + i) for i, (m, k) in enumerate(shapes)]
    Ss = makb200.svd_vals_batched_([makb200.to_device(a) for a in As0])
    torch.cuda.synchronize()
    for a, S in zip(As0, Ss):
        sref = O.svd_vals(a)
        s = S.cpu().numpy()
        assert s.shape == sref.shape and np.all(np.diff(s) <= 0)
        assert np.max(np.abs(s - sref)) / sref[0] <= 10 * max(a.shape) * EPS


def test_svd_vals_graded_spectrum():
    """Singular values spanning 12 decades between random unitaries: absolute accuracy eps * sigma_1."""
    import makb200
    n = 128
    Uq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 5))
    Vq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 6))
    sv = 10.0 ** (-12 * np.arange(n) / n)
    A = (Uq * sv) @ Vq
    Sv = makb200.svd_vals(makb200.to_device(A)).cpu().numpy()
    assert np.max(np.abs(Sv - sv)) / sv[0] <= O.tol_for(n)
