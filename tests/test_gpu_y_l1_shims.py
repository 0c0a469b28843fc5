"""L1 (LAPACK-shaped) shims and the callers that go through them, on the B200 vs the oracle:
geqrf!/ungqr!/unmqr!(::B200) (ext pattern MatrixAlgebraKitCUDAExt.jl:32-34; consumers qr.jl:160-175, 236-262),
svd_full! (svd.jl:202-212, job 'A'), and the LAPACK-named algorithm tags with driver = B200()
(BASELINE: LAPACK_DivideAndConquer, LAPACK_MultipleRelativelyRobustRepresentations).  Tolerance 10 n eps."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (37, 54), (300, 200), (513, 129), (700, 700)])
def test_geqrf_ungqr_vs_oracle_blocksize1(m, n, dtype):
    """makb200_geqrf + makb200_orgqr = the reference's `blocksize = 1` branch (geqrf! + ungqr!, qr.jl:160-175)."""
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=21)
    A = makb200.to_device(A0)
    A, tau = makb200.geqrf_(A)
    k = min(m, n)
    Q = makb200.ungqr_(A, tau, makb200.colmajor_empty(m, k, A.dtype, A.device))
    torch.cuda.synchronize()
    Qn = makb200.to_numpy(Q)
    Rn = np.triu(makb200.to_numpy(A)[:k, :])
    Qo, Ro = O.qr_householder(A0, mode="compact", blocksize=1)
    tol = O.tol_for(m, n)
    assert O.rel_resid(A0, Qn, Rn) <= tol and O.orth_err(Qn) <= tol
    d = np.diagonal(Rn)
    assert np.all(d.real >= 0) and np.all(np.abs(d.imag) == 0)          # non-negative-beta reflectors: gauge for free
    assert np.linalg.norm(Rn - Ro) <= 500 * tol * np.linalg.norm(Ro)
    assert np.linalg.norm(Qn - Qo) <= 500 * tol * np.sqrt(k)
    # full Q (ncols = m) from the same reflectors
    if m > k:
        Qf = makb200.to_numpy(makb200.ungqr_(A, tau, makb200.colmajor_empty(m, m, A.dtype, A.device)))
        assert O.orth_err(Qf) <= tol and np.linalg.norm(Qf[:, :k] - Qn) <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,nc", [(54, 37, 20), (300, 200, 77), (400, 400, 150), (200, 300, 64)])
def test_unmqr_left_vs_explicit_q(m, n, nc, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=22)
    C0 = O.randn_matrix(m, nc, dtype, seed=23)
    A, tau = makb200.geqrf_(makb200.to_device(A0))
    Qf = makb200.to_numpy(makb200.ungqr_(A, tau, makb200.colmajor_empty(m, m, A.dtype, A.device)))
    tol = O.tol_for(m, max(n, nc))
    for trans in ("N", "C"):
        C = makb200.unmqr_("L", trans, A, tau, makb200.to_device(C0))
        torch.cuda.synchronize()
        ref = (Qf if trans == "N" else Qf.conj().T) @ C0
        assert np.linalg.norm(makb200.to_numpy(C) - ref) <= tol * np.linalg.norm(C0)
    with pytest.raises(ValueError):
        makb200.unmqr_("R", "N", A, tau, makb200.to_device(C0))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (300, 100), (64, 1), (200, 199)])
def test_qr_null_by_geqrf_unmqr(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=24)
    N = makb200.to_numpy(makb200.qr_null(makb200.to_device(A0)))
    tol = O.tol_for(m, n)
    assert N.shape == (m, m - n)
    assert O.orth_err(N) <= tol
    assert np.linalg.norm(A0.conj().T @ N) <= tol * np.linalg.norm(A0)
    Qo, _ = O.qr_full(A0)
    No = Qo[:, n:]
    assert np.linalg.norm(N @ N.conj().T - No @ No.conj().T) <= 100 * tol     # same subspace as the reference's N


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (37, 54), (300, 120), (120, 300), (1, 5), (5, 1)])
def test_svd_full_vs_oracle(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=25)
    U, S, Vh = makb200.svd_full(makb200.to_device(A0))
    torch.cuda.synchronize()
    Un, Sn, Vhn = makb200.to_numpy(U), makb200.to_numpy(S), makb200.to_numpy(Vh)
    Uo, So, Vho = O.svd_full(A0)
    k = min(m, n)
    tol = O.tol_for(m, n)
    assert Un.shape == (m, m) and Sn.shape == (m, n) and Vhn.shape == (n, n)
    assert np.array_equal(Sn, np.diag(np.diagonal(Sn)) if m == n else Sn * np.eye(m, n))
    assert np.max(np.abs(np.diagonal(Sn) - np.diagonal(So))) / So[0, 0] <= tol
    assert np.linalg.norm(A0 - Un @ Sn @ Vhn) / np.linalg.norm(A0) <= tol
    assert O.orth_err(Un) <= tol and O.orth_err(Vhn, "right") <= tol
    # leading triplets: the compact decomposition (gauge-fixed vectors, compared scaled by the gap)
    sv = np.diagonal(So)
    gap = np.minimum(np.abs(np.diff(sv, prepend=np.inf)), np.abs(np.diff(sv, append=-np.inf)))
    gap = np.minimum(gap, sv) if m != n else gap
    err = np.maximum(np.linalg.norm(Un[:, :k] - Uo[:, :k], axis=0), np.linalg.norm(Vhn[:k] - Vho[:k], axis=1))
    assert np.all(err <= 200 * max(m, n) * O.EPS * sv[0] / gap + 1e-12)
    # extra columns / rows: same subspace as LAPACK's, own gauge (entry of maximal modulus real positive)
    if m > k:
        E, Eo = Un[:, k:], Uo[:, k:]
        assert np.linalg.norm(E @ E.conj().T - Eo @ Eo.conj().T) <= 100 * tol
        piv = O._argmaxabs_cols(E)
        assert np.all(piv.real > 0) and np.all(np.abs(piv.imag) <= 1e-15)
    if n > k:
        E, Eo = Vhn[k:], Vho[k:]
        assert np.linalg.norm(E.conj().T @ E - Eo.conj().T @ Eo) <= 100 * tol
        piv = O._argmaxabs_cols(E.T)
        assert np.all(piv.real > 0) and np.all(np.abs(piv.imag) <= 1e-15)


def test_svd_full_preallocated_and_empty():
    import makb200
    A0 = O.randn_matrix(20, 12, "c128", seed=26)
    outs = makb200.svd.initialize_output_full(makb200.to_device(A0))
    U, S, Vh = makb200.svd_full_(makb200.to_device(A0), outs)
    assert U is outs[0] and S is outs[1] and Vh is outs[2]
    U, S, Vh = makb200.svd_full(makb200.to_device(np.zeros((0, 4))))
    assert tuple(U.shape) == (0, 0) and tuple(S.shape) == (0, 4) and np.allclose(makb200.to_numpy(Vh), np.eye(4))


def test_lapack_named_tags_with_the_b200_driver():
    """DivideAndConquer(driver = B200()) on svd and RobustRepresentations(driver = B200()) on eigh - what the
    BASELINE-named aliases expand to - run the B200 kernels; tags the driver does not provide throw."""
    import makb200
    A0 = O.randn_matrix(60, 40, "f64", seed=27)
    So = O.svd_vals(A0)
    for alg in (makb200.DivideAndConquer(driver=makb200.B200()), makb200.SafeDivideAndConquer(driver=makb200.B200()),
                "DivideAndConquer", makb200.B200_SVDViaPolar()):
        U, S, Vh = makb200.svd_compact(makb200.to_device(A0), alg=alg)
        assert np.max(np.abs(S.cpu().numpy() - So)) <= O.tol_for(60, 40) * So[0]
    for alg in ("QRIteration", "Jacobi"):
        with pytest.raises(ValueError):
            makb200.svd_compact(makb200.to_device(A0), alg=alg)
    H0 = O.rand_hermitian(50, "c128", seed=28)
    wo = O.eigh_vals(H0)
    for alg in (makb200.RobustRepresentations(driver=makb200.B200()), "RobustRepresentations", makb200.B200_DivideAndConquer()):
        D, V = makb200.eigh_full(makb200.to_device(H0), alg=alg)
        assert np.max(np.abs(D.cpu().numpy() - wo)) <= O.tol_for(50) * np.abs(wo).max()
    for alg in ("QRIteration", "Jacobi"):
        with pytest.raises(ValueError):
            makb200.eigh_full(makb200.to_device(H0), alg=alg)
