"""TSQR local step (CholeskyQR2 on the DMMA GEMM) vs the oracle's qr_compact on the same matrix;
tolerance 10*n*eps (n = max(m, n) as in north_star; here bounded by m)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(5000, 256), (100000, 64), (3000, 300), (257, 129)])
def test_tsqr_single_rank_vs_oracle(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=5)
    Q, R = makb200.tsqr_(makb200.to_device(A0))
    torch.cuda.synchronize()
    Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
    Qo, Ro = O.qr_compact(A0)
    tol = O.tol_for(m, n)
    assert O.rel_resid(A0, Qn, Rn) <= tol
    assert O.orth_err(Qn) <= tol
    assert np.array_equal(Rn, np.triu(Rn)) and np.all(np.diagonal(Rn).real > 0)
    assert np.linalg.norm(Rn - Ro) <= 100 * tol * np.linalg.norm(Ro)
    assert np.linalg.norm(Qn - Qo) <= 100 * tol


def test_tsqr_reports_breakdown():
    import makb200
    m, n = 2000, 40
    U, _ = O.qr_compact(O.randn_matrix(m, n, "f64", 1))
    V, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 2))
    A0 = (U * 10.0 ** (-14 * np.arange(n) / n)) @ V   # kappa = 1e14: outside CholeskyQR2's range
    with pytest.raises(makb200.MakError):
        makb200.tsqr_(makb200.to_device(A0))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("kappa_exp,robust", [(9, 1), (13, True), (0, True)])
def test_tsqr_robust_shifted_cholqr(kappa_exp, robust, dtype):
    """Ill-conditioned shards: the shifted-CholeskyQR local step (robust=...) must give the LAPACK
    Householder factorization (oracle) where plain CholeskyQR2 breaks down; residual and
    orthogonality at the usual 10*n*eps.  R is compared column-scaled (graded columns)."""
    import makb200
    m, n = 4000, 48
    U, _ = O.qr_compact(O.randn_matrix(m, n, dtype, 1))
    V, _ = O.qr_compact(O.randn_matrix(n, n, dtype, 2))
    A0 = np.asfortranarray((U * 10.0 ** (-kappa_exp * np.arange(n) / n)) @ V.conj().T)
    Q, R = makb200.tsqr_(makb200.to_device(A0), robust=robust)
    torch.cuda.synchronize()
    Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
    tol = O.tol_for(m, n)
    assert O.rel_resid(A0, Qn, Rn) <= tol
    assert O.orth_err(Qn) <= tol
    assert np.array_equal(Rn, np.triu(Rn)) and np.all(np.diagonal(Rn).real > 0)
    if kappa_exp == 0:
        Qo, Ro = O.qr_compact(A0)
        assert np.linalg.norm(Rn - Ro) <= 100 * tol * np.linalg.norm(Ro)
        assert np.linalg.norm(Qn - Qo) <= 100 * tol
