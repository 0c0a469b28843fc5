"""qr_compact!/qr_full! on B200 vs the LAPACK-replay oracle.
Tolerances (north_star): ||A-QR||/||A||, ||Q^H Q - I||_F <= 10*n*eps with n = max(m,n); the
gauge-fixed factors are unique for full-rank A, so Q and R are also compared directly with the
oracle's (tolerance 100*n*eps*cond-ish slack stated inline)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu

SIZES = [(54, 37), (54, 54), (54, 63)]  # the reference's sizes, test/decompositions/qr.jl:21-22


def _run(fn_name, A_np, **kw):
    import makb200
    A = makb200.to_device(A_np)
    Q, R = getattr(makb200, fn_name)(A, **kw)
    torch.cuda.synchronize()
    return makb200.to_numpy(Q), makb200.to_numpy(R), A


def _check(A, Q, R, Qo, Ro, full=False):
    m, n = A.shape
    k = min(m, n)
    tol = O.tol_for(m, n)
    assert O.rel_resid(A, Q, R) <= tol
    assert O.orth_err(Q) <= tol
    assert np.array_equal(R, np.triu(R))
    d = np.diagonal(R)[:k]
    assert np.all(d.real >= 0) and np.all(d.imag == 0)
    # direct comparison with the oracle's gauge-fixed factors (first k columns / rows)
    cmp_tol = 200 * max(m, n) * O.EPS * np.linalg.cond(A[:, :k]) ** 0 * 50
    assert np.linalg.norm(Q[:, :k] - Qo[:, :k]) <= cmp_tol * np.sqrt(k)
    assert np.linalg.norm(R[:k] - Ro[:k]) <= cmp_tol * np.linalg.norm(Ro)
    if full and m > k:
        # trailing columns span the orthogonal complement (gauge-free comparison)
        assert np.linalg.norm(Qo[:, :k].conj().T @ Q[:, k:]) <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", SIZES + [(1, 1), (2, 5), (300, 200), (200, 300), (513, 129), (1000, 1000)])
def test_qr_compact_vs_oracle(m, n, dtype):
    A = O.randn_matrix(m, n, dtype, seed=123)
    Q, R, Adev = _run("qr_compact", A)
    assert Q.shape == (m, min(m, n)) and R.shape == (min(m, n), n)
    import makb200
    assert np.array_equal(makb200.to_numpy(Adev), A)  # out-of-place call leaves A untouched
    Qo, Ro = O.qr_compact(A)
    _check(A, Q, R, Qo, Ro)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", SIZES + [(300, 100)])
def test_qr_full_vs_oracle(m, n, dtype):
    A = O.randn_matrix(m, n, dtype, seed=7)
    Q, R, _ = _run("qr_full", A)
    assert Q.shape == (m, m) and R.shape == (m, n)
    Qo, Ro = O.qr_full(A)
    _check(A, Q, R, Qo, Ro, full=True)


def test_qr_inplace_outputs_and_skip_r():
    import makb200
    A0 = O.randn_matrix(54, 37, "c128", seed=3)
    A = makb200.to_device(A0)
    Q = makb200.colmajor_empty(54, 37, torch.complex128, A.device)
    R = makb200.colmajor_empty(37, 37, torch.complex128, A.device)
    Q2, R2 = makb200.qr_compact_(A, (Q, R))
    assert Q2 is Q and R2 is R
    Qn = makb200.to_numpy(Q)
    # R not requested: zero-length R (qr.jl:15,26,149); Q still gauge-fixed
    A = makb200.to_device(A0)
    Rnone = makb200.colmajor_empty(0, 0, torch.complex128, A.device)
    Q3, _ = makb200.qr_compact_(A, (makb200.colmajor_empty(54, 37, torch.complex128, A.device), Rnone))
    assert np.allclose(makb200.to_numpy(Q3), Qn, atol=1e-14)


def test_qr_capability_negatives_and_errors():
    import makb200
    A = makb200.to_device(O.randn_matrix(20, 10, "f64", 1))
    with pytest.raises(ValueError):
        makb200.qr_compact(A, pivoted=True)       # test/testsuite/decompositions/qr.jl:57-77
    with pytest.raises(ValueError):
        makb200.qr_compact(A, blocksize=8)
    with pytest.raises(ValueError):
        makb200.qr_compact(A, alg=makb200.Householder(), positive=True)  # kwargs + instance
    with pytest.raises(ValueError):
        makb200.qr_compact_(A, (makb200.colmajor_empty(20, 9, torch.float64, A.device),
                                makb200.colmajor_empty(10, 10, torch.float64, A.device)))
    with pytest.raises(TypeError):
        makb200.qr_compact(A.to(torch.float32))


def test_qr_special_matrices():
    import makb200
    for A in (np.zeros((9, 5)), np.eye(7, 4), -np.eye(6), np.triu(O.randn_matrix(8, 8, "f64", 2))):
        Q, R, _ = _run("qr_compact", A)
        assert O.orth_err(Q) < 1e-13 and np.linalg.norm(A - Q @ R) < 1e-13
        assert np.all(np.diagonal(R) >= 0)
    # graded spectrum (parity only): sigma_i = 10^(-12 i/n)
    n = 96
    U, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 5))
    V, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 6))
    A = U @ np.diag(10.0 ** (-12 * np.arange(n) / n)) @ V
    Q, R, _ = _run("qr_compact", A)
    assert O.rel_resid(A, Q, R) < O.tol_for(n) and O.orth_err(Q) < O.tol_for(n)


def test_qr_strided_view_input():
    import makb200
    big = makb200.to_device(O.randn_matrix(100, 80, "f64", 9))
    A = big[7:61, 3:40]
    A0 = makb200.to_numpy(A)
    Q, R = makb200.qr_compact(A)
    assert O.rel_resid(A0, makb200.to_numpy(Q), makb200.to_numpy(R)) < O.tol_for(54, 37)
    # in-place on a view with lda != m
    work = makb200.to_device(O.randn_matrix(100, 80, "f64", 9))
    Av = work[7:61, 3:40]
    Qv, Rv = makb200.qr_compact_(Av)
    assert np.allclose(makb200.to_numpy(Qv), makb200.to_numpy(Q), atol=1e-13)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_qr_batched_vs_oracle(dtype):
    import makb200
    rng = np.random.default_rng(4)
    sizes = [(16, 16), (17, 9), (9, 17), (32, 32), (54, 37), (54, 63), (64, 64), (100, 100), (1, 1), (130, 130),
             (200, 150)]
    sizes += [(int(s), int(s)) for s in rng.integers(16, 96, size=40)]
    As0 = [O.randn_matrix(m, n, dtype, seed=100 + i) for i, (m, n) in enumerate(sizes)]
    As = [makb200.to_device(a) for a in As0]
    QRs = makb200.qr_compact_batched_(As)
    torch.cuda.synchronize()
    for a, (Q, R) in zip(As0, QRs):
        Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
        Qo, Ro = O.qr_compact(a)
        _check(a, Qn, Rn, Qo, Ro)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_qr_batched_blocked_lockstep(dtype):
    """Mid-size ragged blocks (lock-step blocked path of batched_blocked.cu): tall, wide, square, widths
    that are not multiples of the column step, graded columns, a zero block, and a strided view."""
    import makb200
    sizes = [(96, 96), (97, 101), (130, 130), (200, 150), (150, 200), (257, 257), (300, 100), (100, 300),
             (512, 512), (333, 129), (128, 128), (99, 160)]
    As0 = [O.randn_matrix(m, n, dtype, seed=300 + i) for i, (m, n) in enumerate(sizes)]
    As0[2] = As0[2] * (10.0 ** (-12 * np.arange(130) / 130))[None, :]   # graded columns
    As0[4] = np.zeros_like(As0[4])                                      # zero block
    As = [makb200.to_device(a) for a in As0]
    # strided view: lda != m
    big = makb200.to_device(O.randn_matrix(260, 140, dtype, seed=399))
    As.append(big[10:210, 5:135]); As0.append(makb200.to_numpy(big)[10:210, 5:135].copy())
    QRs = makb200.qr_compact_batched_(As)
    torch.cuda.synchronize()
    for idx, (a, (Q, R)) in enumerate(zip(As0, QRs)):
        Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
        m, n = a.shape
        tol = O.tol_for(m, n)
        assert np.allclose(np.tril(Rn, -1), 0)
        assert np.all(np.real(np.diagonal(Rn)) >= 0) and np.allclose(np.imag(np.diagonal(Rn)), 0)
        if idx == 4:
            assert np.allclose(Rn, 0) and O.orth_err(Qn) <= tol
            continue
        assert O.rel_resid(a, Qn, Rn) <= tol
        assert O.orth_err(Qn) <= tol
        if idx != 2 and m >= n:
            Qo, Ro = O.qr_compact(a)
            assert np.linalg.norm(Rn - Ro) <= 100 * tol * np.linalg.norm(Ro)
            assert np.linalg.norm(Qn - Qo) <= 100 * tol * np.sqrt(n)
