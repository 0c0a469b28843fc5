"""SURVEY §8(f) rank-1 rows on B200: qr_null!, the LQ family through lq_via_qr!, orth/null routers —
checked against the oracle (LQ of A = adjoint of the oracle QR of A^H) with the 10*n*eps tolerance."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (37, 54), (200, 310)])
def test_lq_and_null(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=31 + m)
    tol = O.tol_for(m, n)
    k = min(m, n)
    L, Q = makb200.lq_compact(makb200.to_device(A0))
    Ln, Qn = makb200.to_numpy(L), makb200.to_numpy(Q)
    Qo, Ro = O.qr_compact(A0.conj().T)
    assert Ln.shape == (m, k) and Qn.shape == (k, n)
    assert O.rel_resid(A0, Ln, Qn) <= tol and O.orth_err(Qn, "right") <= tol
    assert np.array_equal(Ln, np.tril(Ln)) and np.all(np.diagonal(Ln).real >= 0)
    assert np.linalg.norm(Ln - Ro.conj().T) <= 100 * tol * np.linalg.norm(Ro)
    assert np.linalg.norm(Qn - Qo.conj().T) <= 100 * tol
    Lf, Qf = makb200.lq_full(makb200.to_device(A0))
    Lf, Qf = makb200.to_numpy(Lf), makb200.to_numpy(Qf)
    assert Qf.shape == (n, n) and O.rel_resid(A0, Lf, Qf) <= tol and O.orth_err(Qf) <= tol
    # null spaces
    N = makb200.to_numpy(makb200.qr_null(makb200.to_device(A0)))
    assert N.shape == (m, m - k)
    if m > k:
        assert np.linalg.norm(N.conj().T @ A0) <= tol * np.linalg.norm(A0) and O.orth_err(N) <= tol
    Nh = makb200.to_numpy(makb200.lq_null(makb200.to_device(A0)))
    assert Nh.shape == (n - k, n)
    if n > k:
        assert np.linalg.norm(A0 @ Nh.conj().T) <= tol * np.linalg.norm(A0) and O.orth_err(Nh, "right") <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_orth_routers(dtype):
    import makb200
    A0 = O.randn_matrix(60, 40, dtype, seed=77)
    tol = O.tol_for(60, 40)
    for kind in ("qr", "polar", "svd"):
        V, Cm = makb200.left_orth(makb200.to_device(A0), kind=kind)
        V, Cm = makb200.to_numpy(V), makb200.to_numpy(Cm)
        assert O.rel_resid(A0, V, Cm) <= tol and O.orth_err(V) <= tol
    for kind in ("lq", "svd"):
        Cm, Vh = makb200.right_orth(makb200.to_device(A0.conj().T.copy(order="F")), kind=kind)
        Cm, Vh = makb200.to_numpy(Cm), makb200.to_numpy(Vh)
        assert O.rel_resid(A0.conj().T, Cm, Vh) <= tol and O.orth_err(Vh, "right") <= tol
    V, Cm = makb200.left_orth(makb200.to_device(A0), kind="svd", trunc=makb200.truncrank(10))
    assert tuple(V.shape) == (60, 10) and tuple(Cm.shape) == (10, 40)
    N = makb200.to_numpy(makb200.left_null(makb200.to_device(A0)))
    assert N.shape == (60, 20) and np.linalg.norm(N.conj().T @ A0) <= tol * np.linalg.norm(A0)
    Nh = makb200.to_numpy(makb200.right_null(makb200.to_device(A0.conj().T.copy(order="F"))))
    assert Nh.shape == (20, 60) and np.linalg.norm(A0.conj().T @ Nh.conj().T) <= tol * np.linalg.norm(A0)
