"""EXPERIMENTAL second stage of the two-stage tridiagonalisation on the GPU (csrc/sbr.cu): the
tridiagonal must have the band matrix's eigenvalues, and X = Q2 Z (reflectors applied on the host in
reverse generation order, layout of csrc/sbr_core.h) must be its eigenvectors; tolerance 10*n*eps."""
import numpy as np
import pytest
import torch
from scipy.linalg import eigh_tridiagonal

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _band(n, b, dtype, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    if dtype == "c128":
        A = A + 1j * rng.standard_normal((n, n))
    A = (A + A.conj().T) / 2
    i, j = np.indices((n, n))
    A[np.abs(i - j) > b] = 0
    return np.asfortranarray(A)


def _apply_q2(V2, tau2, n, b, Z):
    X = Z.astype(V2.dtype).copy()
    for s in range(n - 2, -1, -1):
        nt = (n - 1 - s + b - 1) // b
        for k in range(nt - 1, -1, -1):
            r0 = s + 1 + k * b
            L = min(b, n - r0)
            v, tau = V2[r0:r0 + L, s], tau2[k, s]
            X[r0:r0 + L] -= np.outer(v, tau * (v.conj() @ X[r0:r0 + L]))
    return X


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,b", [(1, 1), (2, 1), (5, 8), (17, 4), (64, 16), (100, 7), (200, 16), (300, 64), (257, 32)])
def test_sbr_chase_vs_numpy(n, b, dtype):
    import makb200
    A = _band(n, b, dtype, seed=n * 131 + b)
    d, e, V2, tau2 = makb200.sbr_chase_(makb200.to_device(A), b)
    torch.cuda.synchronize()
    d, e = d.cpu().numpy(), e.cpu().numpy()
    V2, tau2 = makb200.to_numpy(V2), makb200.to_numpy(tau2)
    tol = 10 * max(n, 2) * EPS
    wref = np.linalg.eigvalsh(A)
    nrm = max(np.abs(wref).max(), 1e-300)
    w, Z = eigh_tridiagonal(d, e) if n > 1 else (d.copy(), np.ones((1, 1)))
    assert np.max(np.abs(w - wref)) / nrm <= tol
    X = _apply_q2(V2, tau2, n, b, Z)
    assert np.linalg.norm(A @ X - X * w) / nrm <= tol * np.sqrt(n)
    assert np.linalg.norm(X.conj().T @ X - np.eye(n)) <= tol * np.sqrt(n)
