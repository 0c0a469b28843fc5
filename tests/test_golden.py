"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
  - CPU: the oracle still reproduces them (guards the checker against drift of scipy/OpenBLAS);
  - GPU: the CUDA path matches them through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import mak_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "golden_*.npz")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    A = g["A"]
    Q, R = O.qr_compact(A)
    assert np.linalg.norm(Q - g["Q"]) < 1e-12 and np.linalg.norm(R - g["R"]) < 1e-12
    U, S, Vh = O.svd_compact(A)
    assert np.max(np.abs(S - g["S"])) < 1e-12
    if "w" in g:
        w, V = O.eigh_full(g["H"])
        assert np.max(np.abs(w - g["w"])) < 1e-12


def test_kat():
    g = np.load(os.path.join(HERE, "golden", "kat_eigh3.npz"))
    w, _ = O.eigh_full(g["A"])
    np.testing.assert_allclose(w, g["w"], atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_cuda_matches_golden(path):
    import makb200
    import torch
    g = np.load(path)
    A = g["A"]
    m, n = A.shape
    tol = O.tol_for(m, n)
    Q, R = makb200.qr_compact(makb200.to_device(A))
    assert np.linalg.norm(makb200.to_numpy(Q) - g["Q"]) <= 100 * tol
    assert np.linalg.norm(makb200.to_numpy(R) - g["R"]) <= 100 * tol * np.linalg.norm(g["R"])
    Qf, Rf = makb200.qr_full(makb200.to_device(A))
    k = min(m, n)
    assert np.linalg.norm(makb200.to_numpy(Qf)[:, :k] - g["Qf"][:, :k]) <= 100 * tol
    U, S, Vh = makb200.svd_compact(makb200.to_device(A))
    assert np.max(np.abs(S.cpu().numpy() - g["S"])) / g["S"][0] <= tol
    assert np.linalg.norm(makb200.to_numpy(U) - g["U"]) <= 1e-9
    assert np.linalg.norm(makb200.to_numpy(Vh) - g["Vh"]) <= 1e-9
    if "W" in g:
        W, P = makb200.left_polar(makb200.to_device(A))
        assert np.linalg.norm(makb200.to_numpy(W) - g["W"]) <= 1e-10
        assert np.linalg.norm(makb200.to_numpy(P) - g["P"]) <= 1e-10 * np.linalg.norm(g["P"])
    if "w" in g:
        D, V = makb200.eigh_full(makb200.to_device(g["H"]))
        assert np.max(np.abs(D.cpu().numpy() - g["w"])) / np.abs(g["w"]).max() <= tol
        assert np.linalg.norm(makb200.to_numpy(V) - g["V"]) <= 1e-9
    torch.cuda.synchronize()
