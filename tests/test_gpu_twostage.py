"""EXPERIMENTAL two-stage eigh path (MAKB200_EIGH_TWOSTAGE, read once per process -> run in a
subprocess): stage 1 (dense -> band), the diamond-blocked Q2 application, and the assembled
eigh_full! against numpy, tolerance 10*n*eps."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import makb200
b = 16
out = []
for dtype in ("f64", "c128"):
    for n in (40, 150, 257):
        rng = np.random.default_rng(n)
        G = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if dtype == "c128" else 0)
        A0 = np.asfortranarray((G + G.conj().T) / 2)
        w0 = np.linalg.eigvalsh(A0)
        A = makb200.to_device(A0)
        makb200.sy2sb_(A, b)
        An = makb200.to_numpy(A)
        i, j = np.indices((n, n))
        B = np.where((i - j >= 0) & (i - j <= b), An, 0)
        B = B + np.tril(B, -1).conj().T
        B[np.diag_indices(n)] = B.diagonal().real
        band_err = float(np.abs(np.linalg.eigvalsh(B) - w0).max() / np.abs(w0).max())
        D, V = makb200.eigh_full(makb200.to_device(A0))
        torch.cuda.synchronize()
        w = (torch.diagonal(D) if D.dim() == 2 else D).cpu().numpy().real
        Vn = makb200.to_numpy(V)
        out.append(dict(dtype=dtype, n=n, band=band_err, vals=float(np.abs(w - w0).max() / np.abs(w0).max()),
                        resid=float(np.linalg.norm(A0 @ Vn - Vn * w) / np.abs(w0).max()),
                        orth=float(np.linalg.norm(Vn.conj().T @ Vn - np.eye(n))),
                        gauge=bool(np.all(np.abs(Vn[np.abs(Vn).argmax(axis=0), np.arange(n)].imag) < 1e-14))))
print("RESULT " + json.dumps(out))
"""


def test_two_stage_eigh_subprocess():
    env = dict(os.environ, MAKB200_EIGH_TWOSTAGE="16")
    p = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for r in json.loads(line[len("RESULT "):]):
        n = r["n"]
        tol = 10 * n * 2.220446049250313e-16
        assert r["band"] <= tol, r
        assert r["vals"] <= tol, r
        assert r["resid"] <= tol * n ** 0.5, r
        assert r["orth"] <= tol * n ** 0.5, r
        assert r["gauge"], r
