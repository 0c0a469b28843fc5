"""Round-2 BRING-UP tests: kernels written after the round-1 GPU budget was spent.  Their logic is
validated on the CPU emulator (tests/test_emu_kernels_cpu.py, all CTAs as co-resident fibers) but they
have not run on a B200 yet, so they are opt-in at run time (MAKB200_CHASE_PERSISTENT, MAKB200_Q2_FUSED)
and these tests only run with MAKB200_BRINGUP=1 (tools/jobs_r2/job_r2a.sh sets it).  Sorted last on purpose."""
import os

import numpy as np
import pytest
import torch
from scipy.linalg import eigh_tridiagonal

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MAKB200_BRINGUP") != "1",
                                 reason="bring-up kernels (CPU-emulator validated only); set MAKB200_BRINGUP=1")]
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


@pytest.fixture
def persistent_chase():
    os.environ["MAKB200_CHASE_PERSISTENT"] = "1"
    yield
    os.environ.pop("MAKB200_CHASE_PERSISTENT", None)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n,b", [(2, 1), (5, 8), (17, 4), (64, 16), (100, 7), (300, 64), (257, 32), (1500, 32), (2048, 64)])
def test_persistent_chase_matches_wave_kernel(persistent_chase, n, b, dtype):
    """Same reflectors and tridiagonal as the per-wavefront kernel (to rounding), and
    the band matrix's eigenvalues within 10 n eps."""
    import makb200
    A = _band(n, b, dtype, seed=n * 131 + b)
    d, e, V2, tau2 = makb200.sbr_chase_(makb200.to_device(A), b)
    torch.cuda.synchronize()
    os.environ.pop("MAKB200_CHASE_PERSISTENT", None)
    d0, e0, V0, t0 = makb200.sbr_chase_(makb200.to_device(A), b)
    torch.cuda.synchronize()
    # same task arithmetic, but two separately compiled kernels: compare to rounding, not bitwise
    scale = float(d0.abs().max())
    assert float((d - d0).abs().max()) <= 1e-11 * scale and float((e - e0).abs().max()) <= 1e-11 * scale
    assert float((V2 - V0).abs().max()) <= 1e-9 and float((tau2 - t0).abs().max()) <= 1e-9
    w = eigh_tridiagonal(d.cpu().numpy(), e.cpu().numpy(), eigvals_only=True)
    wref = np.linalg.eigvalsh(A)
    assert np.abs(w - wref).max() / np.abs(wref).max() <= 10 * n * EPS


def _apply_q2_host(V2, tau2, n, b, Z):
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
@pytest.mark.parametrize("n,b,g,ncols,cw", [(50, 8, 8, 40, 32), (200, 16, 16, 200, 64), (300, 32, 16, 77, 64),
                                            (257, 32, 32, 257, 64), (400, 64, 32, 130, 32)])
def test_fused_q2_slab_vs_host_reflectors(n, b, g, ncols, cw, dtype):
    """Z <- Q2 Z through the slab kernel against one reflector at a time on the host."""
    import makb200
    if dtype == "c128" and (b, g) == (64, 32) and cw == 64:
        pytest.skip("shared memory")
    A = _band(n, b, dtype, seed=n + b)
    d, e, V2, tau2 = makb200.sbr_chase_(makb200.to_device(A), b)
    rng = np.random.default_rng(n)
    Z0 = rng.standard_normal((n, ncols)) + (1j * rng.standard_normal((n, ncols)) if dtype == "c128" else 0)
    Z0 = np.asfortranarray(Z0)
    os.environ["MAKB200_Q2_FUSED"] = "1"
    os.environ["MAKB200_Q2_CW"] = str(cw)
    try:
        Z = makb200.sbr_apply_q2_(V2, tau2, b, makb200.to_device(Z0), g=g)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MAKB200_Q2_FUSED", None)
        os.environ.pop("MAKB200_Q2_CW", None)
    Xref = _apply_q2_host(makb200.to_numpy(V2), makb200.to_numpy(tau2), n, b, Z0)
    assert np.abs(makb200.to_numpy(Z) - Xref).max() <= 10 * n * EPS * np.abs(Xref).max()


def test_two_stage_eigh_with_bringup_kernels():
    """Assembled two-stage eigh_full! with the persistent chase and the fused Q2 kernel (env read per process)."""
    import subprocess
    import sys
    import json
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_twostage", os.path.join(here, "test_gpu_twostage.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    SCRIPT, ROOT = mod.SCRIPT, mod.ROOT
    env = dict(os.environ, MAKB200_EIGH_TWOSTAGE="16", MAKB200_CHASE_PERSISTENT="1", MAKB200_Q2_FUSED="1")
    p = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for r in json.loads(line[len("RESULT "):]):
        tol = 10 * r["n"] * EPS
        assert r["vals"] <= tol and r["resid"] <= tol * r["n"] ** 0.5 and r["orth"] <= tol * r["n"] ** 0.5 and r["gauge"], r


@pytest.fixture
def bhetrd():
    os.environ["MAKB200_BHETRD"] = "1"
    yield
    os.environ.pop("MAKB200_BHETRD", None)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_batched_eigh_with_one_launch_tridiagonalisation(bhetrd, dtype):
    """eigh_full! of mid-size blocks (beyond the one-CTA Jacobi kernel) with MAKB200_BHETRD=1: every such block
    is tridiagonalised by csrc/bhetrd.cuh in ONE launch, then stedc + back-transformation per block."""
    import makb200
    from oracle import mak_oracle as O
    ns = [20, 70, 90, 129, 200, 257, 300, 33, 512]
    As0 = [O.rand_hermitian(n, dtype, seed=600 + n) for n in ns]
    # only the upper triangle may be read (uplo = 'U'): poison the strict lower part of one block
    As0p = [a.copy() for a in As0]
    As0p[3][np.tril_indices(ns[3], -1)] = 1e3
    outs = makb200.eigh_full_batched_([makb200.to_device(a) for a in As0p], check=False)
    torch.cuda.synchronize()
    for a, (D, V), n in zip(As0, outs, ns):
        w, Vn = D.cpu().numpy(), makb200.to_numpy(V)
        wref = O.eigh_vals(a)
        tol = 10 * n * EPS
        assert np.max(np.abs(w - wref)) / np.abs(wref).max() <= tol
        assert np.linalg.norm(a @ Vn - Vn * w) / np.linalg.norm(a) <= tol
        assert np.linalg.norm(Vn.conj().T @ Vn - np.eye(n)) <= tol
        piv = Vn[np.argmax(np.abs(Vn), axis=0), np.arange(n)]
        assert np.all(np.abs(piv.imag) <= 1e-13) and np.all(piv.real > 0)       # gauge (common/gauge.jl:38-45)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_single_launch_tridiagonalisation_inside_eigh_and_pooled_svd(dtype):
    """MAKB200_BHETRD=2: eigh_t replaces its 2n launches by ONE single-CTA launch (n <= 512), which is what the
    pooled per-block paths of the batched svd / eigh run for mid-size blocks."""
    import makb200
    from oracle import mak_oracle as O
    os.environ["MAKB200_BHETRD"] = "2"
    try:
        for n in (3, 54, 130, 300, 512):
            A0 = O.rand_hermitian(n, dtype, seed=70 + n)
            D, V = makb200.eigh_full(makb200.to_device(A0))
            w, Vn = D.cpu().numpy(), makb200.to_numpy(V)
            wref = O.eigh_vals(A0)
            tol = 10 * n * EPS
            assert np.max(np.abs(w - wref)) / np.abs(wref).max() <= tol
            assert np.linalg.norm(A0 @ Vn - Vn * w) / np.linalg.norm(A0) <= tol
            assert np.linalg.norm(Vn.conj().T @ Vn - np.eye(n)) <= tol
            wv = makb200.eigh_vals(makb200.to_device(A0)).cpu().numpy()
            assert np.max(np.abs(wv - wref)) / np.abs(wref).max() <= tol
        sizes = [(100, 100), (180, 120), (120, 180), (260, 260), (40, 40)]
        As0 = [O.randn_matrix(m, k, dtype, seed=80 + i) for i, (m, k) in enumerate(sizes)]
        outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As0])
        torch.cuda.synchronize()
        for a, (U, S, Vh) in zip(As0, outs):
            Un, Sn, Vn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
            tol = O.tol_for(*a.shape)
            assert np.max(np.abs(Sn - O.svd_vals(a))) / Sn[0] <= tol
            assert O.rel_resid(a, Un * Sn, Vn) <= tol and O.orth_err(Un) <= tol and O.orth_err(Vn, "right") <= tol
    finally:
        os.environ.pop("MAKB200_BHETRD", None)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_panel_blocked_warp_qr(dtype):
    """MAKB200_BQR_WARP_BLK=1: tiny blocks (m, n <= 32) through batched_qr_warp_blk_kernel; same gauge-fixed
    factors as the oracle and as the default warp kernel."""
    import makb200
    from oracle import mak_oracle as O
    rng = np.random.default_rng(5)
    shapes = [(16, 16), (32, 32), (23, 23), (32, 17), (17, 32), (5, 3), (1, 1), (31, 32), (24, 24), (2, 7), (32, 4), (4, 32),
              (9, 9), (28, 27)] + [(int(a), int(b)) for a, b in rng.integers(1, 33, size=(60, 2))]
    As0 = [O.randn_matrix(m, n, dtype, seed=400 + i) for i, (m, n) in enumerate(shapes)]
    os.environ["MAKB200_BQR_WARP_BLK"] = "1"
    try:
        outs = makb200.qr_compact_batched_([makb200.to_device(a) for a in As0])
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MAKB200_BQR_WARP_BLK", None)
    ref = makb200.qr_compact_batched_([makb200.to_device(a) for a in As0])
    torch.cuda.synchronize()
    for a, (Q, R), (Q0, R0) in zip(As0, outs, ref):
        m, n = a.shape
        tol = O.tol_for(m, n)
        Qn, Rn = makb200.to_numpy(Q), makb200.to_numpy(R)
        Qo, Ro = O.qr_compact(a.copy())
        assert O.rel_resid(a, Qn, Rn) <= tol and O.orth_err(Qn) <= tol
        assert np.all(np.tril(Rn, -1) == 0) and np.all(np.real(np.diag(Rn)) >= 0)
        assert np.linalg.norm(Rn - Ro) <= 200 * tol * np.linalg.norm(Ro) and np.linalg.norm(Qn - Qo) <= 200 * tol * np.sqrt(min(m, n))
        assert np.linalg.norm(Qn - makb200.to_numpy(Q0)) <= 200 * tol * np.sqrt(min(m, n))


def test_two_stage_eigh_with_lower_triangle_stage1():
    """MAKB200_SY2SB_LOWER=1: the dense -> band reduction updates only the tiles on/below the diagonal and mirrors
    them (csrc/projections.cuh: mirror_lower_kernel); band eigenvalues and the assembled two-stage eigh_full!."""
    import subprocess
    import sys
    import json
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_twostage", os.path.join(here, "test_gpu_twostage.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    env = dict(os.environ, MAKB200_EIGH_TWOSTAGE="16", MAKB200_SY2SB_LOWER="1")
    p = subprocess.run([sys.executable, "-c", mod.SCRIPT % mod.ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for r in json.loads(line[len("RESULT "):]):
        tol = 10 * r["n"] * EPS
        assert r["band"] <= tol and r["vals"] <= tol and r["resid"] <= tol * r["n"] ** 0.5 and r["orth"] <= tol * r["n"] ** 0.5, r


@pytest.mark.parametrize("extra", [{"MAKB200_SY2SB_LOOKAHEAD": "1"}, {"MAKB200_SY2SB_LOOKAHEAD": "1", "MAKB200_SY2SB_LOWER": "1"}])
def test_two_stage_eigh_with_stage1_lookahead(extra):
    """MAKB200_SY2SB_LOOKAHEAD=1: the next panel's QR runs on the auxiliary stream while the bulk rank-2b update
    drains (alone and together with the lower-triangle update)."""
    import subprocess
    import sys
    import json
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_twostage", os.path.join(here, "test_gpu_twostage.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    env = dict(os.environ, MAKB200_EIGH_TWOSTAGE="16", **extra)
    p = subprocess.run([sys.executable, "-c", mod.SCRIPT % mod.ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for r in json.loads(line[len("RESULT "):]):
        tol = 10 * r["n"] * EPS
        assert r["band"] <= tol and r["vals"] <= tol and r["resid"] <= tol * r["n"] ** 0.5 and r["orth"] <= tol * r["n"] ** 0.5, r
