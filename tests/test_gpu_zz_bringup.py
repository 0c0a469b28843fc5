"""Round-2 BRING-UP tests: kernels written after the round-1 GPU budget was spent.  Their logic is
validated on the CPU emulator (tests/test_emu_kernels_cpu.py, all CTAs as co-resident fibers) but they
have not run on a B200 yet, so they are opt-in at run time (MAKB200_CHASE_PERSISTENT, MAKB200_Q2_FUSED)
and these tests only run with MAKB200_BRINGUP=1 (tools/job_r2a.sh sets it).  Sorted last on purpose."""
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
