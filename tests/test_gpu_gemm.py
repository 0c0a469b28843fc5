"""DMMA GEMM vs a float64 torch/numpy reference (floating point kernel: tolerance stated below)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


def _mk(m, n, dtype, seed):
    import makb200
    return makb200.to_device(O.randn_matrix(m, n, dtype, seed))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("opa,opb", [("N", "N"), ("C", "N"), ("N", "C"), ("T", "T"), ("T", "N")])
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (257, 131, 77), (33, 1000, 5), (5, 7, 3000), (640, 512, 300),
                                   # even leading dimensions -> the TMA-fed kernel (f64): ragged M/N/K edges come from
                                   # the tensor map's zero fill, K = 4096 on one tile takes split-K
                                   (250, 130, 78), (1026, 36, 200), (128, 128, 4096), (66, 514, 130), (2, 2, 2), (18, 4, 1)])
def test_gemm_matches_reference(m, n, k, opa, opb, dtype):
    import makb200
    A = _mk(*((m, k) if opa == "N" else (k, m)), dtype, 1)
    B = _mk(*((k, n) if opb == "N" else (n, k)), dtype, 2)
    Cm = _mk(m, n, dtype, 3)
    alpha, beta = (0.7, -1.3) if dtype == "f64" else (0.7 - 0.2j, -1.3 + 0.5j)
    An, Bn, Cn = makb200.to_numpy(A), makb200.to_numpy(B), makb200.to_numpy(Cm)
    op = {"N": lambda x: x, "T": lambda x: x.T, "C": lambda x: x.conj().T}
    ref = alpha * (op[opa](An) @ op[opb](Bn)) + beta * Cn
    makb200.gemm_(Cm, A, B, alpha, beta, opa, opb)
    torch.cuda.synchronize()
    got = makb200.to_numpy(Cm)
    # tolerance: k * eps * |A||B| growth bound
    scale = np.abs(op[opa](An)) @ np.abs(op[opb](Bn)) + np.abs(Cn) * 2
    assert np.max(np.abs(got - ref) / scale) < 4 * np.finfo(float).eps * max(8, np.sqrt(k))


def test_gemm_strided_views_and_beta_zero():
    import makb200
    big = _mk(300, 300, "f64", 4)
    A = big[3:203, 5:105]          # lda = 300, odd offset -> 8-byte cp.async path
    B = big[10:110, 7:157]
    Cm = makb200.colmajor_empty(200, 150, torch.float64, big.device)
    Cm.fill_(float("nan"))          # beta = 0 must not read C
    makb200.gemm_(Cm, A, B, 1.0, 0.0)
    ref = makb200.to_numpy(A) @ makb200.to_numpy(B)
    assert np.allclose(makb200.to_numpy(Cm), ref, rtol=1e-13, atol=1e-12)
