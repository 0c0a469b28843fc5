"""Rank-deficient input to svd_compact!/svd_trunc! on the single-matrix path (QDWH polar + eigh): the polar
factor of a singular matrix is only a partial isometry, so U is re-orthonormalised when its column norms
show it (polar.cu: col_norm_defect_kernel + Householder QR of U).  LAPACK (the reference path) returns
isometric factors for any input; so must we.  Sorted after the GPU-verified suites (see
test_gpu_y_projections.py)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _lowrank(m, n, r, dtype, seed):
    return np.asfortranarray(O.randn_matrix(m, r, dtype, seed) @ O.randn_matrix(r, n, dtype, seed + 1))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,r", [(120, 90, 40), (90, 120, 40), (200, 200, 199), (150, 150, 1), (300, 100, 60)])
def test_svd_compact_rank_deficient(m, n, r, dtype):
    import makb200
    A0 = _lowrank(m, n, r, dtype, seed=m + n + r)
    U, S, Vh = makb200.svd_compact(makb200.to_device(A0))
    Un, Sn, Vn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
    tol = O.tol_for(m, n)
    So = O.svd_vals(A0)
    assert np.max(np.abs(Sn - So)) / So[0] <= tol
    assert np.all(Sn[r:] <= 100 * tol * So[0])                        # numerically zero beyond the rank
    assert O.rel_resid(A0, Un * Sn, Vn) <= tol
    assert O.orth_err(Un) <= tol and O.orth_err(Vn, "right") <= tol    # isometric factors, as LAPACK returns
    # the leading r triplets are those of the oracle (gauge-fixed; scaled by the gap to the next value)
    Uo, _, Vho = O.svd_compact(A0)
    gaps = np.abs(np.diff(np.append(So[:r], 0.0))) / So[0]
    ok = gaps > 1e-3
    assert np.max(np.abs(np.abs(np.sum(Un[:, :r].conj() * Uo[:, :r], axis=0))[ok] - 1)) <= 1e-8


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_of_zero_and_trunc_of_rank_deficient(dtype):
    import makb200
    Z = np.zeros((100, 80), dtype=np.float64 if dtype == "f64" else np.complex128, order="F")
    U, S, Vh = makb200.svd_compact(makb200.to_device(Z))
    Un, Vn = makb200.to_numpy(U), makb200.to_numpy(Vh)
    assert np.all(S.cpu().numpy() == 0)
    assert O.orth_err(Un) <= O.tol_for(100, 80) and O.orth_err(Vn, "right") <= O.tol_for(100, 80)
    A0 = _lowrank(140, 110, 30, dtype, seed=3)
    U, S, Vh, eps = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.trunctol(rtol=1e-10))
    assert tuple(S.shape) == (30,) and eps <= 1e-10 * float(S[0])
    assert O.rel_resid(A0, makb200.to_numpy(U) * S.cpu().numpy(), makb200.to_numpy(Vh)) <= O.tol_for(140, 110)
    U, S, Vh, eps = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.truncrank(50))   # leading-r path, r > rank
    assert tuple(U.shape) == (140, 50) and O.orth_err(makb200.to_numpy(U)) <= O.tol_for(140, 110)
    assert O.rel_resid(A0, makb200.to_numpy(U) * S.cpu().numpy(), makb200.to_numpy(Vh)) <= O.tol_for(140, 110)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,r", [(120, 90, 40), (150, 150, 149), (100, 80, 0)])
def test_left_polar_and_project_isometric_rank_deficient(m, n, r, dtype):
    """left_polar! / project_isometric! on a singular matrix: QDWH alone gives a partial isometry (||W||_F^2 = rank);
    the host layer detects it (makb200_fro2) and takes the PolarViaSVD recipe (polar.jl:59-70), so W is isometric,
    P Hermitian PSD and A = W P, as with the reference's LAPACK path."""
    import makb200
    A0 = _lowrank(m, n, r, dtype, seed=m + r) if r > 0 else np.zeros((m, n), dtype=O.randn_matrix(1, 1, dtype).dtype, order="F")
    W, P = makb200.left_polar(makb200.to_device(A0))
    Wn, Pn = makb200.to_numpy(W), makb200.to_numpy(P)
    tol = O.tol_for(m, n)
    assert O.orth_err(Wn) <= tol
    assert np.linalg.norm(Wn @ Pn - A0) <= tol * max(np.linalg.norm(A0), 1e-300) or (r == 0 and np.linalg.norm(Wn @ Pn) == 0)
    assert np.linalg.norm(Pn - Pn.conj().T) <= tol * max(np.linalg.norm(Pn), 1e-300) or r == 0
    ev = np.linalg.eigvalsh((Pn + Pn.conj().T) / 2)
    assert ev.min() >= -tol * max(ev.max(), 1e-300) - 1e-300
    Wi = makb200.project_isometric(makb200.to_device(A0))
    assert makb200.isisometric(Wi) and O.orth_err(makb200.to_numpy(Wi)) <= tol
    # a full-rank input keeps taking the pure QDWH path and returns the caller's objects
    A1 = O.randn_matrix(m, n, dtype, seed=1)
    W2 = makb200.colmajor_empty(m, n, W.dtype, "cuda:0")
    P2 = makb200.colmajor_empty(n, n, W.dtype, "cuda:0")
    Wr, Pr = makb200.left_polar_(makb200.to_device(A1), (W2, P2))
    assert Wr is W2 and Pr is P2 and O.orth_err(makb200.to_numpy(W2)) <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_batched_svd_collapsed_columns_are_completed(dtype):
    """One-CTA batched SVD: blocks with exactly-zero singular values - a zero block, a zero
    column, a wide block with a zero row, rank one - must still give an isometric U and Vh (LAPACK contract)."""
    import makb200
    rng = np.random.default_rng(3)

    def rnd(m, n):
        a = rng.standard_normal((m, n))
        return a + 1j * rng.standard_normal((m, n)) if dtype == "c128" else a

    blocks = []
    blocks.append(np.zeros((20, 12)))                               # zero block, tall
    blocks.append(np.zeros((9, 17)))                                # zero block, wide
    a = rnd(30, 16); a[:, 5] = 0; a[:, 11] = 0; blocks.append(a)    # zero columns
    a = rnd(14, 40); a[3, :] = 0; blocks.append(a)                  # wide block with a zero row
    a = np.outer(rnd(24, 1)[:, 0], np.ones(24)); blocks.append(a)   # rank one, square (values exactly repeated)
    a = rnd(32, 32); a[:, 16:] = 0; blocks.append(a)                # half of the columns zero
    blocks.append(rnd(25, 25))                                      # a regular block in the same launch
    blocks = [np.asfortranarray(b.astype(np.complex128 if dtype == "c128" else np.float64)) for b in blocks]
    outs = makb200.svd_compact_batched_([makb200.to_device(b) for b in blocks])
    torch.cuda.synchronize()
    for b, (U, S, Vh) in zip(blocks, outs):
        Un, Sn, Vhn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
        tol = O.tol_for(*b.shape)
        so = O.svd_vals(b)
        assert np.all(np.diff(Sn) <= 0) and np.all(Sn >= 0)
        assert np.max(np.abs(Sn - so)) <= tol * max(so[0], 1.0)
        assert np.linalg.norm(b - (Un * Sn) @ Vhn) <= tol * max(np.linalg.norm(b), 1.0)
        assert O.orth_err(Un) <= tol, "U must be an isometry for rank-deficient blocks too"
        assert O.orth_err(Vhn, "right") <= tol
