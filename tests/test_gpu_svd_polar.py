"""svd_compact!/svd_trunc!/left_polar! on B200 vs the LAPACK-replay oracle.
Tolerances (north_star): ||A - U S Vh||/||A||, ||U^H U - I||_F, ||Vh Vh^H - I||_F,
max|s - s_oracle|/s_1, ||W P - A||/||A||, ||W^H W - I||_F <= 10*n*eps, n = max(m, n)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu


def _svd(A_np, **kw):
    import makb200
    A = makb200.to_device(A_np)
    U, S, Vh = makb200.svd_compact(A, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(makb200.to_numpy(A), A_np)
    return makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)


def _check_svd(A, U, S, Vh, vec_cmp=True):
    m, n = A.shape
    k = min(m, n)
    tol = O.tol_for(m, n)
    Uo, So, Vho = O.svd_compact(A)
    assert U.shape == (m, k) and S.shape == (k,) and Vh.shape == (k, n)
    assert np.all(S >= 0) and np.all(np.diff(S) <= 0)
    assert np.max(np.abs(S - So)) / So[0] <= tol
    assert O.rel_resid(A, U * S, Vh) <= tol
    assert O.orth_err(U) <= tol
    assert O.orth_err(Vh, "right") <= tol
    piv = O._argmaxabs_cols(U)
    assert np.all(piv.real > 0) and np.all(np.abs(piv.imag) <= 1e-15)
    if vec_cmp and m > 1:
        gap = np.minimum(np.abs(np.diff(So, prepend=np.inf)), np.abs(np.diff(So, append=-np.inf)))
        gap = np.minimum(gap, So) if m != n else gap
        err = np.maximum(np.linalg.norm(U - Uo, axis=0), np.linalg.norm(Vh - Vho, axis=1))
        mod = np.abs(Uo)
        top2 = np.sort(mod, axis=0)[-2:]
        tie = (top2[1] - top2[0]) < 1e-8
        assert np.all((err <= 200 * max(m, n) * O.EPS * So[0] / gap + 1e-12) | tie)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (37, 54), (1, 1), (3, 2), (200, 200), (300, 130), (130, 300),
                                 (513, 513)])
def test_svd_compact_vs_oracle(m, n, dtype):
    A = O.randn_matrix(m, n, dtype, seed=123 + m + n)
    U, S, Vh = _svd(A)
    _check_svd(A, U, S, Vh)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(0, 5), (5, 0), (0, 0)])
def test_svd_empty(m, n, dtype):
    # test/decompositions/svd.jl:22, svd.jl:197
    import makb200
    A = makb200.to_device(O.randn_matrix(m, n, dtype, 1))
    U, S, Vh = makb200.svd_compact(A)
    k = min(m, n)
    assert tuple(U.shape) == (m, k) and tuple(Vh.shape) == (k, n) and S.numel() == k


def test_svd_graded_and_vals():
    import makb200
    n = 128
    Uq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 5))
    Vq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 6))
    sv = 10.0 ** (-12 * np.arange(n) / n)
    A = (Uq * sv) @ Vq
    U, S, Vh = _svd(A)
    tol = O.tol_for(n)
    assert np.max(np.abs(S - sv)) / sv[0] <= tol
    assert O.rel_resid(A, U * S, Vh) <= tol and O.orth_err(U) <= tol and O.orth_err(Vh, "right") <= tol
    # svd_vals! (values-only path) on this graded spectrum: tests/test_gpu_y_vals.py
    U2, S2, Vh2 = makb200.svd_compact(makb200.to_device(A), fixgauge=False)
    assert O.rel_resid(A, makb200.to_numpy(U2) * S2.cpu().numpy(), makb200.to_numpy(Vh2)) <= tol


# test_svd_trunc_vs_oracle lives in tests/test_gpu_y_trunc.py (truncrank takes the leading-rank path)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (200, 200), (400, 150), (1, 1), (600, 600)])
def test_left_polar_vs_oracle(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=123 + m)
    A = makb200.to_device(A0)
    W, P = makb200.left_polar(A)
    torch.cuda.synchronize()
    Wn, Pn = makb200.to_numpy(W), makb200.to_numpy(P)
    Wo, Po = O.left_polar(A0)
    tol = O.tol_for(m, n)
    assert O.rel_resid(A0, Wn, Pn) <= tol
    assert O.orth_err(Wn) <= tol
    assert np.array_equal(Pn, Pn.conj().T)
    assert np.linalg.eigvalsh(Pn).min() > 0
    cond = np.linalg.cond(A0)
    assert np.linalg.norm(Wn - Wo) <= 50 * tol * max(1.0, cond / 100)
    assert np.linalg.norm(Pn - Po) / np.linalg.norm(Po) <= 50 * tol


def test_left_polar_skip_p_identity_and_errors():
    import makb200
    A0 = O.randn_matrix(54, 37, "c128", seed=3)
    W = makb200.colmajor_empty(54, 37, torch.complex128, "cuda")
    P = makb200.colmajor_empty(37, 37, torch.complex128, "cuda")
    W2, P2 = makb200.left_polar_(makb200.to_device(A0), (W, P))
    assert W2 is W and P2 is P          # test/testsuite/decompositions/polar.jl:30-31
    Wn = makb200.to_numpy(W)
    Pe = makb200.colmajor_empty(0, 0, torch.complex128, "cuda")
    W3, _ = makb200.left_polar_(makb200.to_device(A0), (makb200.colmajor_empty(54, 37, torch.complex128, "cuda"), Pe))
    assert np.linalg.norm(makb200.to_numpy(W3) - Wn) < 1e-12
    with pytest.raises(ValueError):
        makb200.left_polar(makb200.to_device(O.randn_matrix(3, 5)))
    # PolarViaSVD on the B200 SVD gives the same factors
    W4, P4 = makb200.left_polar(makb200.to_device(A0), alg=makb200.PolarViaSVD())
    assert np.linalg.norm(makb200.to_numpy(W4) - Wn) < 1e-11
    assert np.linalg.norm(makb200.to_numpy(P4) - makb200.to_numpy(P)) < 1e-11
    # ill-conditioned (kappa = 1e10): still an isometry with a small residual
    n = 100
    Uq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 5))
    Vq, _ = O.qr_compact(O.randn_matrix(n, n, "f64", 6))
    A1 = (Uq * 10.0 ** (-10 * np.arange(n) / n)) @ Vq
    W5, P5 = makb200.left_polar(makb200.to_device(A1))
    W5, P5 = makb200.to_numpy(W5), makb200.to_numpy(P5)
    assert O.orth_err(W5) <= O.tol_for(n) and O.rel_resid(A1, W5, P5) <= O.tol_for(n)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_batched_vs_oracle(dtype):
    import makb200
    rng = np.random.default_rng(7)
    sizes = [(16, 16), (17, 9), (9, 17), (32, 32), (54, 37), (37, 54), (64, 64), (1, 1), (70, 70), (150, 120)]
    sizes += [(int(s), int(s)) for s in rng.integers(16, 72, size=24)]
    As0 = [O.randn_matrix(m, n, dtype, seed=300 + i) for i, (m, n) in enumerate(sizes)]
    outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As0])
    torch.cuda.synchronize()
    for a, (U, S, Vh) in zip(As0, outs):
        _check_svd(a, makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh))
    # batched svd_trunc! (device-side truncation search): tests/test_gpu_y_trunc.py


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_batched_pooled_threads(dtype):
    """Many blocks too large for the one-CTA kernel: they fan out over the stream pool with one host
    thread per stream (capi.cu: run_pooled); ragged sizes so the threads finish out of order."""
    import makb200
    rng = np.random.default_rng(11)
    dims = [int(v) for v in rng.integers(90, 200, size=20)]
    sizes = [(d, d) for d in dims[:14]] + [(d, d - 30) for d in dims[14:17]] + [(d - 30, d) for d in dims[17:]]
    sizes += [(24, 24), (40, 33)]          # small ones mixed in (one-CTA Jacobi kernel)
    As0 = [O.randn_matrix(m, n, dtype, seed=700 + i) for i, (m, n) in enumerate(sizes)]
    outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As0])
    torch.cuda.synchronize()
    for a, (U, S, Vh) in zip(As0, outs):
        _check_svd(a, makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh), vec_cmp=False)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_and_eigh_batched_graph_replay(dtype, monkeypatch):
    """Opt-in CUDA-graph replay of the per-block path (capi.cu: run_graphed): one captured svd_t / eigh_t per shape and
    stream slot, replayed per block between a staged copy-in and copy-out; rank-deficient blocks are redone uncaptured."""
    import makb200
    monkeypatch.setenv("MAKB200_BATCH_GRAPHS", "1")
    rng = np.random.default_rng(5)
    dims = [int(v) for v in rng.integers(90, 180, size=6)] * 3            # repeated shapes: the graphs are reused
    sizes = [(d, d) for d in dims[:12]] + [(d, d - 20) for d in dims[12:15]] + [(d - 20, d) for d in dims[15:]]
    As0 = [O.randn_matrix(m, n, dtype, seed=900 + i) for i, (m, n) in enumerate(sizes)]
    As0[3] = np.asfortranarray(As0[3][:, :1] @ As0[3][:1, :])             # rank one: takes the redo path
    for rep in range(2):                                                  # second call: cached graphs
        outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As0])
        torch.cuda.synchronize()
        for a, (U, S, Vh) in zip(As0, outs):
            Un, Sn, Vhn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
            tol = O.tol_for(*a.shape)
            so = O.svd_vals(a)
            assert np.max(np.abs(Sn - so)) <= tol * so[0]
            assert np.linalg.norm(a - (Un * Sn) @ Vhn) <= tol * np.linalg.norm(a)
            assert O.orth_err(Un) <= tol and O.orth_err(Vhn, "right") <= tol
    Hs0 = [O.rand_hermitian(d, dtype, seed=950 + i) for i, d in enumerate(dims[:10])]
    outs = makb200.eigh_full_batched_([makb200.to_device(a) for a in Hs0])
    torch.cuda.synchronize()
    for a, (D, V) in zip(Hs0, outs):
        w, Vn = D.cpu().numpy(), makb200.to_numpy(V)
        tol = O.tol_for(a.shape[0])
        assert np.max(np.abs(w - O.eigh_vals(a))) <= tol * np.abs(w).max()
        assert np.linalg.norm(a @ Vn - Vn * w) <= tol * np.linalg.norm(a) and O.orth_err(Vn) <= tol


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("lockstep", ["1", "0", "noest"])
def test_svd_batched_lockstep_qdwh(dtype, lockstep, monkeypatch, capfd):
    """Mid-size blocks (beyond the one-CTA Jacobi kernel) in a chunk of >= 8: phase 1 of the phased batched SVD is the
    lock-step QDWH (csrc/polar_lockstep_plan.h: one grouped GEMM / batched kernel per step for ALL blocks; the graded,
    rank-one and zero blocks fall back to the l0 = eps schedule, the others share the schedule of the chunk's smallest
    sigma_min estimate); with MAKB200_SVD_LOCKSTEP=0 the same blocks take the per-block chain.  Ragged square and tall blocks, odd sizes (unaligned
    Float64 leading dimensions), and the hard cases: rank one, zero, graded kappa = 1e10, tiny and huge scale."""
    import makb200
    if lockstep == "noest":      # lock-step on the fixed l0 = eps schedule (no per-chunk sigma_min estimate)
        monkeypatch.setenv("MAKB200_LS_ESTIMATE", "0")
        lockstep = "1"
    monkeypatch.setenv("MAKB200_SVD_LOCKSTEP", lockstep)
    monkeypatch.setenv("MAKB200_LOCKSTEP_VERBOSE", "1")
    rng = np.random.default_rng(21)
    dims = [int(v) for v in rng.integers(90, 330, size=14)] + [129, 257, 300]
    sizes = [(d, d) for d in dims] + [(260, 200), (331, 97), (150, 140)]
    As0 = [O.randn_matrix(m, n, dtype, seed=1100 + i) for i, (m, n) in enumerate(sizes)]
    n0 = 120
    Uq, _ = O.qr_compact(O.randn_matrix(n0, n0, dtype, 5))
    Vq, _ = O.qr_compact(O.randn_matrix(n0, n0, dtype, 6))
    As0.append(np.asfortranarray((Uq * 10.0 ** (-10 * np.arange(n0) / n0)) @ Vq))              # graded
    As0.append(np.asfortranarray(As0[0][:, :1] @ As0[0][:1, :]))                                # rank one
    As0.append(np.asfortranarray(np.zeros((100, 100), dtype=As0[0].dtype)))                     # zero
    As0.append(np.asfortranarray(1e-100 * O.randn_matrix(111, 111, dtype, 7)))                  # tiny
    As0.append(np.asfortranarray(1e100 * O.randn_matrix(96, 96, dtype, 8)))                     # huge
    outs = makb200.svd_compact_batched_([makb200.to_device(a) for a in As0])
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert ("in lock-step" in err) == (lockstep == "1"), err
    for a, (U, S, Vh) in zip(As0, outs):
        Un, Sn, Vhn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
        tol = O.tol_for(*a.shape)
        so = O.svd_vals(a)
        nrm = np.linalg.norm(a)
        assert np.all(np.isfinite(Un)) and np.all(np.isfinite(Vhn))
        assert np.all(Sn >= 0) and np.all(np.diff(Sn) <= 0)
        assert np.max(np.abs(Sn - so)) <= tol * max(so[0], 1e-300)
        assert np.linalg.norm(a - (Un * Sn) @ Vhn) <= tol * nrm
        assert O.orth_err(Un) <= tol and O.orth_err(Vhn, "right") <= tol
