"""svd_trunc! with a rank known up front (truncrank(r); svd.jl:226-237, truncation.jl:54-58): the
leading-r path (makb200_svd_leading: partial back-transformation, m x r x n product for U) must give
what the reference's "full compact SVD, then slice" gives.  Sorted after the GPU-verified suites
(see test_gpu_y_projections.py)."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n,r", [(54, 37, 17), (37, 54, 17), (54, 54, 1), (54, 54, 53), (200, 200, 64),
                                   (300, 130, 33), (130, 300, 100), (5, 3, 2), (600, 600, 100)])
def test_svd_trunc_rank_matches_full_then_slice(m, n, r, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=m + 3 * n + r)
    U, S, Vh, eps = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.truncrank(r))
    Un, Sn, Vn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
    assert Un.shape == (m, r) and Sn.shape == (r,) and Vn.shape == (r, n)
    tol = O.tol_for(m, n)
    # the reference's recipe on the LAPACK oracle
    Uo, So, Vho, epso = O.svd_trunc(A0, O.truncrank(r))
    assert np.max(np.abs(Sn - So)) / So[0] <= tol
    assert abs(eps - epso) <= tol * So[0] * np.sqrt(min(m, n))
    assert O.orth_err(Un) <= tol and O.orth_err(Vn, "right") <= tol
    # best rank-r approximation: ||A - U S Vh||_F = eps
    assert abs(np.linalg.norm(A0 - (Un * Sn) @ Vn) - epso) <= tol * np.linalg.norm(A0)
    # gauge (common/gauge.jl:69-77): the entry of largest modulus of every column of U is real positive
    piv = Un[np.argmax(np.abs(Un), axis=0), np.arange(r)]
    assert np.all(np.abs(piv.imag) <= 1e-14) and np.all(piv.real > 0)
    # the same triplets as the full decomposition sliced (vectors up to rounding: same algorithm, fewer columns)
    Uf, Sf, Vf = makb200.svd_compact(makb200.to_device(A0))
    Sf = Sf.cpu().numpy()
    assert np.max(np.abs(Sn - Sf[:r])) <= 4 * EPS * Sf[0]
    gap = np.min(np.abs(np.diff(Sf[: r + 1]))) / Sf[0] if r < min(m, n) else np.min(np.abs(np.diff(Sf))) / Sf[0]
    vtol = 100 * tol / max(gap, 1e-8)
    assert np.linalg.norm(Un - makb200.to_numpy(Uf)[:, :r]) <= vtol
    assert np.linalg.norm(Vn - makb200.to_numpy(Vf)[:r, :]) <= vtol


def test_svd_trunc_rank_no_error_kwargs_and_preallocated():
    import makb200
    A0 = O.randn_matrix(80, 60, "f64", seed=5)
    U, S, Vh = makb200.svd_trunc_no_error(makb200.to_device(A0), trunc=makb200.truncrank(10))
    assert tuple(U.shape) == (80, 10) and tuple(S.shape) == (10,) and tuple(Vh.shape) == (10, 60)
    So = O.svd_vals(A0)
    np.testing.assert_allclose(S.cpu().numpy(), So[:10], rtol=1e-12)
    # keyword form (TruncationStrategy(; maxrank), interface/truncation.jl:37-66) takes the same path
    U2, S2, Vh2, eps2 = makb200.svd_trunc(makb200.to_device(A0), trunc={"maxrank": 10})
    assert tuple(U2.shape) == (80, 10) and abs(eps2 - np.linalg.norm(So[10:])) <= 1e-12 * So[0]
    # rank >= k: nothing to truncate, full path
    U3, S3, Vh3, eps3 = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.truncrank(100))
    assert tuple(U3.shape) == (80, 60) and eps3 == 0.0
    # caller-provided full-size outputs: the reference's full decomposition into them, then the slice
    out = makb200.svd.initialize_output(makb200.to_device(A0))
    U4, S4, Vh4, eps4 = makb200.svd_trunc_(makb200.to_device(A0), out, trunc=makb200.truncrank(10))
    assert tuple(U4.shape) == (80, 10) and abs(eps4 - eps2) <= 1e-12 * So[0]
    assert O.orth_err(makb200.to_numpy(out[0])) <= O.tol_for(80, 60)       # the full U was computed
    # fixgauge=False is honoured
    U5, S5, Vh5, _ = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.truncrank(10), fixgauge=False)
    assert O.rel_resid(makb200.to_numpy(U2) * S2.cpu().numpy(), makb200.to_numpy(U5) * S5.cpu().numpy(),
                       makb200.to_numpy(Vh5) @ makb200.to_numpy(Vh2).conj().T) <= 1e-10


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_trunc_batched_device_select(dtype):
    """Batched svd_trunc!: rank and error of every block from makb200_trunc_select_batched (one launch, one
    read) must equal the per-block host search of the reference recipe on the oracle's values."""
    import makb200
    from makb200 import truncation as T
    rng = np.random.default_rng(3)
    sizes = [(16, 16), (17, 9), (9, 17), (32, 32), (54, 37), (64, 64), (1, 1), (70, 70)]
    sizes += [(int(s), int(s)) for s in rng.integers(16, 64, size=12)]
    As0 = [O.randn_matrix(m, n, dtype, seed=900 + i) for i, (m, n) in enumerate(sizes)]
    for trunc in (makb200.truncrank(5), makb200.trunctol(rtol=0.21), makb200.truncerror(rtol=0.33),
                  {"atol": 1.7, "maxrank": 12, "minrank": 2}, {"maxerror": 2.9, "maxrank": 30}):
        strategy = T.select_truncation(trunc)
        assert T.device_spec(strategy) is not None
        res = makb200.svd_trunc_batched_([makb200.to_device(a) for a in As0], trunc)
        assert len(res) == len(As0)
        for a, (U, S, Vh, eps) in zip(As0, res):
            So = O.svd_vals(a)
            ind = T._find(So, strategy, svd=True)
            r = len(ind)
            assert tuple(U.shape) == (a.shape[0], r) and tuple(S.shape) == (r,) and tuple(Vh.shape) == (r, a.shape[1]), \
                (trunc, a.shape, r, tuple(S.shape))
            np.testing.assert_allclose(S.cpu().numpy(), So[:r], rtol=1e-11, atol=1e-13)
            assert abs(eps - np.linalg.norm(So[r:])) <= 1e-11 * So[0]
            Un, Vn = U.cpu().numpy(), Vh.cpu().numpy()
            if r:
                assert O.orth_err(Un) <= O.tol_for(*a.shape) and O.orth_err(Vn, "right") <= O.tol_for(*a.shape)
                assert abs(np.linalg.norm(a - (Un * S.cpu().numpy()) @ Vn) - np.linalg.norm(So[r:])) <= 1e-11 * So[0]
    # per-block rank caps (BASELINE config 3: truncrank(n_i // 2) per block), alone and on top of a tolerance
    caps = [min(a.shape) // 2 for a in As0]
    for trunc in (None, makb200.trunctol(rtol=0.21)):
        res = makb200.svd_trunc_batched_([makb200.to_device(a) for a in As0], trunc, maxranks=caps)
        for a, cap, (U, S, Vh, eps) in zip(As0, caps, res):
            So = O.svd_vals(a)
            st = T.truncrank(cap) if trunc is None else T.trunc_and(trunc, T.truncrank(cap))
            r = len(T._find(So, st, svd=True))
            assert tuple(U.shape) == (a.shape[0], r) and tuple(Vh.shape) == (r, a.shape[1])
            np.testing.assert_allclose(S.cpu().numpy(), So[:r], rtol=1e-11, atol=1e-13)
            assert abs(eps - np.linalg.norm(So[r:])) <= 1e-11 * So[0]
    # a strategy outside the prefix family takes the per-block host path, same return shape
    res = makb200.svd_trunc_batched_([makb200.to_device(a) for a in As0[:4]], makb200.trunctol(atol=1.0, keep_below=True))
    for a, (U, S, Vh, eps) in zip(As0[:4], res):
        So = O.svd_vals(a)
        keep = So <= 1.0
        np.testing.assert_allclose(S.cpu().numpy(), So[keep], rtol=1e-11, atol=1e-13)
        assert abs(eps - np.linalg.norm(So[~keep])) <= 1e-11 * So[0]


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_eigh_batched_values_only_including_large_blocks(dtype):
    """makb200_eigh_batched with V == NULL: small blocks through the one-CTA kernel, blocks beyond its
    shared-memory limit through makb200_eigh's values-only (Sturm) path."""
    import ctypes as C
    import makb200
    from makb200 import _core
    ns = [8, 24, 40, 64, 90, 150, 260]
    As0 = [O.rand_hermitian(n, dtype, seed=40 + n) for n in ns]
    As = [makb200.to_device(a) for a in As0]
    Ws = [torch.empty(n, dtype=torch.float64, device="cuda:0") for n in ns]
    h = _core.Handle.get(As[0].device)
    b = len(ns)
    IA, VP = C.c_int * b, C.c_void_p * b
    n_ = IA(*ns)
    lda = IA(*[_core.ld(A) for A in As])
    dt = _core.dtype_code(As[0])
    lw = h.lib.makb200_eigh_batched_worksize(h.h, dt, b, n_)
    work = h.workspace(lw)
    rc = h.lib.makb200_eigh_batched(h.h, dt, 0, b, n_, VP(*[A.data_ptr() for A in As]), lda,
                                    VP(*[W.data_ptr() for W in Ws]), None, None, C.c_void_p(0), _core.ptr(work),
                                    work.numel())
    h.check(rc, "makb200_eigh_batched")
    torch.cuda.synchronize()
    for a, W, n in zip(As0, Ws, ns):
        wref = O.eigh_vals(a)
        assert np.max(np.abs(W.cpu().numpy() - wref)) / np.abs(wref).max() <= 10 * n * EPS


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_svd_trunc_vs_oracle(dtype):
    import makb200
    A0 = O.randn_matrix(54, 37, dtype, seed=9)
    S0 = O.svd_vals(A0)
    r = 17
    U, S, Vh, eps = makb200.svd_trunc(makb200.to_device(A0), trunc=makb200.truncrank(r))
    Uo, So, Vho, epso = O.svd_trunc(A0, O.truncrank(r))
    Un, Sn, Vn = makb200.to_numpy(U), S.cpu().numpy(), makb200.to_numpy(Vh)
    assert Un.shape == (54, r) and Vn.shape == (r, 37)
    np.testing.assert_allclose(Sn, S0[:r], rtol=1e-12)
    np.testing.assert_allclose(eps, epso, rtol=1e-10)
    np.testing.assert_allclose(np.linalg.norm(A0 - (Un * Sn) @ Vn, 2), S0[r], rtol=1e-9)
    assert np.linalg.norm(Un - Uo) < 1e-9 and np.linalg.norm(Vn - Vho) < 1e-9
    # equivalence of the strategies (svd.jl:156-197)
    for tr in (makb200.trunctol(atol=S0[r] + 1e-9), makb200.truncerror(atol=np.linalg.norm(S0[r:]) + 1e-9),
               {"maxrank": r}):
        U2, S2, Vh2 = makb200.svd_trunc_no_error(makb200.to_device(A0), trunc=tr)
        assert S2.numel() == r
    # fixed spectrum fixture (svd.jl:198-254)
    Uq, _ = O.qr_compact(O.randn_matrix(4, 4, dtype, 1))
    Vq, _ = O.qr_compact(O.randn_matrix(4, 4, dtype, 2))
    Sd = np.array([0.9, 0.3, 0.1, 0.01])
    A4 = (Uq * Sd) @ Vq
    for tr, keep in (({"rtol": 0.2, "maxrank": 1}, 1), ({"rtol": 0.2, "maxrank": 3}, 2), ({"rtol": 0.5, "minrank": 3}, 3),
                     ({"rtol": 0.2, "minrank": 1}, 2), (makb200.trunctol(atol=0.2), 2)):
        U4, S4, V4, e4 = makb200.svd_trunc(makb200.to_device(A4), trunc=tr)
        np.testing.assert_allclose(S4.cpu().numpy(), Sd[:keep], rtol=1e-12)
        np.testing.assert_allclose(e4, np.linalg.norm(Sd[keep:]), rtol=1e-10)
    with pytest.raises(ValueError):
        talg = makb200.TruncatedAlgorithm(makb200.SVDViaPolar(), makb200.trunctol(atol=0.2))
        makb200.svd_trunc(makb200.to_device(A4), alg=talg, trunc={"maxrank": 2})
