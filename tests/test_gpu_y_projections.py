"""project_hermitian! / project_antihermitian! / project_isometric! and ishermitian / isantihermitian /
isisometric / isunitary through the C ABI (reference: implementations/projections.jl,
common/matrixproperties.jl; its tests: test/testsuite/projections.jl).  Files named test_gpu_y_* were
added after the round's last GPU run (kernel logic validated on the CPU emulator), so they sort after
the suites that have already run on a B200."""
import numpy as np
import pytest
import torch

from oracle import mak_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _ref(A, anti):
    return (A - A.conj().T) / 2 if anti else (A + A.conj().T) / 2


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("n", [1, 31, 32, 54, 97, 300])
def test_project_hermitian_bit_exact_and_in_place(n, dtype):
    import makb200
    A0 = O.randn_matrix(n, n, dtype, seed=n)
    for anti, f, f_ in ((0, makb200.project_hermitian, makb200.project_hermitian_),
                        (1, makb200.project_antihermitian, makb200.project_antihermitian_)):
        ref = _ref(A0, anti)
        A = makb200.to_device(A0)
        B = f(A)                                            # out of place: A untouched
        assert np.array_equal(makb200.to_numpy(B), ref) and np.array_equal(makb200.to_numpy(A), A0)
        out = f_(A)                                         # in place: returns A itself (projections.jl:38-43)
        assert out is A and np.array_equal(makb200.to_numpy(A), ref)
        # explicit distinct output, strided input view
        big = makb200.colmajor_zeros(n + 7, n, A.dtype, "cuda:0")
        big[:n, :] = makb200.to_device(A0)
        B2 = makb200.colmajor_empty(n, n, A.dtype, "cuda:0")
        assert f_(big[:n, :], B2) is B2 and np.array_equal(makb200.to_numpy(B2), ref)
        assert (makb200.ishermitian, makb200.isantihermitian)[anti](B2)          # exact test on the result
        if n > 1:
            assert not (makb200.ishermitian, makb200.isantihermitian)[1 - anti](B2)


def test_project_hermitian_errors_and_empty():
    import makb200
    with pytest.raises(ValueError):
        makb200.project_hermitian(makb200.to_device(O.randn_matrix(5, 4, "f64", 1)))
    A = makb200.to_device(O.randn_matrix(6, 6, "f64", 1))
    with pytest.raises(ValueError):
        makb200.project_hermitian_(A, makb200.colmajor_empty(5, 5, torch.float64, "cuda:0"))
    E = makb200.colmajor_zeros(0, 0, torch.float64, "cuda:0")
    assert makb200.project_hermitian_(E) is E and makb200.ishermitian(E) and makb200.isantihermitian(E)


@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_ishermitian_exact_and_approx(dtype):
    import makb200
    n = 120
    G = O.randn_matrix(n, n, dtype, seed=4)
    H = _ref(G, 0)
    K = _ref(G, 1)
    Hd, Kd, Gd = (makb200.to_device(x) for x in (H, K, G))
    assert makb200.ishermitian(Hd) and not makb200.ishermitian(Gd) and not makb200.ishermitian(Kd)
    assert makb200.isantihermitian(Kd) and not makb200.isantihermitian(Gd) and not makb200.isantihermitian(Hd)
    Hp = H.copy()
    Hp[n - 1, 0] += 1e-10
    Hpd = makb200.to_device(Hp)
    assert not makb200.ishermitian(Hpd)                                  # exact: one entry off
    assert makb200.ishermitian(Hpd, atol=1e-9) and not makb200.ishermitian(Hpd, atol=1e-12)
    assert makb200.ishermitian(Hpd, rtol=1e-10) and not makb200.ishermitian(Hpd, rtol=1e-14)
    Hq = H.copy()
    Hq[n - 1, 0] += 1e-13                                                # below default_hermitian_tol ~ 3e-12 (defaults.jl:44)
    Hqd = makb200.to_device(Hq)
    assert not makb200.ishermitian(Hqd) and makb200.ishermitian(Hqd, atol=None) and not makb200.ishermitian(Hpd, atol=None)
    # the numbers behind the decision against numpy
    d, mx, fro, bad = makb200.projections.hermitian_props(Gd)
    assert np.isclose(d, np.linalg.norm(K), rtol=1e-12) and np.isclose(mx, np.abs(G).max(), rtol=1e-15)
    assert np.isclose(fro, np.linalg.norm(G), rtol=1e-12) and bad == np.count_nonzero(np.triu(G != G.conj().T))
    # consistent with the check inside eigh_full! (eigh.jl:11-18)
    with pytest.raises(makb200.DomainError):
        makb200.eigh_full(Gd)
    makb200.eigh_full(makb200.project_hermitian(Gd))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(54, 37), (54, 54), (200, 130)])
def test_project_isometric_and_isisometric(m, n, dtype):
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=m + n)
    A = makb200.to_device(A0)
    W = makb200.project_isometric(A)
    Wn = makb200.to_numpy(W)
    assert np.array_equal(makb200.to_numpy(A), A0)
    assert O.orth_err(Wn) <= O.tol_for(m, n)
    Wo, _ = O.left_polar(A0)
    assert np.linalg.norm(Wn - Wo) <= 1e3 * O.tol_for(m, n)             # the polar factor is unique
    W2 = makb200.colmajor_empty(m, n, A.dtype, "cuda:0")
    assert makb200.project_isometric_(makb200.to_device(A0), W2) is W2   # same output object
    assert makb200.isisometric(W) and makb200.isisometric(W, side="left") and not makb200.isisometric(A)
    assert makb200.isisometric(W, side="right") == (m == n)
    assert makb200.isunitary(W) == (m == n)
    Wh = makb200.to_device(np.asfortranarray(Wn.conj().T))
    assert makb200.isisometric(Wh, side="right") and makb200.is_right_isometric(Wh)
    with pytest.raises(ValueError):
        makb200.isisometric(W, side="up")
    with pytest.raises(ValueError):                                      # m < n (projections.jl:27-28)
        makb200.project_isometric(makb200.to_device(O.randn_matrix(n, m + 1, dtype, 2)))


@pytest.mark.parametrize("dtype", ["f64", "c128"])
@pytest.mark.parametrize("m,n", [(1, 1), (54, 37), (37, 54), (200, 200)])
def test_one_uppertriangular_lowertriangular(m, n, dtype):
    """one! / uppertriangular! / lowertriangular! (src/common/initialization.jl:11-36), one launch each, in place,
    strided views included."""
    import makb200
    A0 = O.randn_matrix(m, n, dtype, seed=m * n)
    want = {makb200.one_: np.eye(m, n), makb200.uppertriangular_: np.triu(A0), makb200.lowertriangular_: np.tril(A0)}
    for f, ref in want.items():
        big = makb200.colmajor_zeros(m + 5, n, makb200.to_device(A0).dtype, "cuda:0")
        big[:m, :] = makb200.to_device(A0)
        big[m:, :] = 7.0
        view = big[:m, :]
        assert f(view) is view
        out = makb200.to_numpy(big)
        assert np.array_equal(out[:m], ref.astype(A0.dtype)) and np.all(out[m:] == 7.0)
    E = makb200.colmajor_zeros(0, 4, torch.float64, "cuda:0")
    assert makb200.one_(E) is E
