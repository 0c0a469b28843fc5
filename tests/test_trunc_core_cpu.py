"""The batched truncation search (csrc/trunc_core.h) compiled with g++ against the host-side
``findtruncated_svd`` / ``truncation_error!`` restatement (matrixalgebrakit.jl_b200/truncation.py, which
tests/test_oracle.py pins to the reference's literals, test/common/truncate.jl:33-95)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import makb200
from makb200 import truncation as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "trunc_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "trunc_host.cpp")])
    return ctypes.CDLL(out)


def _run(lib, spectra, spec, caps=None):
    b = len(spectra)
    k = (ctypes.c_int * b)(*[len(s) for s in spectra])
    ptrs = (ctypes.c_void_p * b)(*[s.ctypes.data for s in spectra])
    rank = (ctypes.c_int * b)()
    eps = (ctypes.c_double * b)()
    capv = (ctypes.c_int * b)(*caps) if caps is not None else None
    assert lib.trunc_select_host(b, k, ptrs, ctypes.byref(spec), capv, rank, eps) == 0
    return list(rank), list(eps)


def _spectra():
    rng = np.random.default_rng(0)
    out = [np.sort(np.abs(rng.standard_normal(n)))[::-1].copy() for n in (1, 2, 16, 37, 100, 512)]
    out.append(np.array([0.9, 0.3, 0.1, 0.01]))                       # the reference's fixed spectrum (svd.jl:198-254)
    out.append(10.0 ** (-np.arange(40) / 3.0))                        # graded
    out.append(np.zeros(5))
    out.append(np.zeros(0))
    out.append(np.array([1.0, 1.0, 1.0, 0.5, 0.5]))                   # ties
    return out


STRATEGIES = [
    T.notrunc(), T.truncrank(3), T.truncrank(0), T.truncrank(1000),
    T.trunctol(atol=0.2), T.trunctol(rtol=0.05), T.trunctol(atol=0.3, rtol=0.01, p=1), T.trunctol(rtol=0.1, p=3),
    T.truncerror(atol=0.25), T.truncerror(rtol=0.13), T.truncerror(atol=0.05, rtol=0.02, p=1), T.truncerror(atol=1e9),
    T.trunc_and(T.truncrank(5), T.trunctol(atol=0.15)),
    T.trunc_and(T.truncrank(20), T.trunctol(rtol=0.01), T.truncerror(rtol=0.05)),
    T.trunc_or(T.trunctol(atol=0.5), T.truncrank(2)),
    T.select_truncation({"atol": 0.2, "maxrank": 7, "minrank": 2}),
    T.select_truncation({"rtol": 1e-3, "maxerror": 0.1}),
    T.select_truncation({"minrank": 3}),
    T.select_truncation({"maxrank": 4}),
]


@pytest.mark.parametrize("si", range(len(STRATEGIES)))
def test_device_rank_and_error_match_host_search(lib, si):
    s = STRATEGIES[si]
    spec = T.device_spec(s)
    assert spec is not None, s
    spectra = _spectra()
    ranks, eps = _run(lib, spectra, spec)
    for v, r, e in zip(spectra, ranks, eps):
        ind = T._find(v, s, svd=True)
        assert np.array_equal(ind, np.arange(r)), (s, v[:8], r, ind)
        assert np.isclose(e, np.linalg.norm(v[r:]), rtol=1e-14, atol=0)


def test_strategies_outside_the_prefix_family_fall_back():
    for s in (T.truncrank(3, rev=False), T.trunctol(atol=0.1, keep_below=True),
              T.trunc_or(T.trunctol(atol=0.1), T.truncerror(atol=0.1)),
              T.trunc_and(T.trunctol(atol=0.1), T.trunctol(rtol=0.1)),
              T.trunctol(atol=0.1, p=np.inf)):
        assert T.device_spec(s) is None
    spec = T.device_spec(T.select_truncation({"atol": 0.2, "maxrank": 7, "minrank": 2, "maxerror": 0.3}))
    assert (spec.maxrank, spec.minrank, spec.by_value, spec.by_error) == (7, 2, 1, 1)
    assert (spec.vatol, spec.eatol) == (0.2, 0.3)


def test_per_block_rank_caps(lib):
    """BASELINE config 3: block i is truncated at truncrank(n_i // 2), optionally on top of a shared tolerance."""
    spectra = [s for s in _spectra() if len(s) > 0]
    caps = [len(s) // 2 for s in spectra]
    for base in (T.notrunc(), T.trunctol(rtol=0.05), T.select_truncation({"atol": 0.2, "maxrank": 7, "minrank": 2})):
        ranks, eps = _run(lib, spectra, T.device_spec(base), caps)
        for v, r, e, cap in zip(spectra, ranks, eps, caps):
            if isinstance(base, T.TruncationUnion):       # (tol & rank & cap) | minrank
                inner = T.trunc_and(base.components[0], T.truncrank(cap))
                st = T.trunc_or(inner, *[c for c in base.components[1:]])
            else:
                st = T.trunc_and(base, T.truncrank(cap))
            ind = T._find(v, st, svd=True)
            assert np.array_equal(ind, np.arange(r)), (base, cap, r, ind)
            assert np.isclose(e, np.linalg.norm(v[r:]), rtol=1e-14, atol=0)
