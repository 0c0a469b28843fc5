"""Batched tridiagonal divide & conquer (csrc/stedc.cu: stedc_batched) replayed on the CPU: the product's host tables
(csrc/stedc_batch_tables.h: blocks laid end to end, one tree per block, a level's merges gathered over all blocks) and
the product's work-item bodies (csrc/stedc_core.h), with the eigenvector matrices stored as NaN-initialised strips of
leading dimension n_max exactly as on the device.  Every block's eigenpairs are compared with numpy: a merge that
crossed a block boundary, an overlap of two blocks' strip regions or a block read from the wrong ping-pong buffer would
break them."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "stedc_batched_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "stedc_batched_host.cpp")])
    return ctypes.CDLL(out)


def _run(lib, blocks):
    ns = [len(d) for d, _ in blocks]
    n = (ctypes.c_int * len(ns))(*ns)
    d = np.concatenate([b[0] for b in blocks]).astype(np.float64)
    e = np.concatenate([np.append(b[1], 0.0) for b in blocks]).astype(np.float64)
    w = np.full(sum(ns), np.nan)
    V = np.full(sum(k * k for k in ns), np.nan)
    stats = (ctypes.c_int * 3)()
    vp = ctypes.c_void_p
    rc = lib.stedc_batched_host(len(ns), n, d.ctypes.data_as(vp), e.ctypes.data_as(vp), w.ctypes.data_as(vp),
                                V.ctypes.data_as(vp), stats)
    assert rc == 0
    outs, o, vo = [], 0, 0
    for k in ns:
        outs.append((w[o:o + k].copy(), V[vo:vo + k * k].reshape(k, k).T.copy()))
        o += k
        vo += k * k
    return outs, list(stats)


def _check(d, e, w, Z):
    n = len(d)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    tol = 10 * n * EPS
    nrm = max(np.abs(np.linalg.eigvalsh(T)).max(), 1e-300)
    assert np.all(np.isfinite(w)) and np.all(np.isfinite(Z))
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - np.linalg.eigvalsh(T))) <= tol * nrm
    assert np.linalg.norm(T @ Z - Z * w) <= tol * nrm * np.sqrt(n)
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= tol


def test_ragged_blocks_with_different_tree_depths(lib):
    rng = np.random.default_rng(1)
    ns = [200, 33, 64, 65, 129, 17, 100, 256, 31, 32, 1, 2, 97]      # depths 0 .. 3 in one pass, both ping-pong parities
    blocks = [(rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))) for n in ns]
    outs, stats = _run(lib, blocks)
    for (d, e), (w, Z) in zip(blocks, outs):
        _check(d, e, w, Z)
    assert stats[1] == 3                                             # deepest tree: 256 / 32
    assert stats[0] == stats[2] - len(ns)                            # merges = leaves - blocks


def test_special_spectra_and_scales(lib):
    rng = np.random.default_rng(2)
    n = 150
    blocks = [
        (np.ones(n), np.zeros(n - 1)),                               # identity: everything deflates
        (np.zeros(n), np.zeros(n - 1)),                              # zero
        (2.0 * np.ones(n), -np.ones(n - 1)),                         # 1-D Laplacian (known spectrum)
        (np.repeat([1.0, 2.0, 3.0], n // 3), 1e-9 * rng.standard_normal(n - 1)),   # clusters glued by tiny couplings
        (1e-100 * rng.standard_normal(n), 1e-100 * rng.standard_normal(n - 1)),    # per-block scaling
        (1e100 * rng.standard_normal(n), 1e100 * rng.standard_normal(n - 1)),
        (rng.standard_normal(70), rng.standard_normal(69)),
    ]
    outs, _ = _run(lib, blocks)
    for (d, e), (w, Z) in zip(blocks, outs):
        _check(d, e, w, Z)
    k = np.arange(1, n + 1)
    assert np.max(np.abs(outs[2][0] - (2 - 2 * np.cos(k * np.pi / (n + 1))))) <= 10 * n * EPS * 4


def test_many_equal_blocks(lib):
    rng = np.random.default_rng(3)
    blocks = [(rng.standard_normal(80), rng.standard_normal(79)) for _ in range(40)]
    outs, stats = _run(lib, blocks)
    for (d, e), (w, Z) in zip(blocks, outs):
        _check(d, e, w, Z)
    assert stats[2] == 40 * 4
