"""Tile geometry of the persistent tridiagonalisation column kernel (csrc/trd_tiles.h): the product header compiled
with g++ and driven by tests/cpu_harness/trd_tiles_host.cpp, which replays chunking, band/strip walk, partial-slot
addressing and the partial sums of csrc/trd2.cuh.  Checked: every element of the trailing lower triangle is visited
exactly once, no partial slot is written twice or read unwritten, and y = A22 v, v^H A22 v match numpy."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "trd_tiles_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "cpu_harness", "trd_tiles_host.cpp")])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("n,row0,BH,CW,G", [
    (64, 1, 256, 16, 296), (300, 1, 256, 16, 296), (300, 299, 256, 16, 296), (300, 255, 256, 16, 296),
    (300, 256, 256, 16, 296), (300, 257, 256, 16, 296), (1000, 1, 256, 16, 296), (1000, 17, 256, 8, 296),
    (1000, 511, 256, 8, 4), (1000, 512, 256, 16, 7), (777, 300, 256, 16, 2), (1100, 33, 256, 16, 296),
    (1537, 1, 256, 16, 296), (1537, 770, 256, 8, 296), (1024, 1, 256, 16, 13), (1024, 1023, 256, 16, 13),
    (520, 100, 64, 8, 5), (130, 3, 32, 8, 3), (129, 128, 32, 16, 9), (97, 1, 32, 16, 1),
])
def test_tile_replay_matches_dense(lib, n, row0, BH, CW, G):
    rng = np.random.default_rng(n + row0)
    M = rng.standard_normal((n, n))
    M = (M + M.T) / 2
    A = np.asfortranarray(np.tril(M) + np.triu(np.full((n, n), np.nan), 1))   # the upper triangle must never be used
    A[:, :row0] = np.nan if row0 > 0 else A[:, :row0]                           # nor anything left of the trailing block
    A = np.asfortranarray(np.where(np.isnan(A), 1e300, A))
    v = np.zeros(n)
    v[row0:] = rng.standard_normal(n - row0)
    y = np.zeros(n)
    q = ctypes.c_double()
    vp = ctypes.c_void_p
    bad = lib.trd_tiles_replay(n, row0, BH, CW, G, A.ctypes.data_as(vp), v.ctypes.data_as(vp), y.ctypes.data_as(vp),
                               ctypes.byref(q))
    assert bad == 0
    yref = M[row0:, row0:] @ v[row0:]
    assert np.linalg.norm(y[row0:] - yref) <= 1e-12 * max(1.0, np.linalg.norm(yref))
    assert abs(q.value - v[row0:] @ yref) <= 1e-10 * max(1.0, abs(v[row0:] @ yref))
