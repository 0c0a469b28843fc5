"""Host-layer marshalling check without a GPU: every operator of the Python mirror is driven with CPU
tensors and a stand-in handle whose C handle is NULL.  ctypes converts and checks all arguments against
``_lib.SIGNATURES`` (count and types) before the call, and every C entry point returns -1 ("argument 1":
NULL handle) before touching CUDA, so this pins the binding of each call site — in particular of the
entry points added after the last GPU run — on the CPU box.  No numerical claim is made here."""
import ctypes as C

import numpy as np
import pytest
import torch

import makb200
from makb200 import _core


class _FakeHandle:
    def __init__(self):
        self.lib = makb200._lib.load()
        self.h = C.c_void_p(0)
        self.device = torch.device("cpu")
        self.calls = []

    def workspace(self, nbytes):
        return torch.empty(max(int(nbytes), 1024), dtype=torch.uint8)

    def check(self, rc, what):
        self.calls.append((what, rc))


@pytest.fixture
def fake(monkeypatch):
    fh = _FakeHandle()
    monkeypatch.setattr(_core.Handle, "get", classmethod(lambda cls, device: fh))
    return fh


def _cm(m, n, dtype=torch.float64, fill=None):
    t = torch.zeros((n, m), dtype=dtype).t()
    if fill is not None:
        t.copy_(torch.as_tensor(fill, dtype=dtype))
    return t


def _herm(n, dtype=torch.float64):
    g = torch.randn(n, n, dtype=dtype)
    return _cm(n, n, dtype, (g + g.conj().t()) / 2)


@pytest.mark.parametrize("dtype", [torch.float64, torch.complex128])
def test_every_call_site_marshals(fake, dtype, monkeypatch):
    # the Hermitian pre-check reads two doubles back; with a NULL handle the C side writes nothing
    monkeypatch.setattr(makb200.eigh, "hermitian_defect", lambda A: (0.0, 1.0))
    A = _cm(12, 8, dtype, torch.randn(12, 8, dtype=dtype))
    H = _herm(9, dtype)
    makb200.qr_compact(A)
    makb200.qr_full(A)
    makb200.qr_null(A)
    makb200.lq_compact(_cm(8, 12, dtype))
    makb200.left_orth(A)
    makb200.svd_compact(A)
    makb200.svd_vals(A)
    makb200.svd_trunc_no_error(A, trunc=makb200.truncrank(3))          # makb200_svd_leading
    makb200.svd_trunc_no_error(A, trunc=makb200.trunctol(atol=0.5))    # full path + host search
    makb200.eigh_full(H)
    makb200.eigh_vals(H)                                               # makb200_eigh with V = NULL
    makb200.left_polar(A)
    makb200.project_hermitian(H)
    makb200.project_antihermitian_(H)
    makb200.project_isometric(A)
    makb200.gemm_(_cm(8, 8, dtype), A, A, opa="C", opb="N")
    makb200.qr_compact_batched_([_cm(6, 6, dtype), _cm(40, 30, dtype)])
    makb200.svd_compact_batched_([_cm(6, 6, dtype), _cm(40, 30, dtype)])
    makb200.svd_vals_batched_([_cm(6, 6, dtype), _cm(40, 30, dtype)])
    makb200.eigh_full_batched_([_herm(6, dtype), _herm(70, dtype)], check=False)
    makb200.eigh_vals_batched_([_herm(6, dtype), _herm(70, dtype)], check=False)
    names = [w for w, _ in fake.calls]
    for want in ("makb200_qr", "makb200_svd", "makb200_svd_leading", "makb200_eigh", "makb200_polar_qdwh",
                 "makb200_project_hermitian", "makb200_gemm", "makb200_svd_batched",
                 "makb200_eigh_batched", "makb200_adjoint"):
        assert want in names, (want, names)
    assert any(w.startswith("makb200_qr_batched") for w in names), names
    assert all(rc == -1 for _, rc in fake.calls), fake.calls             # NULL handle: argument 1


def test_property_and_truncation_call_sites(fake, monkeypatch):
    H = _herm(7)
    # out.tolist() of an uninitialised buffer: only the call itself is under test
    makb200.projections.hermitian_props(H)
    makb200.projections.hermitian_props(H, anti=True)
    makb200.is_left_isometric(_cm(9, 4))
    makb200.is_right_isometric(_cm(4, 9))
    spec = makb200.truncation.device_spec(makb200.select_truncation({"atol": 0.1, "maxrank": 3, "minrank": 1}))
    ranks, eps = makb200.truncation.trunc_select_batched_([torch.rand(5, dtype=torch.float64), torch.rand(0, dtype=torch.float64),
                                                           torch.rand(9, dtype=torch.float64)], spec)
    assert len(ranks) == 3 and len(eps) == 3
    names = [w for w, _ in fake.calls]
    for want in ("makb200_hermitian_props", "makb200_gram_defect", "makb200_gemm", "makb200_trunc_select_batched"):
        assert want in names
    assert all(rc == -1 for _, rc in fake.calls), fake.calls


def test_trunc_spec_layout_matches_the_header():
    """ctypes mirror of makb200_trunc_spec: 4 ints then 6 doubles, 64 bytes, no hidden padding."""
    S = makb200.truncation.TruncSpec
    assert C.sizeof(S) == 64
    assert [f[0] for f in S._fields_] == ["maxrank", "minrank", "by_value", "by_error", "vatol", "vrtol", "vp",
                                          "eatol", "ertol", "ep"]
    assert S.vatol.offset == 16 and S.ep.offset == 56
