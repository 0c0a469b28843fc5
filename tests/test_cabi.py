"""CPU-side boundary checks: the C-ABI library loads, exports every symbol include/makb200.h
declares, and the Python binding table covers all of them (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "makb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(makb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    import makb200
    lib = makb200._lib.load()
    names = _declared()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in makb200.h but not exported"
        assert nm in makb200._lib.SIGNATURES, f"{nm} has no ctypes signature"
    assert lib.makb200_version() >= 100


def test_no_cpu_fallback():
    import makb200
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = torch.zeros((4, 4), dtype=torch.float64).t()
    with pytest.raises(Exception):
        makb200.qr_compact(A)          # CPU tensors are rejected: no fallback path exists
    rc = makb200._lib.load().makb200_create(ctypes.byref(ctypes.c_void_p()), 0)
    assert rc != 0                      # no device -> create fails loudly


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the shipped package may import or call it."""
    pkg = os.path.join(ROOT, "matrixalgebrakit.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mak_oracle" not in txt, f


def test_algorithm_selection_semantics():
    # src/algorithms.jl:106-124, test/common/algorithms.jl:7-62
    import makb200 as M
    assert M.select_algorithm("qr_compact", None).name == "Householder"
    assert M.select_algorithm("svd_compact", None).name == "SVDViaPolar"
    assert M.select_algorithm("eigh_full", None).name == "DivideAndConquer"
    assert M.select_algorithm("left_polar", None).name == "QDWH"
    assert M.select_algorithm("qr_compact", None, positive=False).get("positive") is False
    assert M.select_algorithm("qr_compact", None, "Householder", positive=True).get("positive") is True
    assert M.select_algorithm("qr_compact", None, M.Householder).name == "Householder"
    assert M.select_algorithm("qr_compact", None, {"positive": False}).get("positive") is False
    alg = M.Householder(positive=True)
    assert M.select_algorithm("qr_compact", None, alg) is alg
    with pytest.raises(ValueError):
        M.select_algorithm("qr_compact", None, alg, positive=False)
    with pytest.raises(ValueError):
        M.select_algorithm("qr_compact", None, {"positive": False}, blocksize=1)
    with pytest.raises(ValueError):
        M.select_algorithm("qr_compact", None, "NoSuchAlg")
    with pytest.raises(ValueError):
        M.Householder(nonsense=1)


def test_truncation_host_logic():
    # test/common/truncate.jl:33-95 (0-based)
    import numpy as np
    import torch
    import makb200 as M
    v = torch.tensor([1, 0.9, 0.5, -0.3, 0.01], dtype=torch.float64)
    assert M.findtruncated(v, M.truncrank(2)).tolist() == [0, 1]
    assert M.findtruncated(v, M.truncrank(2, rev=False)).tolist() == [4, 3]
    assert M.findtruncated_svd(v, M.truncrank(2)).tolist() == [0, 1]
    assert M.findtruncated(v, M.trunctol(atol=0.4)).tolist() == [0, 1, 2]
    assert M.findtruncated_svd(v.abs(), M.trunctol(atol=0.4)).tolist() == [0, 1, 2]
    assert M.findtruncated(v, M.trunctol(atol=0.4, keep_below=True)).tolist() == [3, 4]
    assert M.findtruncated_svd(v.abs(), M.trunctol(atol=0.4, keep_below=True)).tolist() == [3, 4]
    v = torch.tensor([0.01, 1, 0.9, -0.3, 0.5], dtype=torch.float64)
    assert M.findtruncated(v, M.trunctol(atol=0.4)).tolist() == [1, 2, 4]
    assert M.findtruncated(v, M.trunctol(atol=0.2)).tolist() == [1, 2, 3, 4]
    assert M.findtruncated(v, M.trunctol(atol=0.2, keep_below=True)).tolist() == [0]
    assert set(M.findtruncated(v, M.truncerror(atol=0.2)).tolist()) == {1, 2, 3, 4}
    vs = torch.tensor(np.sort(np.abs(v.numpy()))[::-1].copy())
    assert M.findtruncated_svd(vs, M.truncerror(atol=0.2)).tolist() == [0, 1, 2, 3]
    v2 = torch.tensor([1.0, 0.9, 0.5, 0.3, 0.01], dtype=torch.float64)
    assert M.findtruncated_svd(v2, M.trunc_or(M.trunctol(atol=0.4), M.truncrank(4))).tolist() == [0, 1, 2, 3]
    assert M.findtruncated_svd(v2, M.trunc_or(M.trunctol(atol=0.4), M.truncrank(2))).tolist() == [0, 1, 2]
    assert isinstance(M.trunc_or(M.notrunc(), M.truncrank(3)), type(M.notrunc()))
    s = M.select_truncation({"atol": 0.4, "maxrank": 2})
    assert M.findtruncated_svd(v2, s).tolist() == [0, 1]
    s = M.select_truncation({"atol": 0.4, "minrank": 4})
    assert M.findtruncated_svd(v2, s).tolist() == [0, 1, 2, 3]
