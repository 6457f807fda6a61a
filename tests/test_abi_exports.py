"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/denovo_b200.h declares, struct layouts match the header,
and the product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import denovo_b200 as D
from denovo_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "denovo_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dnv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = D.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.EXPORTS) == names
    assert lib.dnv_abi_version() == 2


def test_struct_sizes_match_header():
    assert C.sizeof(L.Params) == 40
    assert C.sizeof(L.Stats) == 56
    assert D.CALL_DT.itemsize == 80 and D.SITE_DT.itemsize == 48
    assert D.VCALL_DT.itemsize == 32 and D.PCALL_DT.itemsize == 24


def test_no_oracle_dependency_in_product():
    """The product tree must not reference the test oracle."""
    pkg = os.path.join(ROOT, "denovo-variant-caller-java_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", ".java", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "denovo_oracle.h" not in txt, f


def test_create_validates_arguments():
    lib = D.load()
    h = C.c_void_p()
    for bad in (L.Params(0.0, 1e-8, 1.0, 1, 0, 0, 0), L.Params(1e-2, 1.5, 1.0, 1, 0, 0, 0),
                L.Params(1e-2, 1e-8, 1.0, 7, 0, 0, 0)):
        assert lib.dnv_create(C.byref(bad), C.byref(h)) == D.DNV_E_INVALID
        assert b"dnv_create" in lib.dnv_last_error()
    assert lib.dnv_create(None, C.byref(h)) == D.DNV_E_INVALID


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(D.DenovoError) as ei:
        D.TrioEngine()
    assert ei.value.code == D.DNV_E_CUDA
    assert "no CPU fallback" in str(ei.value)
