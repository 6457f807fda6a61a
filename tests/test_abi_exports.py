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


def test_java_binding_only_names_exported_symbols_and_matching_offsets():
    """The Java (Panama FFM) sources cannot be compiled here (no JDK): at least every downcall they name must be
    an exported, declared entry point, and the byte offsets they hard-code must be the header's."""
    jdir = os.path.join(ROOT, "denovo-variant-caller-java_b200", "java")
    names, srcs = set(), {}
    for dp, _, fs in os.walk(jdir):
        for f in fs:
            if f.endswith(".java"):
                txt = open(os.path.join(dp, f)).read()
                srcs[f] = txt
                names |= set(re.findall(r'handle\("(dnv_[a-z0-9_]+)"', txt))
    assert {"dnv_group_create", "dnv_group_infer_counts", "dnv_pileup", "dnv_mendel_filter", "dnv_alloc_pinned"} <= names
    declared = set(header_functions())
    assert names <= declared, names - declared
    c, p, v = D.CALL_DT.fields, D.PCALL_DT.fields, D.VCALL_DT.fields
    assert (c["site"][1], c["counts"][1], c["dad"][1], c["mom"][1], c["child"][1]) == (0, 8, 32, 33, 34)   # GpuReadCaller.details
    assert (p["position"][1], p["child"][1], p["mom"][1], p["dad"][1], p["is_denovo"][1]) == (0, 16, 17, 18, 19)  # callShard
    assert (v["start"][1], v["end"][1], v["variant_id"][1], v["gt0"][1], v["gt1"][1], v["ngt"][1], v["person"][1],
            v["filter_pass"][1], v["ref_len"][1], v["ref_base"][1], v["alt0_len"][1], v["alt0_base"][1],
            v["alt1_len"][1], v["alt1_base"][1]) == (0, 8, 16, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30)    # pack()
    assert "extends DenovoCaller" in srcs["GpuReadCaller.java"] and "extends DenovoCaller" in srcs["GpuVariantCaller.java"]
    assert "public void execute()" in srcs["GpuVariantCaller.java"]
