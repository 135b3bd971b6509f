"""The C-ABI library loads and exports every symbol include/zkb.h declares; without a GPU the product
refuses to run (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

from zokrates_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zkb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_product_library_exports_all_symbols():
    assert os.path.exists(_lib.DEFAULT_LIB), "run __graft_entry__.build() first"
    lib = _lib.Library(_lib.DEFAULT_LIB)
    for sym in header_symbols():
        assert hasattr(lib.dll, sym), sym
    assert lib.dll.zkb_abi_version() == 1
    assert lib.curve_sizes(0) == [32, 32, 256, lib.curve_sizes(0)[3]]
    assert lib.curve_sizes(1)[:3] == [32, 48, 384]
    with pytest.raises(_lib.ZkbError):
        lib.curve_sizes(7)


def test_no_cpu_fallback_without_gpu():
    lib = _lib.Library(_lib.DEFAULT_LIB)
    if lib.dll.zkb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.ZkbError) as e:
        _lib.Context(0, 0, lib)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(_lib.ZkbError):
        _lib.Library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "zokrates_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libzkoracle" not in src and "zkoracle.c" not in src, f
    for f in os.listdir(os.path.join(ROOT, "tools")):          # the file-level front doors are product code too
        if f.startswith("zkb_") and f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f
