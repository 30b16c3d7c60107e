"""The C-ABI library builds, loads, and exports every symbol include/comorag_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "comorag_b200.h")).read()
    return sorted(set(re.findall(r"CRAG_API\s+[\w\s\*]+?\b(crag_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("crag_search_topk", "crag_merge_topk", "crag_encoder_forward", "crag_gemm_bf16", "crag_pool_normalize",
                 "crag_search_workspace_bytes", "crag_version", "crag_last_error"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    from comorag_b200 import build
    path = build.build()
    lib = ctypes.CDLL(str(path))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported by {path}"


def test_ctypes_table_matches_header():
    from comorag_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()
    lib = _native.load()
    assert lib.crag_version() >= 1000
    assert lib.crag_last_error() is not None


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "comorag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
