"""The drop-in boundary is C: include/comorag_b200.h must be a valid C99 and C++ header with no framework types in
it, and a host with no Python / torch in it (examples/c_host_search.cu) must link against the library and use it."""
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "comorag_b200.h")


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c99"), ("g++", "-std=c++11")])
def test_header_is_plain_c_and_cpp(tmp_path, compiler, std):
    if shutil.which(compiler) is None:
        pytest.skip(f"{compiler} not installed")
    ext = "c" if compiler == "gcc" else "cpp"
    src = tmp_path / f"use_header.{ext}"
    # take the address of a few entry points so their prototypes are really parsed and type-checked
    src.write_text('#include "comorag_b200.h"\n'
                   "typedef int (*search_fn)(const void*, int64_t, int, int64_t, int64_t, const void*, int, int, int64_t*,\n"
                   "                         float*, float*, void*, size_t, crag_stream_t);\n"
                   "search_fn f = crag_search_topk;\n"
                   "size_t (*g)(int, int) = crag_search_workspace_bytes;\n"
                   "crag_encoder model;\n"
                   "crag_classifier_head head;\n"
                   "int main(void) { return (f != 0 && g != 0 && sizeof model > 0 && sizeof head > 0) ? CRAG_OK : CRAG_ERR_INVALID; }\n")
    r = subprocess.run([compiler, std, "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I",
                        os.path.dirname(HEADER), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_header_has_no_framework_types():
    text = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)     # comments may mention PyTorch; declarations may not
    for banned in ("torch", "at::", "c10::", "Tensor", "std::", "PyObject", "#include <cuda"):
        assert banned not in code, f"{banned!r} appears in a declaration of the C ABI header"
    assert set(re.findall(r"#include\s+<([^>]+)>", code)) == {"stddef.h", "stdint.h"}


def test_c_host_builds_against_the_library_and_fails_loudly_without_a_gpu():
    from comorag_b200 import build
    build.build()
    exe = build.build_examples()
    assert os.access(exe, os.X_OK)
    needed = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libcomorag_b200.so" in needed and "not found" not in needed
    for lib in ("libtorch", "libpython", "libc10"):
        assert lib not in needed, f"the C host links {lib}"
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is tests/test_z_c_host_gpu.py")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CUDA device" in r.stderr      # no CPU path: refuses, does not fake an answer
