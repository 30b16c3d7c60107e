"""The search kernel's selector primitives on the CPU: csrc/topk.cuh and csrc/pool_floor.cuh are pure SIMT code (loads,
compares, shuffles, warp reductions), so the SAME headers the kernel includes are compiled for the host with the warp
intrinsics mapped onto 32 cooperatively scheduled lanes (tests/warp_emu) and checked against std::sort models: the
bitonic sort, the flush / insert list maintenance, and the pooled admission floors -- above all the property exactness
rests on, that at least k published keys reach the floor a refresh returns (tests/warp_emu/selector_emu_test.cpp)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "warp_emu")
CSRC = os.path.join(ROOT, "comorag_b200", "csrc")


@pytest.fixture(scope="module")
def emu_binary(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not installed")
    exe = tmp_path_factory.mktemp("warp_emu") / "selector_emu_test"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unknown-pragmas", "-I", os.path.join(EMU, "stub"), "-I", CSRC,
                        os.path.join(EMU, "selector_emu_test.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_selector_primitives_match_their_models_on_emulated_lanes(emu_binary):
    r = subprocess.run([str(emu_binary), "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ALL OK")
    for group in ("warp_sort_desc<8>", "flush_query<128, 128>", "first-tile flush_query<128, 0>", "insert_few<128, 128>",
                  "select_stream<128, 128>", "lane_kth_of_pool<4>", "pooled_floor_batch8", "pooled_kth_key / pooled_max_kth"):
        assert f"ok  {group}" in r.stdout, group


def test_the_emulated_headers_are_the_ones_the_kernel_includes():
    """No copy of the selector lives under tests/: the emulation includes csrc/topk.cuh and csrc/pool_floor.cuh, and
    search.cu includes the same two files."""
    test_src = open(os.path.join(EMU, "selector_emu_test.cpp")).read()
    assert '#include "topk.cuh"' in test_src and '#include "pool_floor.cuh"' in test_src
    kernel_src = open(os.path.join(CSRC, "search.cu")).read()
    assert '#include "topk.cuh"' in kernel_src and '#include "pool_floor.cuh"' in kernel_src
    for name in os.listdir(EMU):
        assert not name.endswith(".cuh"), f"{name}: kernel headers must not be duplicated under tests/"


def test_emulation_catches_a_broken_floor(emu_binary, tmp_path):
    """Mutation check: a floor bisection that counts `>` instead of `>=` (a floor one key too high -- it would drop a
    true top-k row) must fail the property test."""
    mutated = tmp_path / "csrc"
    mutated.mkdir()
    for h in ("topk.cuh", "pool_floor.cuh"):
        shutil.copy(os.path.join(CSRC, h), mutated / h)
    src = (mutated / "pool_floor.cuh").read_text()
    needle = "for (int i = 0; i < NC; ++i) c += (hi[j][i] >= cand) ? 1 : 0;"
    assert src.count(needle) == 1
    (mutated / "pool_floor.cuh").write_text(src.replace(needle, needle.replace(">= cand", "> cand")))
    exe = tmp_path / "mutant"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wno-unknown-pragmas", "-I", os.path.join(EMU, "stub"), "-I", str(mutated),
                        os.path.join(EMU, "selector_emu_test.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "FAILED" in r.stderr
