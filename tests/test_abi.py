"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/dr3lk.h
declares, and fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re
import subprocess

import pytest

from _common import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "dr3lk.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dr3lk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(dr3):
    decl = _declared_symbols()
    assert decl == sorted(dr3.SYMBOLS)
    L = dr3.lib()
    for s in decl:
        assert hasattr(L, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", dr3.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dr3lk_[a-z0-9_]+)", out))
    assert exported == set(decl)


def test_level_sizes_follow_opencv_early_stop(dr3):
    assert dr3.lk_level_sizes(1240, 376, (21, 21), 3) == [(1240, 376), (620, 188), (310, 94), (155, 47)]
    assert dr3.lk_level_sizes(1241, 376, (21, 21), 3) == [(1241, 376), (621, 188), (311, 94), (156, 47)]
    # the reference's literal call: 30x30, maxLevel 4 -> level 4 would be 78x24 <= 30 -> effective maxLevel 3
    assert len(dr3.lk_level_sizes(1240, 376, (30, 30), 4)) == 4
    assert len(dr3.lk_level_sizes(3840, 2160, (31, 31), 4)) == 5
    assert dr3.lk_level_sizes(20, 20, (21, 21), 3) == [(20, 20)]
    import oracle
    for (w, h, win, ml) in [(1240, 376, (21, 21), 3), (500, 375, (9, 15), 6), (64, 48, (5, 5), 10), (3840, 2160, (31, 31), 4)]:
        assert dr3.lk_level_sizes(w, h, win, ml) == oracle.lk_level_sizes(w, h, win, ml)


def test_no_cpu_fallback(dr3):
    """Without a CUDA device context creation must fail; with one this test is a no-op."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dr3.Dr3lkError) as e:
        dr3.Context(0)
    assert e.value.code == dr3.E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (or cv2)."""
    pkg = os.path.join(ROOT, "3dr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|cv2)\b", src, flags=re.M), fn
                assert not re.search(r"#\s*include\s*[<\"][^>\"]*oracle", src), fn
                assert "liboracle" not in src and "dlopen" not in src, fn


def test_shard_range_rule_without_a_gpu(dr3):
    """dr3lk_shard_range (the block partition of dr3lk_multi) is host arithmetic: the rule of SURVEY.md 8e and of
    3dr_b200/sharding.py, for every rank of every world size, incl. more ranks than pairs"""
    from importlib import import_module
    sharding = import_module("3dr_b200.sharding")
    for n in (0, 1, 3, 7, 64, 4096, 32767):
        for world in (1, 2, 3, 4, 8, 64):
            blocks = [dr3.shard_range(n, r, world) for r in range(world)]
            assert blocks == [sharding.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(blocks, blocks[1:]))
            for p in range(0, n, max(1, n // 50)):  # pair p lives on rank floor(p * world / n) ... within one pair of rounding
                owner = [r for r, (lo, hi) in enumerate(blocks) if lo <= p < hi]
                assert len(owner) == 1


def test_multi_context_fails_loudly_without_a_device(dr3):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dr3.Dr3lkError) as e:
        dr3.MultiContext([0, 0])
    assert e.value.code == dr3.E_CUDA and "no CPU fallback" in str(e.value)


def test_opencv_typed_shim_compiles(tmp_path):
    """include/dr3lk_opencv.hpp (the two call surfaces with the reference's own cv:: types) and the call-site mirror
    3dr_b200/host/opencv_callsite.cpp compile warning-free against the stand-in OpenCV types, and the header is inert
    when no opencv2/core.hpp is on the include path"""
    inc = os.path.join(ROOT, "include")
    src = os.path.join(ROOT, "3dr_b200", "host", "opencv_callsite.cpp")
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "mock_opencv"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "dr3lk_opencv.hpp"\n#ifdef DR3LK_HAVE_OPENCV\n#error "no OpenCV here"\n#endif\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I", inc, str(probe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
