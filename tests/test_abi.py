"""CPU-only: the C-ABI library builds, loads and exports every symbol include/zippy_b200.h
declares; the product path refuses to run (loudly) without a CUDA device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    from zippy_b200 import _native
    L = _native.lib()
    hdr = open(os.path.join(ROOT, "include", "zippy_b200.h")).read()
    declared = set(re.findall(r"\b(zb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_bounds_and_messages():
    from zippy_b200 import _native
    L = _native.lib()
    assert L.zb200_deflate_bound(0) >= 5
    assert L.zb200_deflate_bound(65536) >= 65536 + 10
    assert L.zb200_compress_bound(100, 2) >= 100 + 5 + 36 + 8
    assert L.zb200_strerror(3) == b"Invalid buffer, unable to uncompress"      # internal.nim:191-192
    assert L.zb200_strerror(14) == b"Checksum verification failed"             # gzip.nim:81
    assert L.zb200_strerror(9) == b"Unable to detect compressed data format"   # zippy.nim:125


def test_no_cpu_fallback():
    import zippy_b200 as z
    from zippy_b200 import _native
    if _native.lib().zb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(z.ZippyError):
        z.Context()
    with pytest.raises(z.ZippyError):
        z.compress(b"hello")


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zippy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def _build_cpp_api_test():
    import subprocess
    src = os.path.join(ROOT, "tests", "native", "cpp_api_test.cpp")
    exe = os.path.join(ROOT, "tests", "native", "cpp_api_test")
    libdir = os.path.join(ROOT, "zippy_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, src, "-L" + libdir, "-l:libzippy_b200.so",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_compiles_and_links():
    # include/zippy_b200.hpp is the compiled-language host side (Nim is unavailable here)
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(_build_cpp_api_test())
