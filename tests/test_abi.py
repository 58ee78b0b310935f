"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/orb_b200.h declares; without a GPU the product fails loudly
instead of falling back to a CPU path."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from multi_orb_slam_b200 import build
    build.build()
    from multi_orb_slam_b200 import _lib
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(orb[xmdp]_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib.lib, name), f"{name} declared in orb_b200.h but not exported"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)


def test_struct_layouts_match_header(lib):
    import ctypes as C
    assert lib.KP_DTYPE.itemsize == 24 and lib.MP_DTYPE.itemsize == 28
    assert C.sizeof(lib.Config) == 36 and C.sizeof(lib.Bounds) == 16


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multi_orb_slam_b200.extractor import ORBextractor
    from multi_orb_slam_b200.matcher import ORBmatcher
    with pytest.raises(lib.OrbError):
        ORBextractor(1000, 1.2, 8, 20, 7, image_size=(640, 480))
    with pytest.raises(lib.OrbError):
        ORBmatcher(0.9, True)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multi_orb_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in src and "liborb_oracle" not in src and "liborb_ref" not in src, f
                assert not re.search(r'#include\s+"[^"]*oracle/', src), f
