"""The C-ABI library loads, exports every symbol include/b200lidar.h declares, and the ctypes
prototypes in lidarcrafter_b200/_lib.py agree with the header (argument counts)."""
import ctypes
import os
import re

import pytest

from lidarcrafter_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200lidar.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int|size_t|const char\*)\s+(b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_header_declares_the_expected_entry_points():
    fns = header_functions()
    assert len(fns) >= 20
    assert set(fns) == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name


def test_ctypes_prototypes_match_header_arity():
    for name, n in header_functions().items():
        assert len(_lib.PROTOTYPES[name][1]) == n, name


def test_loader_fails_loudly_without_the_library(tmp_path):
    with pytest.raises(_lib.B200LidarError):
        _lib.Lib(str(tmp_path / "missing.so"))


def test_cpu_tensor_is_rejected_by_the_product_path():
    import torch
    from helpers import make_unet
    m, _ = make_unet((8, 1024), (1, 1, 1, 1))
    assert _lib._TEST_LIB is None
    with pytest.raises(_lib.B200LidarError):
        m(torch.zeros(1, 2, 8, 1024), torch.zeros(1))


def test_version_call_needs_no_gpu():
    assert _lib.Lib().cdll.b200_version() >= 100
