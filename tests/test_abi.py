"""CPU test: libc3d.so loads and exports every symbol include/c3d.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for f in os.listdir(os.path.join(ROOT, "include")):
        if f.endswith(".h"):
            src = open(os.path.join(ROOT, "include", f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(c3d_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from omni3d_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 6
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/*.h but not exported"
    assert sorted(_lib.EXPORTS) == decl, "omni3d_b200/_lib.py EXPORTS out of sync with include/c3d.h"
    L.c3d_abi_version.restype = ctypes.c_int32
    assert L.c3d_abi_version() >= 1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from omni3d_b200 import box3d, _lib
    with pytest.raises(_lib.C3DError):
        box3d.box3d_overlap(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "omni3d_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\".*oracle", s, flags=re.M):
                    bad.append(f)
    assert not bad, bad
