"""The C-ABI library loads and exports every symbol include/mesa_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mesa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from mesa_b200 import build

    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert "mb_gol_step_u8" in names and "mb_version" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mesa_b200.h but not exported"


def test_binding_table_matches_header():
    from mesa_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared()
    h = _native.lib()
    assert h.mb_version() >= 100
    assert _native.last_error() == "" or isinstance(_native.last_error(), str)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mesa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
