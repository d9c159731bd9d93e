"""-m "not gpu": the C-ABI library builds, loads, and exports every symbol include/spvipes_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from spvipes_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "spvipes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(spv_\w+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spvipes_b200.h but not exported"


def test_binding_table_matches_header(lib_path):
    from spvipes_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared()
    src = open(os.path.join(ROOT, "include", "spvipes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in _lib._SIGS.items():
        m = re.search(r"\b(?:int|long long)\s+" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        decl = m.group(1).strip()
        n = 0 if decl in ("", "void") else len(decl.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, binding has {len(args)}"


def test_abi_version(lib_path):
    from spvipes_b200 import _lib
    assert _lib.load().spv_abi_version() == 1
