"""The C-ABI shared library loads and exports exactly what include/e4s_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "e4s_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(e4s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from e4s_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/e4s_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes table and header disagree"


def test_version_and_arch_strings():
    from e4s_b200 import _lib
    lib = _lib.load()
    assert lib.e4s_version() >= 100
    assert lib.e4s_build_arch() == b"sm_100a"


def test_sass_is_sm100a_only():
    import subprocess
    from e4s_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
