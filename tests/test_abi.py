"""CPU-only: the C-ABI shared library loads, exports every symbol include/dadetect_b200.h declares, and
the ctypes signature table (da-detect_b200/_lib.py) agrees with the header prototypes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "dadetect_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|long long|size_t)\s+(dd_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        sig = ""
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    sig += "p"
                elif a.startswith("float"):
                    sig += "f"
                elif a.startswith("double"):
                    sig += "d"
                elif a.startswith("long long"):
                    sig += "q"
                elif a.startswith("int"):
                    sig += "i"
                else:
                    raise AssertionError("unhandled parameter type: " + a)
        protos[name] = (ret, sig)
    return protos


def test_header_matches_ctypes_table():
    from dadetect_b200 import _lib
    protos = header_prototypes()
    assert set(protos) == set(_lib._SIGS), set(protos) ^ set(_lib._SIGS)
    for name, (ret, sig) in protos.items():
        assert _lib._SIGS[name][1] == sig, (name, _lib._SIGS[name][1], sig)


def test_library_loads_and_exports_every_symbol():
    from dadetect_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_prototypes():
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.dd_abi_version() == 1
    assert loaded.dd_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch):
    from dadetect_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdadetect_b200.so")
    with pytest.raises(ImportError):
        _lib.load()
