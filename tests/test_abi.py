"""The C-ABI library loads and exports every symbol include/dvis_b200.h declares; the ctypes binding mirrors the
header one to one (argument counts).  No compute calls: runs without a GPU."""
import ctypes
import os
import re

from dvis_plus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dvis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(dvis_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if args == ["void"] else args
    return out


def test_library_exports_every_declared_symbol():
    fns = header_functions()
    assert {"dvis_msda_forward", "dvis_msda_backward", "dvis_msda_fused_forward", "dvis_mask_logits",
            "dvis_add_layernorm", "dvis_abi_version", "dvis_last_error"} <= set(fns)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in include/dvis_b200.h but not exported"


def test_binding_matches_header_arity():
    fns = header_functions()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in fns, name
        assert len(argtypes) == len(fns[name]), (name, len(argtypes), len(fns[name]))
    for name in fns:
        if name not in ("dvis_abi_version", "dvis_last_error"):
            assert name in _lib.SIGNATURES, f"{name} has no ctypes binding"


def test_abi_version_and_error_string():
    lib = _lib.lib()
    assert lib.dvis_abi_version() == _lib.ABI_VERSION
    # argument validation happens on the host before any launch: a null pointer is rejected without a GPU
    rc = lib.dvis_msda_forward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, None, None, None)
    assert rc == 1 and b"null" in lib.dvis_last_error()
    rc = lib.dvis_mask_logits(None, None, 1, 1, 64, 8, None, 0, None)
    assert rc == 1


def test_missing_library_fails_loudly(monkeypatch):
    import importlib
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdvis_b200.so")
    monkeypatch.setattr(_lib, "_lib", None)
    try:
        _lib.lib()
        raise AssertionError("expected RuntimeError")
    except RuntimeError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    importlib.reload(_lib)
