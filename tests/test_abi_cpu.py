"""The C-ABI boundary without a GPU: libthr.so loads, exports every symbol include/thr.h declares with the
argument count the ctypes binding uses, and fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from triple_hybrid_rag_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _header_decls():
    text = (ROOT / "include" / "thr.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:int|int64_t|const char\*)\s+(thr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_binding_covers_header_exactly():
    decls = _header_decls()
    assert len(decls) >= 17
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, n in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_library_loads_and_exports_every_symbol():
    if not _lib.LIB_PATH.exists():
        pytest.skip("libthr.so not built (run __graft_entry__.build())")
    lib = _lib.load()
    for name in _header_decls():
        assert hasattr(lib, name), name
    m = re.search(r"#define THR_ABI_VERSION (\d+)", (ROOT / "include" / "thr.h").read_text())
    assert lib.thr_abi_version() == int(m.group(1)) == _lib.ABI_VERSION


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    if not _lib.LIB_PATH.exists():
        pytest.skip("libthr.so not built")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.thr_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.thr_last_error(None)
    from triple_hybrid_rag_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_product_package_never_imports_the_oracle():
    for p in (ROOT / "triple_hybrid_rag_b200").glob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
