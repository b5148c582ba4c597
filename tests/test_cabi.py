"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the headers declare,
mirrors the reference's struct layouts, and fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from php_aho_corasick_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text))
    not_functions = {"int", "void", "char", "long", "double", "size_t", "uint32_t", "uint64_t", "defined"}
    return {n for n in names if not n.startswith("AC_") and not n.endswith("_f") and n not in not_functions}


def test_library_exports_every_declared_symbol():
    L = native.lib()
    decl = declared_functions("acb200.h") | declared_functions("acb200_php.h")
    assert set(native.EXPORTS) <= decl
    for name in sorted(decl):
        assert hasattr(L, name), f"libacb200.so does not export {name}"
    assert len(decl) >= 28


def test_struct_layouts_match_the_reference_abi():
    # src/multifast/actypes.h:47-113 on LP64
    assert C.sizeof(native.AcText) == 16
    assert C.sizeof(native.AcPattId) == 16
    assert C.sizeof(native.AcPattern) == 56
    assert C.sizeof(native.AcMatch) == 24
    assert native.AcPattern.aux.offset == 48 and native.AcPattern.id.offset == 32
    assert C.sizeof(native.Event) == 16


def test_drivers_built_from_one_source_link_against_both_apis():
    from oracle import pydriver
    assert pydriver.available("oracle")
    assert pydriver.available("gpu"), "oracle/libgpu_driver.so (driver.c linked to libacb200.so) missing"


def test_no_gpu_means_loud_failure_not_fallback():
    L = native.lib()
    if L.acb200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    a = native.Automaton()
    assert a.add(b"abc") == 0
    with pytest.raises(native.AcError) as e:
        a.finalize()
    assert "CUDA" in str(e.value) or "device" in str(e.value)
    rc, got = a.search_callback(b"zabc")
    assert rc == -1 and got == []
    with pytest.raises(native.AcError):
        a.search_events(b"zabc")


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "php_aho_corasick_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                for needle in ("import oracle", "from oracle", "liboracle", "libref_driver", "oracle/_ref", "pydriver"):
                    assert needle not in src, f"{f} references the checker ({needle})"
