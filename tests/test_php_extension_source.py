"""The real Zend extension source (php/php_ahocorasick_b200.c).  This image has no php-dev, so the file is type-checked
against declaration-only stand-ins of the Zend headers (php/zend_stub/) and compiled to an object whose undefined
symbols must be exactly the Zend API it uses plus the C-ABI of include/acb200.h; php/build.sh builds and runs the
reference's .phpt files where phpize exists (and must say so and exit 0 where it does not)."""
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "php", "php_ahocorasick_b200.c")
INC = ["-I" + os.path.join(ROOT, "php", "zend_stub"), "-I" + os.path.join(ROOT, "include")]


def test_extension_source_type_checks_against_the_zend_api_shape():
    r = subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only",
                        "-DCOMPILE_DL_AHOCORASICK", *INC, SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_extension_object_needs_only_zend_and_the_c_abi(tmp_path):
    obj = str(tmp_path / "ext.o")
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-fPIC", "-DCOMPILE_DL_AHOCORASICK", *INC, "-c", SRC, "-o", obj], check=True)
    syms = subprocess.run(["nm", "-u", obj], capture_output=True, text=True, check=True).stdout.split()
    undefined = {s for s in syms if s != "U"}
    matcher = {s for s in undefined if s.startswith(("ac_trie_", "acb200_"))}
    # the five calls of the reference's glue (src/php_ahocorasick.c:812, 484, 140, 745, 504) + the batch entry + diagnostics
    assert matcher == {"ac_trie_create", "ac_trie_add", "ac_trie_finalize", "ac_trie_search", "ac_trie_release",
                       "ac_trie_search_batch", "acb200_last_error", "acb200_version"}
    header = open(os.path.join(ROOT, "include", "acb200.h")).read()
    for s in matcher:
        assert re.search(r"\b%s\s*\(" % s, header), s
    defined = subprocess.run(["nm", "--defined-only", obj], capture_output=True, text=True, check=True).stdout
    for fn in ("ahocorasick_init", "ahocorasick_add_patterns", "ahocorasick_finalize", "ahocorasick_match",
               "ahocorasick_match_batch", "ahocorasick_isValid", "ahocorasick_deinit"):
        assert "zif_" + fn in defined, fn
    assert "get_module" in defined and "ahocorasick_module_entry" in defined


def test_result_record_and_messages_follow_the_reference():
    """What PHP scripts and the six .phpt files observe: key order pos, key|keyIdx, aux, start_postion (sic), value and
    the warning / exception texts (reference src/php_ahocorasick.c:222-329, 404, 555-584, 697, 702, 910, 915)."""
    text = open(SRC).read()
    order = [text.index(k) for k in ('"pos"', '"key"', '"keyIdx"', '"aux"', '"start_postion"', '"value"')]
    body = text[text.index("static void aho_append_hit"):]
    order = [body.index(k) for k in ('"pos"', '"key"', '"keyIdx"', '"aux"', '"start_postion"', '"value"')]
    assert order == sorted(order)
    for msg in ("Invalid resource.", "Not initialized.", "Invalid pattern structure! Cannot initialize.",
                "No value was specified for pattern index: %ld",
                "Pattern can have either numeric or string identifier, not both! Pattern index: %ld",
                "ignoreCase attribute is deprecated and is ignored. Pattern index: %ld",
                "Cannot add a new pattern to finalized search structure", "Cannot add a new pattern, not initialized",
                "Invalid type of pattern ID given (long required), type: %s, pattern index: %ld",
                "Pattern %s has to be a string, type: %s, pattern index: %ld"):
        assert msg in text, msg


def test_build_script_is_gated_on_the_php_toolchain():
    if shutil.which("php-config") and shutil.which("phpize"):
        import pytest
        pytest.skip("a PHP toolchain is present: run php/build.sh by hand (it needs a GPU for make test)")
    r = subprocess.run(["sh", os.path.join(ROOT, "php", "build.sh")], capture_output=True, text=True)
    assert r.returncode == 0 and "skipping the extension build" in r.stdout
