"""Host-side rules of the PHP-level API (no GPU needed: the automaton is only uploaded at finalize/match).
Texts and return values follow src/php_ahocorasick.c."""
import warnings

import pytest

from php_aho_corasick_b200.php_api import (AhoException, AhoWarning, ahocorasick_add_patterns, ahocorasick_deinit,
                                           ahocorasick_init, ahocorasick_isValid)


def call(fn, *a):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        r = fn(*a)
    return r, [str(x.message) for x in w if issubclass(x.category, AhoWarning)]


def test_init_accepts_reference_shapes():
    c, w = call(ahocorasick_init, [{"key": "ab", "value": "alfa"}, {"id": 0, "value": "zeta"}, {"value": "lfa"},
                                   {"VALUE": "x", "KEY": "k"}, ["bare"], {"value": "v", "aux": object()}])
    assert c is not False and w == []
    assert ahocorasick_isValid(c) is True
    assert ahocorasick_deinit(c) is True
    assert ahocorasick_isValid(c) is False and ahocorasick_deinit(c) is False     # tests/test2.phpt:295-298


def test_structural_errors_give_warning_and_false():
    r, w = call(ahocorasick_init, ["not-an-array"])
    assert r is False and w == ["Invalid pattern structure! Cannot initialize."]
    r, w = call(ahocorasick_init, [{"value": "a"}, {"bogus": 1, "value": "b"}])
    assert r is False and "unrecognized sub-array key" in w[0] and w[0].endswith("Pattern index: 1")
    r, w = call(ahocorasick_init, [{"key": "k"}])
    assert r is False and w == ["No value was specified for pattern index: 0"]
    r, w = call(ahocorasick_init, [{"key": "k", "id": 3, "value": "v"}])
    assert r is False and w == ["Pattern can have either numeric or string identifier, not both! Pattern index: 0"]
    r, w = call(ahocorasick_init, [{"value": "v", "ignoreCase": True}])
    assert r is not False and w == ["ignoreCase attribute is deprecated and is ignored. Pattern index: 0"]


def test_type_errors_raise_aho_exception():
    with pytest.raises(AhoException, match=r"Invalid type of pattern ID given \(long required\), type: string, pattern index: 0"):
        ahocorasick_init([{"id": "7", "value": "v"}])
    with pytest.raises(AhoException, match=r"Pattern value has to be a string, type: long, pattern index: 1"):
        ahocorasick_init([{"value": "ok"}, {"value": 5}])
    with pytest.raises(AhoException, match=r"Pattern key has to be a string, type: array, pattern index: 0"):
        ahocorasick_init([{"key": [1], "value": "v"}])
    with pytest.raises(AhoException, match=r"type: double"):
        ahocorasick_init([{"id": 1.5, "value": "v"}])
    with pytest.raises(AhoException, match=r"type: null"):
        ahocorasick_init([{"value": None}])


def test_add_patterns_rules_before_finalize():
    c, _ = call(ahocorasick_init, [])
    assert c is not False
    r, w = call(ahocorasick_add_patterns, c, [{"value": "x"}])
    assert r is True and w == []
    r, w = call(ahocorasick_add_patterns, c, [{"value": "ok"}, "bad"])
    assert r is False and w == ["Invalid pattern structure! Cannot initialize."]
    assert ahocorasick_deinit(c) is True
    r, w = call(ahocorasick_add_patterns, c, [{"value": "y"}])
    assert r is False and w == ["Cannot add a new pattern, not initialized"]
