"""The reference's six PHPT tests, replayed through the mirrored PHP API on the GPU and compared
record by record (key order included) with the golden var_dump output parsed from the reference."""
import warnings

import pytest

from php_aho_corasick_b200.php_api import (ahocorasick_add_patterns, ahocorasick_deinit, ahocorasick_finalize,
                                           ahocorasick_init, ahocorasick_isValid, ahocorasick_match,
                                           ahocorasick_match_batch)
from tests.helpers import golden_record, load_golden

pytestmark = pytest.mark.gpu
G = load_golden()


def check(got, expect):
    assert got is not False
    assert len(got) == len(expect)
    for g, e in zip(got, expect):
        order, rec = golden_record(e)
        assert list(g.keys()) == order
        assert g == rec


def open_case(case):
    c = ahocorasick_init(case["init"])
    assert c is not False
    for call in case.get("add_patterns", []):
        assert ahocorasick_add_patterns(c, call) is True
    return c


@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test6"])
def test_phpt_golden_outputs(name):
    for case in G[name]["cases"]:
        c = open_case(case)
        for m in case["matches"]:
            if "findAll" in m:
                got = ahocorasick_match(m["haystack"], c, m["findAll"])
            else:
                got = ahocorasick_match(m["haystack"], c)
            check(got, m["expect"])
        # the batched entry point returns the same arrays in one launch
        batch = ahocorasick_match_batch([m["haystack"] for m in case["matches"] if m.get("findAll", True)], c)
        for got, m in zip(batch, [m for m in case["matches"] if m.get("findAll", True)]):
            check(got, m["expect"])
        life = case.get("lifecycle")
        if life:
            assert ahocorasick_isValid(c) is life["isValid"]
            assert ahocorasick_deinit(c) is life["deinit"]
            assert ahocorasick_isValid(c) is life["isValid_after"]
            assert ahocorasick_deinit(c) is life["deinit_again"]
        else:
            assert ahocorasick_deinit(c) is True


def test_phpt4_no_state_leaks_over_many_handles_and_calls():
    t = G["test4"]
    for _ in range(t["outer"]):
        c = ahocorasick_init(t["init"])
        for _ in range(t["inner"] // 10):          # 20 x 100 matches keeps the GPU suite short
            d = ahocorasick_match(t["haystack"], c)
            assert d and len(d) == t["hits"]
        assert ahocorasick_deinit(c) is True


def test_phpt5_multibyte_haystacks_without_hits():
    t = G["test5"]
    c = ahocorasick_init(t["init"])
    for h in t["haystacks"]:
        assert ahocorasick_match(h, c) == []
    assert ahocorasick_match_batch(t["haystacks"], c) == [[] for _ in t["haystacks"]]
    ahocorasick_deinit(c)


def test_find_all_false_returns_all_patterns_of_the_first_event():
    # src/php_ahocorasick.c:588 — the callback stops after the first EVENT, which may carry several patterns
    c = ahocorasick_init([{"value": "alfa"}, {"value": "lfa"}, {"value": "a"}, {"value": "zz"}])
    got = ahocorasick_match("xxzzalfa zz", c, False)
    assert [(g["pos"], g["value"]) for g in got] == [(4, "zz")]
    got = ahocorasick_match("alfa zz", c, False)
    assert [(g["pos"], g["value"]) for g in got] == [(1, "a")]
    got = ahocorasick_match("xlfa", c, False)
    assert [(g["pos"], g["value"]) for g in got] == [(4, "lfa"), (4, "a")]
    assert ahocorasick_match("", c, False) == [] and ahocorasick_match("qqq", c) == []


def test_finalize_lifecycle_and_late_add():
    c = ahocorasick_init([{"value": "ab"}])
    assert ahocorasick_finalize(c) is True and ahocorasick_finalize(c) is False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert ahocorasick_add_patterns(c, [{"value": "cd"}]) is False
    assert [str(x.message) for x in w] == ["Cannot add a new pattern to finalized search structure"]
    assert [g["pos"] for g in ahocorasick_match("abcdab", c)] == [2, 6]
    assert ahocorasick_deinit(c) is True
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert ahocorasick_match("ab", c) is False
    assert [str(x.message) for x in w] == ["Invalid resource."]


def test_binary_safe_patterns_and_haystacks():
    c = ahocorasick_init([{"id": 1, "value": b"\x00\xff\x00"}, {"id": 2, "value": b"\xff"}, {"id": 3, "value": b"a\x00b"}])
    got = ahocorasick_match(b"\x00\xff\x00a\x00b\xff", c)
    assert [(g["pos"], g["keyIdx"]) for g in got] == [(2, 2), (3, 1), (6, 3), (7, 2)]
