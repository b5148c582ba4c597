"""Host logic of the direct verification of flagged words (no GPU): gram_table.hpp's table + comparison, evaluated
on the host through acb200_direct_probe (the very function the walk kernel runs), must agree with the CPU oracle for
EVERY aligned word of a text — not only the flagged ones:

  verdict 0  -> the oracle has no event at the W end offsets the word owns;
  verdict 1  -> the oracle has exactly one event there, at the reported end, the reported state lists the oracle's
                patterns in the oracle's order, and that state is the pattern's own trie node (the deepest trie
                node that is a suffix of the text there is exactly the pattern — what an automaton walk reaches);
  verdict 2  -> undecided, the kernel walks (must stay the exception for dictionaries without nested patterns).
"""
import random

import numpy as np
import pytest

from php_aho_corasick_b200.native import Automaton
from tests.helpers import oracle_hits


def build_host(pats):
    a = Automaton()
    a.add_php_order(pats)
    a.L.ac_trie_finalize(a.h)            # the host half of finalize runs without a GPU
    return a


def check_text(a, pats, text, expect_mostly_direct):
    inf = a.info()
    W = inf.filter_word
    assert W in (4, 8) and inf.direct_keys > 0
    exp = oracle_hits([pats], [np.frombuffer(text, dtype=np.uint8)])[0]
    by_end = {}
    for p, o in zip(exp[0], exp[1]):
        by_end.setdefault(int(p), []).append(int(o))
    accepted = {}
    for o, p in enumerate(pats):                             # duplicates: the last array entry wins (reverse add order)
        accepted[bytes(p)] = o
    prefixes = {bytes(p[:i]) for p in accepted for i in range(1, len(p) + 1)}
    lmax = max(len(p) for p in accepted)
    counts = {-1: 0, 0: 0, 1: 0, 2: 0}
    events_direct = 0
    for k in range(len(text) // W):
        v, end, state = a.direct_probe(text, k)
        counts[v] += 1
        rs = W * (k + 1)
        owned = [p for p in range(rs + 1, rs + W + 1) if p in by_end]
        if v == 0:
            assert owned == [], (k, owned)
        elif v == 1:
            assert owned == [end], (k, owned, end)
            lst = a.state_patterns(state)
            assert [o for o, _ in lst] == by_end[end], (k, lst, by_end[end])
            o, ln = lst[0]
            assert text[end - ln:end] == bytes(pats[o])
            deepest = max(l for l in range(1, min(lmax, end) + 1) if text[end - l:end] in prefixes)
            assert deepest == ln, (k, deepest, ln)           # the walk's state is the pattern's own node
            events_direct += 1
        elif v == -1:
            warm = -(-(lmax - 1) // W) * W
            assert rs < warm or rs + W > len(text)
    if expect_mostly_direct:
        assert counts[2] * 20 <= counts[0] + counts[1] + counts[2], counts
        assert events_direct >= 10, counts
    return counts


@pytest.mark.parametrize("seed", range(10))
def test_direct_verdicts_match_the_oracle(seed):
    rng = random.Random(100 + seed)
    alphabet = rng.choice([b"ab", b"abc", b"abcdef", bytes(range(256)), b"\x00\xff\x80a"])
    min_len = rng.choice([8, 9, 12, 16, 17, 23, 31])
    pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(min_len, min_len + rng.choice([0, 3, 30]))))
            for _ in range(rng.choice([1, 5, 60]))]
    a = build_host(pats)
    text = bytearray(rng.choice(alphabet) for _ in range(5000))
    for _ in range(60):                                      # occurrences at every alignment, some overlapping
        p = rng.choice(pats)
        at = rng.randint(0, len(text) - len(p))
        text[at:at + len(p)] = p
    text[:len(pats[0])] = pats[0]                            # one at offset 0 (clipped window -> not applicable)
    text[len(text) - len(pats[-1]):] = pats[-1]              # and one on the last byte
    check_text(a, pats, bytes(text), expect_mostly_direct=len(alphabet) > 3)


@pytest.mark.parametrize("W", [4, 8])
def test_edge_lengths_duplicates_and_high_bytes(W):
    """Lengths at the limits (exactly 2W, 2W+1, the tail/store boundary at 16/17 bytes, the reference's maximum of
    1024), a duplicate inside one call (the later array entry wins) and bytes >= 0x80 / NUL."""
    rng = random.Random(4242 + W)
    alphabet = bytes([0x00, 0x7f, 0x80, 0xff, 0x41, 0x42])
    lens = [2 * W, 2 * W + 1, 15 if W == 4 else 23, 16, 17, 24, 31, 32, 33, 64, 100, 1023, 1024]
    pats = [bytes(rng.choice(alphabet) for _ in range(L)) for L in lens if L >= 2 * W]
    pats.append(pats[3])                                     # duplicate value: ordinal of the LAST entry is reported
    a = build_host(pats)
    assert a.info().filter_word == W
    text = bytearray(rng.choice(alphabet) for _ in range(9000))
    at = 40
    for p in pats:                                           # every pattern once, at a different alignment each
        text[at:at + len(p)] = p
        at += len(p) + 37 + (at % 5)
        if at + 1100 > len(text):
            at = 3
    for _ in range(40):
        p = pats[rng.randrange(6)]
        pos = rng.randint(0, len(text) - len(p))
        text[pos:pos + len(p)] = p
    counts = check_text(a, pats, bytes(text), expect_mostly_direct=False)
    assert counts[1] >= 20, counts


def test_nested_and_overlapping_patterns_fall_back_to_the_walk():
    # suffix-nested patterns share grams; a pattern that is the suffix of a longer pattern's PREFIX is a failure
    # target (the walk may sit on a deeper node when it ends) — both must come back as "undecided", never wrong
    pats = [b"0123456789abcdef", b"xx0123456789abcdef", b"456789abcdefghij", b"Q0123456789abcdefZZ",
            b"aaaaaaaaaaaaaaaa", b"aaaaaaaaaaaaaaaaaaaa", b"hello, world....", b"say hello, world....!"]
    a = build_host(pats)
    rng = random.Random(7)
    text = bytearray(rng.choice(b"abcdefx0123456789") for _ in range(4000))
    for i in range(80):
        p = pats[i % len(pats)]
        at = rng.randint(0, len(text) - len(p))
        text[at:at + len(p)] = p
    text[1000:1040] = b"a" * 40
    counts = check_text(a, pats, bytes(text), expect_mostly_direct=False)
    assert counts[2] > 0
    inf = a.info()
    assert 0 < inf.direct_walk_keys <= inf.direct_keys


def test_config2_dictionary_is_fully_direct():
    rng = random.Random(5)
    pats = [bytes(rng.choice(b"abcdef") for _ in range(16)) for _ in range(2048)]
    a = build_host(pats)
    inf = a.info()
    assert inf.filter_word == 8
    assert inf.direct_keys + inf.direct_walk_keys >= 2048 * 8 - 64      # nearly all grams distinct
    assert inf.direct_walk_keys < 64
    text = bytearray(rng.choice(b"abcdef") for _ in range(20000))
    for i in range(200):
        p = pats[rng.randrange(len(pats))]
        at = rng.randint(0, len(text) - 16)
        text[at:at + 16] = p
    counts = check_text(a, pats, bytes(text), expect_mostly_direct=True)
    assert counts[1] >= 150


@pytest.mark.parametrize("seed", range(4))
def test_haystack_start_inside_the_window(seed):
    """A haystack that starts inside a word's warm-up window: the candidate must fit into the haystack (bytes before
    hay_begin belong to a neighbour).  Oracle = the reference run on the haystack alone."""
    rng = random.Random(900 + seed)
    W = rng.choice([4, 8])
    lens = (16, 18, 23, 40) if W == 8 else (8, 9, 13, 15, 30)
    pats = [bytes(rng.choice(b"abcdef") for _ in range(rng.choice(lens))) for _ in range(40)]
    a = build_host(pats)
    assert a.info().filter_word == W
    lmax = max(len(p) for p in pats)
    warm = -(-(lmax - 1) // W) * W
    text = bytearray(rng.choice(b"abcdef") for _ in range(3000))
    starts = sorted(rng.sample(range(100, 2900), 30))
    for hb in starts:                                        # occurrences right at, just after and across each start
        p = rng.choice(pats)
        at = hb + rng.choice([0, 0, 1, 2, W - 1, -1, -3, -len(p) + 1])
        text[at:at + len(p)] = p
    text = bytes(text)
    n_events = 0
    for hb in starts:
        exp = oracle_hits([pats], [np.frombuffer(text[hb:], dtype=np.uint8)])[0]
        ends = {}
        for p, o in zip(exp[0], exp[1]):
            ends.setdefault(int(p) + hb, []).append(int(o))
        for k in range(hb // W, min(len(text) // W - 1, (hb + 3 * lmax) // W)):
            rs = W * (k + 1)
            if rs < hb or rs < warm:
                continue
            v, end, state = a.direct_probe(text, k, hb)
            owned = [p for p in range(rs + 1, rs + W + 1) if p in ends]
            if v == 0:
                assert owned == [], (hb, k, owned)
            elif v == 1:
                assert owned == [end], (hb, k, owned, end)
                assert [o for o, _ in a.state_patterns(state)] == ends[end]
                n_events += 1
            else:
                assert v == 2
    assert n_events >= 5


def test_no_table_without_prefilter():
    b = build_host([b"abc", b"abcdefghijklmnopq"])
    assert b.info().direct_keys == 0
    assert b.direct_probe(b"x" * 64, 3)[0] == -1


def test_property_random_dictionaries_with_heavy_overlap():
    """hypothesis: tiny alphabets and patterns cut from one another (shared grams, suffix / prefix nesting, failure
    targets everywhere) — whatever the table decides directly must be what the oracle reports."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, 2 ** 32 - 1), st.sampled_from([4, 8]), st.sampled_from([b"ab", b"abc", b"\x00\xff"]))
    def run(seed, W, alphabet):
        rng = random.Random(seed)
        base = bytes(rng.choice(alphabet) for _ in range(120))
        pats = []
        for _ in range(rng.randint(1, 12)):                  # substrings of one text: maximal overlap between patterns
            L = rng.randint(2 * W, 2 * W + rng.choice([0, 1, 5, 20]))
            at = rng.randint(0, len(base) - L)
            pats.append(base[at:at + L])
        a = build_host(pats)
        if a.info().filter_word != W:                        # (W = 8 needs every pattern >= 16 bytes: true by construction)
            assert W == 4 and min(len(p) for p in pats) >= 16
        text = bytearray(rng.choice(alphabet) for _ in range(600))
        for _ in range(8):
            at = rng.randint(0, len(text) - len(base))
            cut = rng.randint(20, len(base))
            text[at:at + cut] = base[:cut]
        check_text(a, pats, bytes(text), expect_mostly_direct=False)
        a.release()

    run()
