"""Device-side hit expansion (acb200_search_hits): events -> {haystack, pos, start_postion, pattern} records, as the
reference's callback builds them (src/php_ahocorasick.c:555-584), against the CPU oracle."""
import random

import numpy as np
import pytest

from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
from tests.helpers import oracle_hits, split

pytestmark = pytest.mark.gpu


def build(pattern_calls):
    a = Automaton()
    for call in pattern_calls:
        a.add_php_order(call)
    a.finalize()
    return a


def check(a, patterns_calls, hays, flat, off):
    hits = a.search_hits(flat, off)
    exp = oracle_hits(patterns_calls, hays)
    all_pats = [p for call in patterns_calls for p in call]
    ordinal = np.array([a.pattern_ordinal(i) for i in range(a.info().n_patterns)], dtype=np.int64)
    k = 0
    for h, (epos, epat, _nev, _hash) in enumerate(exp):
        n = len(epos)
        got = hits[k:k + n]
        assert np.all(got["text_idx"] == h), h
        assert np.array_equal(got["end"].astype(np.uint64), epos), h
        assert np.array_equal(ordinal[got["pattern"]], epat.astype(np.int64)), h
        lens = np.array([len(all_pats[o]) for o in epat], dtype=np.uint32)
        assert np.array_equal(got["start"], got["end"] - lens), h
        k += n
    assert k == len(hits)
    return hits


def test_readme_vector_hits():
    pats = [p["value"].encode() for p in W.CFG1_PATTERNS]
    a = build([pats])
    hay = np.frombuffer(W.CFG1_HAYSTACK, dtype=np.uint8)
    hits = check(a, [pats], [hay], hay, np.array([0, hay.size], dtype=np.uint64))
    # tests/test1.phpt:60-119 — pos 14,19,24,28,28 ; start_postion 9,15,19,24,25 ; 'alfa' before 'lfa'
    assert hits["end"].tolist() == [14, 19, 24, 28, 28]
    assert hits["start"].tolist() == [9, 15, 19, 24, 25]


def test_cfg2_planted_hits_filtered_and_full_walk():
    needles, hay, off = W.cfg2()
    a = build([needles])
    for mode in (1, -1):
        a.set_filter(mode)
        hits = check(a, [needles], split(hay, off), hay, off)
        assert len(hits) >= 256 * 7 and a.stats().expand_ms > 0


@pytest.mark.parametrize("seed", range(3))
def test_nested_patterns_many_hits_per_event_ragged(seed):
    rng = random.Random(seed)
    for trial in range(8):
        alpha = rng.choice([b"ab", b"abc", b"a", b"\x00\xff\x80a"])
        n = rng.randint(1, 60)
        pats = [bytes(rng.choice(alpha) for _ in range(rng.randint(1, 9))) for _ in range(n)]
        lens = [rng.choice([0, 1, 5, 64, 700, 3000]) for _ in range(rng.randint(1, 9))]
        hays = [np.frombuffer(bytes(rng.choice(alpha) for _ in range(m)), dtype=np.uint8) for m in lens]
        flat = np.concatenate(hays) if sum(lens) else np.zeros(0, np.uint8)
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        half = len(pats) // 2
        a = build([pats[:half], pats[half:]])
        check(a, [pats[:half], pats[half:]], hays, flat, off)
        a.release()


def test_cfg5_shape_expansion_total():
    n = 8192
    pats, hay, off = W.cfg5(n_patterns=1100, hay_bytes=n)
    a = build([pats])
    hits = a.search_hits(hay, off)
    events, total = W.cfg5_expected(n, 1100)
    assert len(hits) == total
    # event at offset p reports a^min(p,1024) .. a^1, longest first
    p = 3000
    seg = hits[hits["end"] == p]
    assert (seg["end"] - seg["start"]).tolist() == list(range(1024, 0, -1))
