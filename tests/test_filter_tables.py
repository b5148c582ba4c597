"""Host logic of the gram prefilter (no GPU): the tables ac_trie_finalize builds must never filter an occurrence away.
For every event of the CPU oracle, the aligned word that owns its end offset — and the byte after that word — must
pass the host-side evaluation of the filter decision (acb200_filter_probe, the same hashes the kernel uses), both with
the next byte known and with "next byte unknown" (the last word of a 512-byte span)."""
import random

import numpy as np
import pytest

from php_aho_corasick_b200.native import Automaton
from tests.helpers import oracle_hits

UNKNOWN = 0x100


def build_host(pats):
    a = Automaton()
    a.add_php_order(pats)
    a.L.ac_trie_finalize(a.h)            # the host half of finalize runs without a GPU
    return a


@pytest.mark.parametrize("seed", range(8))
def test_no_occurrence_is_filtered_away(seed):
    rng = random.Random(seed)
    alphabet = rng.choice([b"ab", b"abc", b"abcdef", bytes(range(256)), b"\x00\xff\x80a"])
    min_len = rng.choice([8, 9, 15, 16, 17, 31])
    pats = [bytes(rng.choice(alphabet) for _ in range(rng.randint(min_len, min_len + rng.choice([0, 3, 30]))))
            for _ in range(rng.choice([1, 5, 60]))]
    a = build_host(pats)
    W = a.info().filter_word
    assert W == (8 if min(len(p) for p in pats) >= 16 else 4)
    text = bytearray(rng.choice(alphabet) for _ in range(6000))
    for _ in range(40):                                     # plant occurrences at every alignment
        p = rng.choice(pats)
        at = rng.randint(0, len(text) - len(p))
        text[at:at + len(p)] = p
    text = bytes(text)
    exp = oracle_hits([pats], [np.frombuffer(text, dtype=np.uint8)])[0]
    ends = sorted(set(int(e) for e in exp[0]))
    assert len(ends) >= 20
    for p_end in ends:
        k = -(-p_end // W) - 2                               # the word with W(k+1) < p_end <= W(k+2)
        assert k >= 0
        word = int.from_bytes(text[W * k:W * k + W], "little")
        nb = text[W * (k + 1)]
        assert a.filter_probe(word, nb) == 1, (p_end, k)
        assert a.filter_probe(word, UNKNOWN) == 1, (p_end, k)


def test_filter_is_selective_and_absent_for_short_patterns():
    rng = random.Random(5)
    pats = [bytes(rng.choice(b"abcdef") for _ in range(16)) for _ in range(2048)]
    a = build_host(pats)
    inf = a.info()
    assert inf.filter_word == 8 and 0.01 < inf.filter_l1_fill < 0.08
    passed = sum(a.filter_probe(rng.getrandbits(64), rng.getrandbits(8)) for _ in range(20000))
    assert passed < 200                                      # random words: ~0.2 % false positives
    b = build_host([b"abc", b"abcdefghijklmnopq"])
    assert b.info().filter_word == 0 and b.filter_probe(0, 0) == -1
