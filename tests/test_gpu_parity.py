"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle, bit-exact, order included."""
import random

import numpy as np
import pytest

from oracle.pydriver import Driver, flatten
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
from tests.helpers import assert_same, oracle_hits, split

pytestmark = pytest.mark.gpu


def build(pattern_calls):
    a = Automaton()
    for call in pattern_calls:
        a.add_php_order(call)
    a.finalize()
    return a


def test_cfg1_readme_vector_through_reference_style_driver():
    pats = [p["value"].encode() for p in W.CFG1_PATTERNS]
    res = {}
    for kind in ("oracle", "gpu"):
        d = Driver(kind)
        d.add_php_order(pats)
        d.finalize()
        r = d.search(W.CFG1_HAYSTACK)
        res[kind] = (r["rc"], r["n_events"], r["hash"], r["pos"].tolist(), r["pat"].tolist(), r["len"].tolist())
        d.release()
    assert res["gpu"] == res["oracle"]
    # tests/test1.phpt:60-119 — pos 14,19,24,28,28 ; 'alfa' before 'lfa' at 28
    assert res["gpu"][3] == [14, 19, 24, 28, 28]
    assert res["gpu"][4] == [2, 4, 5, 0, 6]


@pytest.mark.parametrize("seed", range(6))
def test_random_small_dictionaries_callback_path(seed):
    rng = random.Random(seed)
    for trial in range(40):
        alpha = rng.choice([b"ab", b"abc", bytes(range(256)), b"a", b"\x00\xff\x80a", b"abcdef"])
        n = rng.randint(0, 40)
        pats = [bytes(rng.choice(alpha) for _ in range(rng.randint(0, 9))) for _ in range(n)]
        text = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 700)))
        first = trial % 4 == 3
        got = {}
        for kind in ("oracle", "gpu"):
            d = Driver(kind)
            half = len(pats) // 2
            d.add_php_order(pats[:half])
            d.add_php_order(pats[half:])
            d.finalize()
            r = d.search(text, first_only=first)
            got[kind] = (r["rc"], r["n_events"], r["hash"], r["pos"].tolist(), r["pat"].tolist())
            d.release()
        assert got["gpu"] == got["oracle"], (seed, trial, pats, text)


def test_cfg2_benchmark_shape_planted():
    needles, hay, off = W.cfg2()
    a = build([needles])
    ev = a.search_events(hay, off)
    exp = oracle_hits([needles], split(hay, off))
    assert sum(e[2] for e in exp) >= 256 * 7     # planted needles may overwrite each other
    assert_same(a, ev, 256, exp)
    st = a.stats()
    assert st.kernel_launches >= 1 and st.bytes == hay.size


@pytest.mark.parametrize("chunk,smem", [(16, 0), (48, 4096), (256, 0), (1024, 512), (4096, 0)])
def test_slice_and_table_placement_do_not_change_results(chunk, smem):
    needles, hay, off = W.cfg2(n_hay=16, hay_len=3000, n_needles=300, planted_per_hay=6, seed=11)
    a = build([needles])
    a.set_tuning(chunk, smem)
    ev = a.search_events(hay, off)
    assert_same(a, ev, 16, oracle_hits([needles], split(hay, off)))


@pytest.mark.parametrize("seed", range(5))
def test_overlapping_dictionaries_through_the_batch_walk_with_forced_windows(seed):
    """ac_scan_kernel / ac_scan_tma_kernel (batches take them, single short texts do not): dictionaries over tiny
    alphabets whose patterns nest and overlap — every few bytes an event, walks that leave a forced, tiny
    shared-memory window (sink row) and come back inside one 16-byte group — against the reference, for several
    window sizes, slice lengths and both ways of fetching the text."""
    rng = random.Random(1000 + seed)
    nrng = np.random.default_rng(1000 + seed)
    alpha = rng.choice([b"ab", b"abc", b"a\x00\xff", b"abcdef"])
    pats = list({bytes(rng.choice(alpha) for _ in range(rng.randint(1, rng.choice([4, 9, 20])))) for _ in range(rng.randint(5, 120))})
    lut = np.frombuffer(alpha, dtype=np.uint8)
    lens = [rng.choice([0, 1, 15, 16, 17, 500, 3000, 4096, 9000]) for _ in range(24)]
    hays = [lut[nrng.integers(0, len(alpha), size=n)] for n in lens]
    flat = np.concatenate(hays)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    exp = oracle_hits([pats], hays)
    a = build([pats])
    a.set_filter(-1)
    for tma in (-1, 1):
        a.set_tma(tma)
        for chunk, smem in ((0, 0), (64, 200), (256, 1024), (512, 64), (128, 8192)):
            a.set_tuning(chunk, smem)
            ev = a.search_events(flat, off)
            assert a.stats().filtered == 0
            assert_same(a, ev, len(lens), exp)
    a.release()


@pytest.mark.parametrize("smem", [0, 3000, 96 * 1024])
def test_four_byte_entries_with_and_without_a_window(smem):
    # more than 65,536 states: 4-byte table entries; a forced shared-memory window (window-relative ids, sink row) of a few
    # rows, of 96 KB, and the automatic choice; planted needles walk out of every window
    needles, hay, off = W.cfg2(n_hay=24, hay_len=5000, n_needles=7000, needle_len=14, planted_per_hay=5, seed=23)
    a = build([needles])
    assert a.info().entry_bytes == 4
    a.set_filter(-1)
    a.set_tuning(256, smem)
    ev = a.search_events(hay, off)
    assert a.stats().filtered == 0
    assert_same(a, ev, 24, oracle_hits([needles], split(hay, off)))


@pytest.mark.parametrize("planted", [0, 1, 8])
def test_full_walk_on_uniform_batches(planted):
    # 2,048-needle dictionary (deep states fall out of the shared-memory window), 64 x 8 KiB + one long haystack
    needles, hay, off = W.cfg2(n_hay=64, hay_len=8192, planted_per_hay=planted, seed=77)
    a = build([needles])
    ev = a.search_events(hay, off)
    assert_same(a, ev, 64, oracle_hits([needles], split(hay, off)))
    one = np.array([0, hay.size], dtype=np.uint64)          # the same bytes as ONE haystack: matches may straddle
    ev = a.search_events(hay, one)
    assert_same(a, ev, 1, oracle_hits([needles], [hay]))


def test_ragged_batch_with_empty_haystacks():
    rng = np.random.default_rng(5)
    needles, _, _ = W.cfg2(n_hay=1, hay_len=64, n_needles=64, needle_len=5, planted_per_hay=0, seed=3)
    lens = [0, 1, 4, 5, 0, 0, 17, 1000, 3, 0, 64, 65, 4097, 0]
    hays = [np.frombuffer(b"abcdef", dtype=np.uint8)[rng.integers(0, 6, size=n)] for n in lens]
    flat = np.concatenate(hays) if hays else np.zeros(0, np.uint8)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    a = build([needles])
    for chunk in (0, 16, 64, 32):
        a.set_tuning(chunk, 0)
        ev = a.search_events(flat, off)
        assert_same(a, ev, len(lens), oracle_hits([needles], hays))
        ev1 = a.search_events(flat, off, first_only=True)
        exp1 = oracle_hits([needles], hays, first_only=True)
        assert_same(a, ev1, len(lens), exp1)


def test_cfg3_signature_shape_reduced():
    pats, hay, off = W.cfg3(n_patterns=20_000, hay_bytes=8 << 20, plant_every=1 << 16)
    a = build([pats])
    inf = a.info()
    assert inf.entry_bytes == 4 and inf.n_classes == 256
    ev = a.search_events(hay, off)
    exp = oracle_hits([pats], [hay])
    assert exp[0][2] >= 100
    assert_same(a, ev, 1, exp)


def test_cfg5_adversarial_reduced_and_closed_form():
    n = 1 << 20
    pats, hay, off = W.cfg5(n_patterns=1100, hay_bytes=n)
    a = build([pats])
    assert a.info().n_patterns == 1024          # a^1025.. rejected (AC_PATTRN_MAX_LENGTH)
    ev = a.search_events(hay, off)
    events, hits = W.cfg5_expected(n, 1100)
    assert len(ev) == events
    assert np.array_equal(ev["end"], np.arange(1, n + 1, dtype=np.uint64))
    sizes = np.array([len(a.state_patterns(int(s))) for s in np.unique(ev["state"])])
    uniq, counts = np.unique(ev["state"], return_counts=True)
    assert int((sizes * counts).sum()) == hits
    # event-level hash against the oracle through the reference-style callback path
    res = {}
    for kind in ("oracle", "gpu"):
        d = Driver(kind)
        d.add_php_order(pats)
        d.finalize()
        r = d.search(hay, cap=16)
        res[kind] = (r["rc"], r["n_events"], r["n_hits"], r["hash"])
        d.release()
    assert res["gpu"] == res["oracle"]
    assert res["gpu"][2] == hits


def test_keep_streams_state_across_calls():
    needles, hay, _ = W.cfg2(n_hay=1, hay_len=5000, n_needles=200, planted_per_hay=8, seed=21)
    cuts = [0, 7, 8, 1000, 1003, 4096, 5000]
    res = {}
    for kind in ("oracle", "gpu"):
        d = Driver(kind)
        d.add_php_order(needles)
        d.finalize()
        pos, pat = [], []
        for i in range(len(cuts) - 1):
            r = d.search(hay[cuts[i]:cuts[i + 1]], keep=(i > 0))
            pos += r["pos"].tolist()
            pat += r["pat"].tolist()
        whole = d.search(hay)
        assert pos == whole["pos"].tolist() and pat == whole["pat"].tolist()
        res[kind] = (pos, pat)
        d.release()
    assert res["gpu"] == res["oracle"] and len(res["gpu"][0]) >= 8


def test_not_finalized_returns_minus_one_and_add_statuses():
    for kind in ("oracle", "gpu"):
        d = Driver(kind)
        assert d.add(b"abc") == 0
        assert d.add(b"abc") == 1            # duplicate
        assert d.add(b"") == 3               # empty
        assert d.add(b"x" * 1025) == 2       # too long
        assert d.add(b"x" * 1024) == 0
        assert d.search(b"abc")["rc"] == -1  # not finalized
        d.finalize()
        assert d.add(b"zzz") == 4            # closed
        r = d.search(b"zabcz")
        assert r["rc"] == 0 and r["pos"].tolist() == [4]
        d.release()


def test_one_cta_path_for_short_texts_equals_oracle():
    """ac_small_kernel (one haystack of up to 32 KiB: one copy, one launch, one wait): every size class around its
    slice boundaries, dense events (more than two per slice: the in-place second walk), binary bytes, keep=1
    continuations whose state comes from a longer chunk, and the first size that takes the general path again."""
    rng = np.random.default_rng(99)
    pats = [b"a", b"aa", b"aaaa", b"ab", b"abcab", b"\x00\xff", b"bca" * 5, b"cccccccccccccccccccc"] + \
           [bytes(rng.integers(97, 100, size=int(rng.integers(2, 12))).astype(np.uint8)) for _ in range(60)]
    a = build([pats])
    for n in (1, 2, 15, 16, 17, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 8192, 32767, 32768, 32769):
        hay = rng.integers(97, 100, size=n, dtype=np.uint8)
        if n > 64:
            hay[n // 3:n // 3 + 40] = ord("a")              # a dense burst
            hay[n - 2:] = np.frombuffer(b"\x00\xff", dtype=np.uint8)
        exp = oracle_hits([pats], [hay])
        ev = a.search_events(hay)
        assert (a.stats().kernel_launches == 1 and a.stats().kernel_ms == 0) == (n <= 32768) or n > 32768
        assert_same(a, ev, 1, exp)
        ev1 = a.search_events(hay, first_only=True)
        assert_same(a, ev1, 1, oracle_hits([pats], [hay], first_only=True))
    # keep=1: chunks of mixed sizes, states carried between the one-CTA path and the general one
    hay = rng.integers(97, 100, size=200_000, dtype=np.uint8)
    rc, whole = a.search_callback(hay.tobytes())
    seq, at = [], 0
    for size in (7, 1000, 40_000, 33, 32768, 100_000, 8192, 200_000):
        piece = hay[at:at + size]
        if piece.size == 0:
            break
        rc, g = a.search_callback(piece.tobytes(), keep=(at > 0))
        assert rc == 0
        seq += g
        at += piece.size
    assert at == hay.size and seq == whole
    d = Driver("oracle"); d.add_php_order(pats); d.finalize()
    r = d.search(hay)
    assert [p for p, _ in whole] == sorted(set(int(x) for x in r["pos"]))
    a.release()


def test_text_staged_by_the_tma_unit_equals_plain_loads_and_oracle():
    """ac_scan_tma_kernel (haystack text fetched as 32-byte x 32-slice boxes by the TMA unit into a per-warp shared-memory
    ring) against ac_scan_kernel (per-lane 16-byte loads) and the oracle: equal-length batches whose slices the box can
    serve, ragged batches where most tiles fall back, one long haystack (warm-up box from the row above), dense events."""
    import ctypes as C
    rng = np.random.default_rng(41)
    needles, hay, off = W.cfg2(n_hay=2048, hay_len=8192, planted_per_hay=8, seed=17)
    short = [bytes(rng.integers(97, 103, size=int(rng.integers(2, 33))).astype(np.uint8)) for _ in range(500)] + [b"ab", b"a" * 33]
    cases = [(needles, hay, off), (needles, hay, np.array([0, hay.size], dtype=np.uint64)), (short, hay, off)]
    lens = [5, 40_000, 0, 8192, 1_000_003, 64, 3_000_000]
    rag = np.concatenate([rng.integers(97, 103, size=n, dtype=np.uint8) for n in lens])
    rag[50_000:50_400] = ord("a")                               # a dense burst
    roff = np.zeros(len(lens) + 1, dtype=np.uint64); roff[1:] = np.cumsum(lens)
    cases.append((short, rag, roff))
    for pats, flat, offs in cases:
        a = build([pats])
        a.set_filter(-1)
        got = {}
        for mode in (-1, 1):
            a.L.acb200_set_tma(C.c_void_p(a.h), C.c_int(mode))
            for chunk in (0, 64, 256):
                a.set_tuning(chunk, 0)
                got[(mode, chunk)] = a.search_events(flat, offs)
                assert a.stats().filtered == 0
        ref = got[(-1, 0)]
        for k, ev in got.items():
            assert np.array_equal(ev, ref), k
        n = len(offs) - 1
        if flat.size <= 20_000_000 and n <= 16:
            assert_same(a, ref, n, oracle_hits([pats], split(flat, offs)))
        a.release()
