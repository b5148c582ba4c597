"""GPU parity of the gram-prefilter path (ac_filter_kernel, ac_collect_kernel, ac_walk_kernel with the direct
verification of gram_table.hpp, ac_offsets_kernel, ac_emit_kernel; the asynchronous device call and the chained
multi-GPU step) against the CPU oracle and against
the full automaton walk (ac_scan_kernel): same events, same order, through the C-ABI."""
import random

import numpy as np
import pytest

from oracle.pydriver import Driver
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
from tests.helpers import assert_same, oracle_hits, split

pytestmark = pytest.mark.gpu


def build(pattern_calls, filter_mode):
    a = Automaton()
    for call in pattern_calls:
        a.add_php_order(call)
    a.finalize()
    a.set_filter(filter_mode)
    return a


def rand_bytes(rng, n, alphabet):
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    return lut[rng.integers(0, len(alphabet), size=n)]


def test_cfg2_planted_filtered_equals_oracle_and_full_walk():
    needles, hay, off = W.cfg2()
    exp = oracle_hits([needles], split(hay, off))
    a = build([needles], 1)
    inf = a.info()
    assert inf.filter_word == 8 and inf.min_pattern_len == 16
    ev = a.search_events(hay, off)
    st = a.stats()
    assert st.filtered == 1 and st.kernel_launches == 5 and 0 < st.flagged_words < hay.size // 8 // 10
    assert_same(a, ev, 256, exp)
    a.set_filter(-1)
    ev_full = a.search_events(hay, off)
    assert a.stats().filtered == 0
    assert np.array_equal(ev, ev_full)
    # the same bytes as ONE haystack: matches may straddle the former boundaries
    one = np.array([0, hay.size], dtype=np.uint64)
    a.set_filter(1)
    ev1 = a.search_events(hay, one)
    assert_same(a, ev1, 1, oracle_hits([needles], [hay]))


@pytest.mark.parametrize("seed", range(4))
def test_random_dictionaries_both_word_sizes_ragged_batches(seed):
    rng = np.random.default_rng(100 + seed)
    pyr = random.Random(seed)
    for trial in range(6):
        alphabet = pyr.choice([b"ab", b"abc", b"abcdef", bytes(range(256)), b"\x00\xff\x80a"])
        min_len = pyr.choice([8, 9, 15, 16, 17, 24])
        max_len = min_len + pyr.choice([0, 3, 20, 60])
        n_pat = pyr.choice([1, 3, 40, 400])
        pats = [rand_bytes(rng, pyr.randint(min_len, max_len), alphabet).tobytes() for _ in range(n_pat)]
        lens = [pyr.choice([0, 1, 7, 8, 15, 16, 17, 100, 511, 512, 513, 4095, 20000, 70001]) for _ in range(pyr.randint(1, 12))]
        hays = [rand_bytes(rng, n, alphabet) for n in lens]
        # plant patterns (and pattern prefixes / suffixes) so that there is something to find
        for h in hays:
            for _ in range(max(1, h.size // 700)):
                p = np.frombuffer(pyr.choice(pats), dtype=np.uint8)
                if h.size >= p.size:
                    at = pyr.randint(0, h.size - p.size)
                    h[at:at + p.size] = p
        if hays and hays[-1].size >= 64:
            p = np.frombuffer(pats[0], dtype=np.uint8)
            hays[-1][-p.size:] = p                      # a match that ends on the very last byte
            hays[-1][:p.size] = p                       # and one at offset 0
        flat = np.concatenate(hays) if hays else np.zeros(0, np.uint8)
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        exp = oracle_hits([pats], hays)
        a = build([pats], 1)
        assert a.info().filter_word == (8 if min(len(p) for p in pats) >= 16 else 4)
        ev = a.search_events(flat, off)
        assert a.stats().filtered == (1 if flat.size else 0)
        assert_same(a, ev, len(lens), exp)
        a.set_filter(-1)
        assert np.array_equal(ev, a.search_events(flat, off)), (seed, trial)
        a.release()


def test_calls_of_changing_size_on_one_handle_leave_nothing_behind():
    """No memset runs in front of a prefilter call: the filter kernel's CTA 0 clears the counters and block sums,
    ac_collect_kernel each tile's event count.  Batches of very different sizes and event densities back to back on ONE
    handle (large, tiny, large again with other content, empty of matches) must each equal the full walk."""
    needles, _, _ = W.cfg2(n_hay=1, hay_len=64)
    a = build([needles], 1)
    shapes = [(1024, 8192, 8, 1), (3, 700, 2, 2), (1024, 8192, 1, 3), (40, 8192, 0, 4), (512, 8192, 16, 5), (1, 17, 1, 6)]
    for n_hay, hay_len, planted, seed in shapes:
        _, hay, off = W.cfg2(n_hay=n_hay, hay_len=hay_len, planted_per_hay=planted, seed=seed)
        # the dictionary is the default-seed one: plant ITS needles
        rng = np.random.default_rng(seed)
        for h in range(n_hay):
            for _ in range(planted):
                if hay_len >= 16:
                    at = int(off[h]) + int(rng.integers(0, hay_len - 16 + 1))
                    hay[at:at + 16] = np.frombuffer(needles[int(rng.integers(0, len(needles)))], dtype=np.uint8)
        a.set_filter(1)
        ev = a.search_events(hay, off)
        assert a.stats().filtered == 1
        a.set_filter(-1)
        ref = a.search_events(hay, off)
        assert a.stats().filtered == 0
        assert np.array_equal(ev, ref), (n_hay, hay_len, planted)
        if planted and hay_len >= 16:
            assert len(ev) >= n_hay


def test_find_first_through_the_prefilter_equals_the_first_only_kernel():
    needles, hay, off = W.cfg2(n_hay=300, hay_len=8192, planted_per_hay=3, seed=31)
    hay = hay.copy()
    hay.reshape(300, 8192)[::5] = ord("z")              # every fifth haystack has no hit at all
    exp = oracle_hits([needles], split(hay, off), first_only=True)
    a = build([needles], 1)
    ev = a.search_events(hay, off, first_only=True)
    assert a.stats().filtered == 1
    assert_same(a, ev, 300, exp)
    a.set_filter(-1)
    ev2 = a.search_events(hay, off, first_only=True)
    assert a.stats().filtered == 0 and np.array_equal(ev, ev2)
    assert 200 <= len(ev) <= 240


def test_every_word_flagged_falls_back_to_whole_tile_walks():
    # nested patterns over one repeated byte: every aligned word is a pattern word, every offset an event
    pats = [b"a" * n for n in range(16, 41)]
    hay = np.full(300_000, ord("a"), dtype=np.uint8)
    hay[100_000:100_050] = ord("b")
    a = build([pats], 1)
    ev = a.search_events(hay)
    st = a.stats()
    assert st.filtered == 1 and st.dense_tiles > 0
    assert_same(a, ev, 1, oracle_hits([pats], [hay]))


@pytest.mark.parametrize("word", [8, 4])
def test_dense_and_sparse_tiles_side_by_side_own_every_end_offset_exactly_once(word):
    """A densely flagged 16 KiB tile is walked as whole spans, its neighbours word by word.  The end offsets right
    behind a tile boundary belong to the last word of the tile before it — they must be reported exactly once
    whichever way the two tiles are handled."""
    rng = np.random.default_rng(17)
    L = 2 * word
    filler = bytes(range(0x30, 0x61)) + bytes(range(0x62, 0x7b))       # no 'a': random text stays sparsely flagged
    pats = [b"a" * n for n in range(L, L + 6)] + [rand_bytes(rng, L + 3, filler).tobytes() for _ in range(20)]
    n_tiles = 12
    hay = rand_bytes(rng, n_tiles * 16384, filler)
    for t in (1, 2, 5, 8, 9, 10):                      # dense tiles, some adjacent, some isolated
        hay[t * 16384:(t + 1) * 16384] = ord("a")
    # needles that end 1..word bytes behind every tile boundary, and ones that straddle it further
    p0 = np.frombuffer(pats[-1], dtype=np.uint8)
    for t in range(1, n_tiles):
        for k, d in enumerate((1, word, word + 1, 3 * word)):
            if (t + k) % 2 == 0:
                end = t * 16384 + d
                hay[end - p0.size:end] = p0
    a = build([pats], 1)
    ev = a.search_events(hay)
    st = a.stats()
    assert st.filtered == 1 and 0 < st.dense_tiles < n_tiles
    assert a.info().filter_word == word
    assert_same(a, ev, 1, oracle_hits([pats], [hay]))
    a.set_filter(-1)
    assert np.array_equal(ev, a.search_events(hay))


def test_event_bursts_from_few_flagged_words_stay_ordered():
    # few flagged words per 16 KiB tile, but each yields a run of events (lanes with more than two re-walk and emit)
    rng = np.random.default_rng(9)
    pats = [b"a" * 16, b"a" * 17, rand_bytes(rng, 16, b"bcdef").tobytes()]
    hay = rand_bytes(rng, 1 << 18, b"bcdef")
    for at in range(5000, hay.size - 400, 16384):
        hay[at:at + 230] = ord("a")
    a = build([pats], 1)
    ev = a.search_events(hay)
    st = a.stats()
    assert st.filtered == 1 and st.dense_tiles == 0
    assert len(ev) > 16 * 200
    assert_same(a, ev, 1, oracle_hits([pats], [hay]))


def test_cfg3_signature_shape_reduced_filtered_with_second_level(monkeypatch):
    pats, hay, off = W.cfg3(n_patterns=20_000, hay_bytes=8 << 20, plant_every=1 << 16)
    monkeypatch.setenv("ACB200_L2_MIN_FILL", "0.01")       # force the global second-level bitmap at this reduced size
    a = build([pats], 1)
    inf = a.info()
    assert inf.filter_word == 4 and inf.filter_l2_log2 > 0 and inf.min_pattern_len == 8
    ev = a.search_events(hay, off)
    st = a.stats()
    assert st.filtered == 1
    exp = oracle_hits([pats], [hay])
    assert exp[0][2] >= 100
    assert_same(a, ev, 1, exp)
    # the second level keeps the flagged share far below the first level's fill
    assert st.flagged_words < (hay.size // 4) * 0.02


def test_keep_continuation_after_a_filtered_first_chunk():
    needles, hay, _ = W.cfg2(n_hay=1, hay_len=40_000, n_needles=200, planted_per_hay=8, seed=21)
    cuts = [0, 30_000, 30_007, 30_008, 39_000, 40_000]
    hay[29_990:30_006] = np.frombuffer(needles[5], dtype=np.uint8)      # straddles the first cut
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
        res[kind] = (pos, pat)
        d.release()
    assert res["gpu"] == res["oracle"] and len(res["gpu"][0]) >= 8
    # and with the prefilter forced on the first chunk through the native binding
    a = build([needles], 1)
    rc, got0 = a.search_callback(hay[:30_000].tobytes())
    assert a.stats().filtered == 1
    rc, got1 = a.search_callback(hay[30_000:].tobytes(), keep=True)
    assert a.stats().filtered == 0                     # a continuation does not start at the root
    pos = [p for p, _ in got0] + [p for p, _ in got1]
    assert pos == res["oracle"][0]


def test_device_resident_stream_with_odd_length_is_not_read_past_its_end():
    torch = pytest.importorskip("torch")
    needles, hay, _ = W.cfg2(n_hay=1, hay_len=(1 << 20) + 13, n_needles=500, planted_per_hay=64, seed=5)
    p = np.frombuffer(needles[3], dtype=np.uint8)
    hay[-16:] = p
    a = build([needles], 1)
    t = torch.from_numpy(hay).cuda()
    off = np.array([0, hay.size], dtype=np.uint64)
    ptr, n = a.search_device(t.data_ptr(), off)
    assert a.stats().filtered == 1
    buf = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    a.copy_events(buf.data_ptr(), n)
    torch.cuda.synchronize()
    got = buf.cpu().numpy().view(np.uint32)
    exp = oracle_hits([needles], [hay])[0]
    assert n == exp[2]
    assert int(got[-1, 0]) == hay.size
    ev = a.search_events(hay, off)
    assert np.array_equal(got[:, 0].astype(np.uint64), ev["end"]) and np.array_equal(got[:, 1], ev["state"])


def test_pipelined_flat_search_equals_monolithic_batch_search():
    """ac_trie_search_flat cuts batches above 128 MiB into 64 MiB slabs (upload of slab i+1 overlaps scan and
    replay of slab i); the callback sequence must be the one of the single-launch path."""
    import ctypes as C
    from php_aho_corasick_b200 import native
    needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=8)
    reps = 72                                           # 144 MiB, haystack lengths varied below
    flat = np.tile(hay, reps)
    n = reps * 256
    lens = np.full(n, 8192, dtype=np.uint64)
    lens[::7] -= 3                                      # ragged: slab cuts fall on unaligned offsets
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    flat = np.ascontiguousarray(flat[: int(offsets[-1])])
    a = build([needles], 0)
    t_pipe = a.search_flat_tally(flat.ctypes.data, offsets)
    assert a.stats().bytes == flat.size and a.stats().kernel_launches >= 3 * 5
    # monolithic reference: ac_trie_search_batch gathers the texts and scans them in one launch
    texts = (native.AcText * n)()
    base = flat.ctypes.data
    for i in range(n):
        texts[i].astring = base + int(offsets[i])
        texts[i].length = int(lens[i])
    t_mono = native.Tally()
    cb = C.cast(a.L.acb200_tally_cb, native.BATCH_CB)
    rc = a.L.ac_trie_search_batch(a.h, texts, n, 0, cb, C.cast(C.byref(t_mono), C.c_void_p))
    assert rc == 0
    assert (t_pipe.events, t_pipe.hits, t_pipe.hash) == (t_mono.events, t_mono.hits, t_mono.hash)
    assert t_pipe.events > n * 5
    # findAll=false through both paths
    t1 = a.search_flat_tally(flat.ctypes.data, offsets, first_only=True)
    t2 = native.Tally()
    rc = a.L.ac_trie_search_batch(a.h, texts, n, 1, cb, C.cast(C.byref(t2), C.c_void_p))
    assert rc == 0 and (t1.events, t1.hash) == (t2.events, t2.hash) and t1.events == n


def test_full_size_config2_device_resident_filter_equals_full_walk_and_planted_count():
    """BASELINE config 2 at the size bench.py runs (1 GiB, 131,072 x 8 KiB): the two device paths must produce the
    same event array; every haystack is a copy of one of 256 distinct ones, so the event total is 512 x the
    oracle's total for the 256-haystack block."""
    torch = pytest.importorskip("torch")
    needles, hay, off = W.cfg2()
    exp_block = sum(e[2] for e in oracle_hits([needles], split(hay, off)))
    reps = 512
    big = torch.from_numpy(hay).cuda().repeat(reps)
    a = build([needles], 1)
    _, n1 = a.search_device_uniform(big.data_ptr(), 256 * reps, 8192)
    assert a.stats().filtered == 1 and a.stats().bytes == 1 << 30
    assert n1 == exp_block * reps
    ev1 = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
    a.copy_events(ev1.data_ptr(), n1)
    a.set_filter(-1)
    _, n2 = a.search_device_uniform(big.data_ptr(), 256 * reps, 8192)
    assert a.stats().filtered == 0 and n2 == n1
    ev2 = torch.empty((n2, 2), dtype=torch.int32, device="cuda")
    a.copy_events(ev2.data_ptr(), n2)
    torch.cuda.synchronize()
    assert torch.equal(ev1, ev2)
    # ascending end offsets (viewed as unsigned), and periodic with the block: event i+k == event i + 2 MiB
    ends = ev1[:, 0].to(torch.int64) & 0xFFFFFFFF
    assert bool((ends[1:] > ends[:-1]).all())
    assert torch.equal(ends[exp_block:], ends[:-exp_block] + hay.size)
    assert torch.equal(ev1[exp_block:, 1], ev1[:-exp_block, 1])


def test_config4_shape_reduced_through_the_pipelined_batch_call():
    """BASELINE config 4 shape (64 KiB haystacks, config-2 dictionary) at 8,192 haystacks = 512 MiB through
    ac_trie_search_flat (eight pipelined slabs): totals against the oracle on a sample of haystacks."""
    needles, hay, off = W.cfg2(n_hay=64, hay_len=65536, planted_per_hay=16, seed=44)
    reps = 128
    flat = np.tile(hay, reps)
    offsets = W.offsets_uniform(64 * reps, 65536)
    a = build([needles], 0)
    t = a.search_flat_tally(flat.ctypes.data, offsets)
    st = a.stats()
    assert st.bytes == flat.size and st.filtered == 1
    exp = oracle_hits([needles], split(hay, off))
    assert t.events == reps * sum(e[2] for e in exp)
    assert t.hits == reps * sum(len(e[0]) for e in exp)


def test_full_size_config5_adversarial_closed_form_on_the_device():
    """BASELINE config 5 at full size: a^1..a^4096 submitted (a^1..a^1024 accepted) over 256 MiB of 'a' — one event
    per byte; event i ends at i+1 and reports min(i+1, 1024) patterns.  Checked on the device (2 GiB of events)."""
    torch = pytest.importorskip("torch")
    n = 256 << 20
    pats = [b"a" * (i + 1) for i in range(4096)]
    a = build([pats], 0)
    assert a.info().n_patterns == 1024 and a.info().filter_word == 0      # shortest pattern is 1 byte: full walk only
    text = torch.full((n,), ord("a"), dtype=torch.uint8, device="cuda")
    ptr, ne = a.search_device(text.data_ptr(), np.array([0, n], dtype=np.uint64))
    assert ne == n and a.stats().filtered == 0
    ev = torch.empty((ne, 2), dtype=torch.int32, device="cuda")
    a.copy_events(ev.data_ptr(), ne)
    torch.cuda.synchronize()
    ends = ev[:, 0].to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(ends, torch.arange(1, n + 1, device="cuda", dtype=torch.int64))
    del ends
    # state of event i reports min(i+1, 1024) patterns: 1024 distinct states, the deepest one from offset 1024 on
    states = ev[:, 1]
    head = states[:1024].cpu().numpy().astype(np.uint32)
    sizes = [len(a.state_patterns(int(s))) for s in head]
    assert sizes == list(range(1, 1025))
    assert bool((states[1024:] == states[1023]).all())
    events, hits = W.cfg5_expected(n)
    assert events == ne and hits == sum(sizes) + (n - 1024) * 1024


def test_full_size_config3_signatures_filter_equals_full_walk():
    """BASELINE config 3: 100,000 binary signatures of 8..64 bytes (3.45 M states, 3.5 GB table) over a 1 GiB
    binary haystack with one planted signature per MiB: both device paths give the same events, every planted
    signature is found."""
    torch = pytest.importorskip("torch")
    pats, hay, off = W.cfg3()
    a = build([pats], 1)
    inf = a.info()
    assert inf.n_states > 3_000_000 and inf.filter_word == 4 and inf.filter_l2_log2 > 0
    text = torch.from_numpy(hay).cuda()
    _, n1 = a.search_device(text.data_ptr(), off)
    assert a.stats().filtered == 1
    ev1 = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
    a.copy_events(ev1.data_ptr(), n1)
    a.set_filter(-1)
    _, n2 = a.search_device(text.data_ptr(), off)
    assert a.stats().filtered == 0
    ev2 = torch.empty((n2, 2), dtype=torch.int32, device="cuda")
    a.copy_events(ev2.data_ptr(), n2)
    torch.cuda.synchronize()
    assert n1 == n2 and torch.equal(ev1, ev2)
    assert n1 >= 1000                                   # 1,024 planted (a few may overwrite each other)
    # every event really is a signature ending there (host check of all of them)
    e = ev1.cpu().numpy().view(np.uint32)
    for end, state in e[:: max(1, len(e) // 200)]:
        lst = a.state_patterns(int(state))
        assert lst, state
        o, l = lst[0]
        assert hay[int(end) - l:int(end)].tobytes() == pats[o]


@pytest.mark.parametrize("seed", range(6))
def test_randomised_mixtures_of_dense_and_sparse_regions(seed):
    """Haystacks stitched from random filler, long runs of a repeated pattern byte / pattern prefix, and pattern
    copies at random offsets; ragged batches.  Prefilter path vs oracle vs full walk."""
    rng = np.random.default_rng(1000 + seed)
    pyr = random.Random(seed)
    for trial in range(5):
        word = pyr.choice([4, 8])
        L = 2 * word
        alphabet = pyr.choice([b"ab", b"abcd", bytes(range(256)), bytes(range(0x30, 0x7b))])
        pats = [rand_bytes(rng, pyr.randint(L, L + 40), alphabet).tobytes() for _ in range(pyr.choice([3, 30, 300]))]
        rep = bytes([alphabet[0]])
        pats += [rep * n for n in range(L, L + pyr.choice([1, 5, 30]))]          # nested runs of one byte
        hays = []
        for _ in range(pyr.randint(1, 6)):
            parts = []
            for _ in range(pyr.randint(1, 8)):
                kind = pyr.choice(["fill", "fill", "run", "prefix", "pats"])
                n = pyr.choice([0, 5, 100, 3000, 20000, 70000])
                if kind == "fill":
                    parts.append(rand_bytes(rng, n, alphabet))
                elif kind == "run":
                    parts.append(np.full(n, alphabet[0], dtype=np.uint8))
                elif kind == "prefix":
                    p = np.frombuffer(pyr.choice(pats), dtype=np.uint8)
                    parts.append(np.tile(p[: max(1, p.size - 1)], n // max(1, p.size - 1) + 1)[:n])
                else:
                    parts.append(np.concatenate([np.frombuffer(pyr.choice(pats), dtype=np.uint8) for _ in range(1 + n // 500)]))
            hays.append(np.concatenate(parts) if parts else np.zeros(0, np.uint8))
        lens = [h.size for h in hays]
        flat = np.concatenate(hays) if sum(lens) else np.zeros(0, np.uint8)
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        exp = oracle_hits([pats], hays)
        a = build([pats], 1)
        ev = a.search_events(flat, off)
        assert_same(a, ev, len(lens), exp)
        a.set_filter(-1)
        assert np.array_equal(ev, a.search_events(flat, off)), (seed, trial)
        a.release()


def test_direct_verification_equals_walking_every_flagged_word():
    """gram_table.hpp: flagged words settled by one comparison inside ac_walk_kernel (default) vs every flagged word
    walked (set_direct(-1)) vs the full automaton walk — raw events (end offset AND state id) must be identical."""
    rng = np.random.default_rng(77)
    pyr = random.Random(77)
    cases = []
    needles, hay, off = W.cfg2(n_hay=1024, hay_len=8192, planted_per_hay=8, seed=21)
    cases.append((needles, hay, off))
    # nested / overlapping patterns: shared grams and failure-target nodes must fall back to the walk
    nested = [b"0123456789abcdef", b"xx0123456789abcdef", b"456789abcdefghij", b"Q0123456789abcdefZZ",
              b"aaaaaaaaaaaaaaaa", b"aaaaaaaaaaaaaaaaaaaa", b"hello, world....", b"say hello, world....!"]
    lens = [5, 300, 0, 16, 17, 70001, 4096, 9_000_000]
    hays = [rand_bytes(rng, n, b"abcdefx0123456789") for n in lens]
    for h in hays:
        for _ in range(h.size // 500):
            p = np.frombuffer(pyr.choice(nested), dtype=np.uint8)
            if h.size >= p.size:
                at = pyr.randint(0, h.size - p.size)
                h[at:at + p.size] = p
    hays[-1][5000:5100] = ord("a")
    o = np.zeros(len(lens) + 1, dtype=np.uint64); o[1:] = np.cumsum(lens)
    cases.append((nested, np.concatenate(hays), o))
    # W = 4, binary patterns of 8..64 bytes (config-3 shape): comparisons of up to 16 chunks
    sigs = [rand_bytes(rng, pyr.randint(8, 64), bytes(range(256))).tobytes() for _ in range(3000)]
    big = rand_bytes(rng, 12_000_000, bytes(range(256)))
    for i in range(4000):
        p = np.frombuffer(sigs[i % len(sigs)], dtype=np.uint8)
        at = pyr.randint(0, big.size - p.size)
        big[at:at + p.size] = p
    big[:len(sigs[0])] = np.frombuffer(sigs[0], dtype=np.uint8)
    big[big.size - len(sigs[1]):] = np.frombuffer(sigs[1], dtype=np.uint8)
    lens3 = [3_000_000, 1, 4_999_999, 4_000_000]
    o3 = np.zeros(5, dtype=np.uint64); o3[1:] = np.cumsum(lens3)
    cases.append((sigs, big, o3))
    for pats, flat, offs in cases:
        a = build([pats], 1)
        inf = a.info()
        assert inf.direct_keys > 0
        ev = a.search_events(flat, offs)
        st = a.stats()
        assert st.filtered == 1 and st.kernel_launches == 5 and len(ev) > 100
        a.set_direct(-1)
        ev_walk = a.search_events(flat, offs)
        st = a.stats()
        assert st.filtered == 1 and st.kernel_launches == 5
        assert np.array_equal(ev, ev_walk)
        a.set_filter(-1)
        ev_full = a.search_events(flat, offs)
        assert a.stats().filtered == 0
        assert np.array_equal(ev, ev_full)
        a.release()
    # and against the oracle on the nested case
    pats, flat, offs = cases[1]
    a = build([pats], 1)
    assert_same(a, a.search_events(flat, offs), len(offs) - 1, oracle_hits([pats], split(flat, offs)))


def test_asynchronous_device_search_fills_the_callers_rows():
    """acb200_search_device_uniform_async: nothing is waited for inside the call, row 0 of the caller's device buffer
    receives the count, the rows after it the events of the synchronous call; too few rows -> the count says so and
    only the first rows are written; batches the prefilter cannot take are refused (the caller uses the synchronous
    call)."""
    import torch
    needles, hay, off = W.cfg2(n_hay=2048, hay_len=8192, planted_per_hay=8, seed=33)      # 16 MiB
    a = build([needles], 0)
    dev = torch.from_numpy(hay).to("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    ptr, n = a.search_device_uniform(dev.data_ptr(), 2048, 8192, stream=stream)
    assert a.stats().filtered == 1 and n > 10000
    ref = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
    assert a.copy_events(ref.data_ptr(), n, stream=stream) == n
    torch.cuda.synchronize()
    for rows in (n + 100, n, n // 2):
        buf = torch.full((rows + 1, 2), -1, dtype=torch.int32, device="cuda:0")
        assert a.search_device_uniform_async(dev.data_ptr(), 2048, 8192, buf.data_ptr(), rows, stream=stream)
        got = buf.cpu()                                         # ordered by the STREAM alone (torch's current stream): no device-wide wait
        assert int(got[0, 0]) == n                              # the full count, also when the rows were too few
        a.async_finish(n, int(got[0, 1]))
        st = a.stats()
        assert st.events == n and st.kernel_ms > 0 and st.filtered == 1
        m = min(n, rows)
        assert torch.equal(got[1:1 + m], ref[:m].cpu())
        assert bool((got[1 + m:] == -1).all())                  # nothing written behind the rows that fit
    # a small batch goes to the full walk in automatic mode: no asynchronous form
    small = dev[: 64 * 8192]
    buf = torch.zeros((1025, 2), dtype=torch.int32, device="cuda:0")
    assert not a.search_device_uniform_async(small.data_ptr(), 64, 8192, buf.data_ptr(), 1024, stream=stream)
    a.set_filter(1)                                              # forced prefilter: served
    assert a.search_device_uniform_async(small.data_ptr(), 64, 8192, buf.data_ptr(), 1024, stream=stream)
    torch.cuda.synchronize()
    _, n_small = a.search_device_uniform(small.data_ptr(), 64, 8192, stream=stream)
    assert int(buf[0, 0].cpu()) == n_small


def test_sharded_matcher_chained_step_on_one_rank():
    """dist.ShardedMatcher on a one-rank NCCL group: the first scan_and_gather step is synchronous (it agrees on the
    rows), the following ones chain scan -> all_gather on the stream; the gathered rows must be the events of the
    plain call, and match() must return what search_events returns."""
    import os
    import socket
    import torch
    import torch.distributed as dist
    from php_aho_corasick_b200.dist import ShardedMatcher
    needles, hay, off = W.cfg2(n_hay=2048, hay_len=8192, planted_per_hay=8, seed=44)      # 16 MiB: prefilter in automatic mode
    a = build([needles], 0)
    ref = a.search_events(hay, off)
    stream_end = ref["end"].astype(np.int64) + ref["text_idx"].astype(np.int64) * 8192
    created = False
    if not dist.is_initialized():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    try:
        dev = torch.from_numpy(hay).to("cuda:0")
        stream = torch.cuda.current_stream().cuda_stream
        sm = ShardedMatcher(a)
        for step in range(3):
            n, got = sm.scan_and_gather(dev, off, 0, stream=stream, uniform_len=8192)
            assert n == len(ref) and len(got) == 1 and got[0].shape[0] == n, step
            ev = got[0].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            assert np.array_equal(ev[:, 0], stream_end) and np.array_equal(ev[:, 1], ref["state"].astype(np.int64)), step
            st = a.stats()
            assert st.filtered == 1 and st.events == n and st.kernel_ms > 0
        assert np.array_equal(sm.match(hay, off), ref)
    finally:
        if created:
            dist.destroy_process_group()
