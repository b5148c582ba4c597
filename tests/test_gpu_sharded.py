"""GPU parity of the host pipeline (csrc/shard.hpp + cabi.cpp): one call cut into slabs — cuts inside haystacks
carry an (Lmax-1)-byte halo — and spread over several pipelines / GPUs gives the oracle's events in the oracle's
order, through ac_trie_search, ac_trie_search_flat, ac_trie_search_batch and acb200_search_events.
A one-GPU box runs the multi-pipeline cases with its device listed several times; with >= 2 GPUs the replicas
live on different devices."""
import random

import numpy as np
import pytest

from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton, lib
from tests.helpers import assert_same, oracle_hits, split

pytestmark = pytest.mark.gpu


def build(calls):
    a = Automaton(device=0)
    for c in calls:
        a.add_php_order(c)
    a.finalize()
    return a


def devices_for(k):
    n = lib().acb200_device_count()
    return [i % n for i in range(k)]


def rand_bytes(rng, n, alphabet):
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    return lut[rng.integers(0, len(alphabet), size=n)]


@pytest.mark.parametrize("pipelines,slab", [(1, 4096), (3, 4096), (2, 65536), (4, 1 << 20)])
def test_ragged_batch_in_slabs_equals_oracle(pipelines, slab):
    rng = np.random.default_rng(pipelines * 1000 + slab)
    pyr = random.Random(slab)
    pats = [rand_bytes(rng, pyr.randint(2, 40), b"abcd").tobytes() for _ in range(300)]
    lens = [0, 7, 100_000, 0, 5000, 1, 333_333, 4096, 8192, 0, 70_001, 2_000_000, 12, 0]
    hays = [rand_bytes(rng, n, b"abcd") for n in lens]
    for h in hays:                                   # plant patterns, also across where cuts will fall
        for _ in range(h.size // 2000):
            p = np.frombuffer(pyr.choice(pats), dtype=np.uint8)
            if h.size >= p.size:
                at = pyr.randint(0, h.size - p.size)
                h[at:at + p.size] = p
    off = np.zeros(len(lens) + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
    flat = np.concatenate(hays)
    a = build([pats])
    exp = oracle_hits([pats], hays)
    ev_direct = a.search_events(flat, off)
    assert a.stats().devices == 1
    assert_same(a, ev_direct, len(lens), exp)
    a.set_slab_bytes(slab)
    a.set_devices(devices_for(pipelines))
    ev = a.search_events(flat, off)
    st = a.stats()
    assert st.devices == pipelines and st.bytes >= flat.size       # halo bytes are scanned twice
    assert np.array_equal(ev, ev_direct)
    # findAll=false: the first event of every haystack, also where the haystack continues in a later slab
    ev1 = a.search_events(flat, off, first_only=True)
    assert_same(a, ev1, len(lens), oracle_hits([pats], hays, first_only=True))
    # the scattered-strings entry (ahocorasick_match_batch) and the flat one agree event for event
    t_batch = a.search_batch_tally(hays)
    t_flat = a.search_flat_tally(flat.ctypes.data, off)
    assert (t_batch.events, t_batch.hits, t_batch.hash) == (t_flat.events, t_flat.hits, t_flat.hash)
    assert t_batch.events == len(ev)
    a.release()


def test_one_long_haystack_spread_over_pipelines_and_keep_streaming():
    """config 3 / 5 at several GPUs: ONE haystack, cuts with halo; and ac_trie_search(keep=1) over chunks that are
    themselves cut into slabs continues exactly like the sequential walk."""
    rng = np.random.default_rng(9)
    pyr = random.Random(9)
    pats = [rand_bytes(rng, pyr.randint(8, 64), bytes(range(256))).tobytes() for _ in range(2000)] + [b"a" * k for k in (1, 2, 3, 500)]
    hay = rand_bytes(rng, 3_000_000, bytes(range(256)))
    for i in range(3000):
        p = np.frombuffer(pats[i % len(pats)], dtype=np.uint8)
        at = pyr.randint(0, hay.size - p.size)
        hay[at:at + p.size] = p
    hay[1_000_000:1_004_000] = ord("a")              # a dense burst that straddles cuts
    a = build([pats])
    a.set_slab_bytes(100_000)
    a.set_devices(devices_for(3))
    exp = oracle_hits([pats], [hay])
    ev = a.search_events(hay)
    assert a.stats().devices == 3
    assert_same(a, ev, 1, exp)
    # callback path, whole text
    rc, got = a.search_callback(hay.tobytes())
    assert rc == 0 and [p for p, _ in got] == sorted({int(x) for x in exp[0][0]})
    # keep=1 over three chunks (each cut into slabs again); positions carry the base offset
    cuts = [0, 1_000_123, 1_000_124, 2_345_678, hay.size]
    seq = []
    for i in range(len(cuts) - 1):
        rc, g = a.search_callback(hay[cuts[i]:cuts[i + 1]].tobytes(), keep=(i > 0))
        assert rc == 0
        seq += g
    assert seq == got
    # a callback that stops the search: rc 1, exactly one event delivered
    rc, g1 = a.search_callback(hay.tobytes(), stop_after_first=True)
    assert rc == 1 and g1 == got[:1]
    a.release()


@pytest.mark.parametrize("stage_min,threads", [("1", "5"), ("1", "1"), (str(1 << 40), "3")])
def test_direct_path_staged_by_helper_threads_equals_oracle(monkeypatch, stage_min, threads):
    """A pageable text / scattered strings below half a slab: the handle's parked helper threads copy them into pinned
    staging piece by piece and queue each piece's DMA (ACB200_STAGE_MIN=1 forces that route for every size, a huge value
    the plain cudaMemcpyAsync on the pageable pointer).  Same events either way, as the reference reports them — also
    for a keep=1 continuation, whose carried state enters the staged slab."""
    monkeypatch.setenv("ACB200_STAGE_MIN", stage_min)
    monkeypatch.setenv("ACB200_GATHER_THREADS", threads)
    rng = np.random.default_rng(21)
    pyr = random.Random(21)
    pats = [rand_bytes(rng, pyr.randint(3, 24), b"abc").tobytes() for _ in range(200)] + [b"abcabcabcabcabcabcabcabc"]
    lens = [0, 5, 300_000, 0, 65_536, 65_537, 1, 900_001, 0, 17]
    hays = [rand_bytes(rng, n, b"abc") for n in lens]
    off = np.zeros(len(lens) + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
    flat = np.concatenate(hays)
    a = build([pats])
    exp = oracle_hits([pats], hays)
    ev = a.search_events(flat, off)                  # flat pageable buffer
    assert a.stats().devices == 1
    assert_same(a, ev, len(lens), exp)
    t_batch = a.search_batch_tally(hays)             # separately allocated strings
    t_flat = a.search_flat_tally(flat.ctypes.data, off)
    assert (t_batch.events, t_batch.hits, t_batch.hash) == (t_flat.events, t_flat.hits, t_flat.hash)
    assert t_batch.events == len(ev)
    ev1 = a.search_events(flat, off, first_only=True)
    assert_same(a, ev1, len(lens), oracle_hits([pats], hays, first_only=True))
    # ac_trie_search, whole and as a keep=1 stream of ragged chunks (tiny ones take the one-CTA path unless staged)
    text = hays[7]
    rc, got = a.search_callback(text.tobytes())
    e7 = exp[7]
    assert rc == 0 and [p for p, _ in got] == sorted({int(x) for x in e7[0]})
    cuts = [0, 3, 200_000, 200_001, 700_123, text.size]
    seq = []
    for i in range(len(cuts) - 1):
        rc, g = a.search_callback(text[cuts[i]:cuts[i + 1]].tobytes(), keep=(i > 0))
        assert rc == 0
        seq += g
    assert seq == got
    a.release()


def test_cfg2_block_on_every_visible_gpu():
    n = lib().acb200_device_count()
    needles, hay, off = W.cfg2(n_hay=2048, hay_len=8192, planted_per_hay=8, seed=5)
    a = build([needles])
    ev1 = a.search_events(hay, off)
    a.set_slab_bytes(1 << 20)
    a.set_devices(list(range(n)) if n > 1 else [0, 0])
    ev = a.search_events(hay, off)
    assert a.stats().devices == max(n, 2)
    assert np.array_equal(ev, ev1)
    assert_same(a, ev[ev["text_idx"] < 64], 64, oracle_hits([needles], split(hay, off)[:64]))
    a.release()


def test_one_text_beyond_4_gib_through_the_drop_in_call():
    """The reference scans a text of any size_t length (src/multifast/ahocorasick.c:175-241); a device launch addresses
    its stream with 32 bits.  ac_trie_search cuts such a text into slabs itself: positions beyond 2^32 come back exact,
    a pattern planted across the 4 GiB mark and across a slab cut is found once."""
    total = (4 << 30) + (96 << 20) + 12345
    try:
        hay = np.zeros(total, dtype=np.uint8)
    except MemoryError:
        pytest.skip("not enough host memory for a 4.1 GiB text")
    pats = [b"needle-in-a-very-large-haystack", b"\x01\x02\x03\x04\x05\x06\x07\x08\x09", b"tail!"]
    from php_aho_corasick_b200.native import plan_slabs
    cut = plan_slabs(np.array([0, total], dtype=np.uint64), len(pats[0]) - 1, 1)[0]["end"]
    plant = {                                   # end offset -> pattern
        1000 + len(pats[0]): 0,
        cut + 7: 0,                             # straddles the first slab cut (ends 7 bytes behind it)
        (1 << 32) + 11: 0,                      # straddles the 4 GiB mark
        (1 << 32) + 5000 + len(pats[1]): 1,
        total: 2,                               # the last bytes of the text
    }
    for end, k in plant.items():
        p = np.frombuffer(pats[k], dtype=np.uint8)
        hay[end - p.size:end] = p
    a = build([pats])
    ev = a.search_events(hay)
    got = {int(e["end"]): a.state_patterns(int(e["state"]))[0][0] for e in ev}
    assert got == plant and len(ev) == len(plant)
    assert a.stats().bytes >= total
    a.release()
