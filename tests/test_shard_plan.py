"""Host logic of the slab plan (csrc/shard.hpp, through acb200_plan_slabs — no GPU): the cuts cover the stream,
go round the devices in stream order, stay balanced, respect the halo rule, and — simulated with the CPU oracle standing in for the device — scanning
every slab from the root over [halo | own bytes] and dropping the events that end inside the halo reproduces the
uninterrupted scan of every haystack (the chunk-streaming rule of src/multifast/ahocorasick.c:191-194, 236-238)."""
import random

import numpy as np
import pytest

from oracle.pydriver import Driver
from php_aho_corasick_b200.native import plan_slabs


def _offsets(lens):
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    return off


def _check_cover(plans, off, halo_max, n_dev, slab):
    total = int(off[-1])
    assert plans[0]["begin"] == 0 and plans[-1]["end"] == total
    for i, (a, b) in enumerate(zip(plans, plans[1:])):
        assert a["end"] == b["begin"]
    # slab i goes to device i mod n_dev: all devices work on the same region of the stream at the same time
    assert [p["device_slot"] for p in plans] == [i % n_dev for i in range(len(plans))]
    for p in plans:
        assert 0 < p["end"] - p["begin"] <= slab + slab // 8
        h = p["first_text"]
        assert int(off[h]) <= p["begin"] < int(off[h + 1])
        assert p["halo"] == min(halo_max, p["begin"] - int(off[h]))
        assert int(off[p["end_text"] - 1]) < p["end"] <= total
        assert p["end_text"] == len(off) - 1 or int(off[p["end_text"]]) >= p["end"]
    assert {p["device_slot"] for p in plans} <= set(range(n_dev))


def test_plan_covers_balances_and_snaps_to_haystack_boundaries():
    # config 4 shape: 65,536 x 64 KiB over 8 GPUs -> every cut on a haystack boundary, no halo, equal shares
    off = np.arange(65537, dtype=np.uint64) * np.uint64(65536)
    plans = plan_slabs(off, 1023, 8, 64 << 20)
    _check_cover(plans, off, 1023, 8, 64 << 20)
    assert all(p["halo"] == 0 and p["begin"] % 65536 == 0 for p in plans)
    per_dev = [sum(p["end"] - p["begin"] for p in plans if p["device_slot"] == d) for d in range(8)]
    assert per_dev == [512 << 20] * 8
    # config 3 / 5 shape: ONE large haystack over 4 GPUs -> cuts inside it, each later slab carries Lmax-1 bytes
    off = np.array([0, (1 << 30) + 12345], dtype=np.uint64)
    plans = plan_slabs(off, 63, 4, 64 << 20)
    _check_cover(plans, off, 63, 4, 64 << 20)
    assert plans[0]["halo"] == 0 and all(p["halo"] == 63 for p in plans[1:])
    per_dev = [sum(p["end"] - p["begin"] for p in plans if p["device_slot"] == d) for d in range(4)]
    assert max(per_dev) - min(per_dev) <= len(plans) // 4 and len(plans) % 4 == 0
    # a stream smaller than the device count, empty haystacks at both ends, nothing at all
    off = _offsets([0, 0, 3, 0])
    plans = plan_slabs(off, 7, 8, 4096)
    _check_cover(plans, off, 7, 8, 4096)
    assert plan_slabs(_offsets([0, 0]), 7, 2, 4096) == []


@pytest.mark.parametrize("seed", range(6))
def test_slabs_with_halo_reproduce_the_uninterrupted_scan(seed):
    pyr = random.Random(seed)
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"abc", dtype=np.uint8)
    pats = list({bytes(alphabet[rng.integers(0, 3, size=pyr.randint(1, 9))]) for _ in range(40)})
    lmax = max(len(p) for p in pats)
    lens = [pyr.choice([0, 1, 5, 40, 333, 1000, 4097]) for _ in range(pyr.randint(1, 12))]
    hays = [alphabet[rng.integers(0, 3, size=n)] for n in lens]
    off = _offsets(lens)
    flat = np.concatenate(hays) if sum(lens) else np.zeros(0, np.uint8)
    d = Driver("oracle")
    d.add_php_order(pats)
    d.finalize()
    want = []
    for i, h in enumerate(hays):
        r = d.search(h)
        want += [(i, int(p), int(q)) for p, q in zip(r["pos"], r["pat"])]
    for n_dev, slab in ((1, 256), (3, 100), (8, 4096), (2, 1 << 20)):
        plans = plan_slabs(off, lmax - 1, n_dev, slab)
        if flat.size == 0:
            assert plans == []
            continue
        _check_cover(plans, off, lmax - 1, n_dev, slab)
        got = []
        for p in plans:
            # what the device is handed: the slab's pieces, piece 0 with its halo in front, every piece from the root
            for h in range(p["first_text"], p["end_text"]):
                b = max(int(off[h]), p["begin"])
                e = min(int(off[h + 1]), p["end"])
                halo = p["halo"] if h == p["first_text"] else 0
                if e <= b:
                    continue
                r = d.search(flat[b - halo:e])
                for pos, pat in zip(r["pos"], r["pat"]):
                    if pos > halo:          # ends inside the halo belong to the slab before
                        got.append((h, int(pos) - halo + b - int(off[h]), int(pat)))
        assert got == want, (seed, n_dev, slab)
    d.release()
