"""CPU tests of the ORACLE (oracle/ac_oracle.c): pinned to the reference's golden vectors and, where the
reference tree is mounted, to the reference's own C sources compiled in place (oracle/_ref)."""
import random

import numpy as np
import pytest

from oracle import pydriver
from oracle.pydriver import Driver
from tests.helpers import case_calls, golden_record, load_golden, record_for

G = load_golden()
KINDS = ["oracle"] + (["reference"] if pydriver.available("reference") else [])


def run_case(kind, case):
    calls = case_calls(case)
    specs = [s for call in calls for s in call]
    d = Driver(kind)
    for call in calls:
        d.add_php_order([s["value"].encode("utf-8") for s in call])
    d.finalize()
    outs = []
    for m in case["matches"]:
        r = d.search(m["haystack"].encode("utf-8"), first_only=not m.get("findAll", True))
        outs.append([record_for(specs[int(o)], int(p)) for p, o in zip(r["pos"], r["pat"])])
    d.release()
    return outs


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test6"])
def test_golden_phpt_vectors(kind, name):
    for case in G[name]["cases"]:
        outs = run_case(kind, case)
        for got, m in zip(outs, case["matches"]):
            exp = [golden_record(r) for r in m["expect"]]
            assert len(got) == len(exp), (name, m["haystack"])
            for g, (order, e) in zip(got, exp):
                assert list(g.keys()) == order      # pos, key|keyIdx, aux, start_postion, value
                assert g == e


@pytest.mark.parametrize("kind", KINDS)
def test_golden_test4_and_test5(kind):
    t4 = G["test4"]
    d = Driver(kind)
    d.add_php_order([s["value"].encode() for s in t4["init"]])
    d.finalize()
    for _ in range(50):
        assert d.search(t4["haystack"].encode())["n_hits"] == t4["hits"]
    d.release()
    t5 = G["test5"]
    d = Driver(kind)
    d.add_php_order([s["value"].encode("utf-8") for s in t5["init"]])
    d.finalize()
    for h in t5["haystacks"]:
        assert d.search(h.encode("utf-8"))["n_hits"] == 0
    d.release()


@pytest.mark.skipif(not pydriver.available("reference"), reason="reference sources not mounted / oracle/_ref not built")
@pytest.mark.parametrize("seed", range(8))
def test_restatement_equals_reference_on_random_inputs(seed):
    rng = random.Random(seed)
    for trial in range(60):
        alpha = rng.choice([b"ab", b"abc", bytes(range(256)), b"a", b"\x00\xff\x80a", b"abcdef"])
        pats = [bytes(rng.choice(alpha) for _ in range(rng.randint(0, 10))) for _ in range(rng.randint(0, 60))]
        if trial % 10 == 0:
            pats += [b"a" * 1024, b"a" * 1025, b""]
        text = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 600)))
        outs = []
        for kind in ("oracle", "reference"):
            d = Driver(kind)
            cut = rng.randint(0, len(pats)) if kind == "oracle" else cut
            d.add_php_order(pats[:cut])
            d.add_php_order(pats[cut:])
            d.finalize()
            r = d.search(text, first_only=(trial % 3 == 0))
            outs.append((r["rc"], r["n_events"], r["hash"], r["pos"].tolist(), r["pat"].tolist(), r["len"].tolist()))
            d.release()
        assert outs[0] == outs[1], (seed, trial)


@pytest.mark.skipif(not pydriver.available("reference"), reason="reference sources not mounted / oracle/_ref not built")
def test_restatement_equals_reference_on_config_shapes():
    from php_aho_corasick_b200 import workloads as W
    shapes = [W.cfg2(n_hay=4, hay_len=8192), W.cfg3(n_patterns=3000, hay_bytes=1 << 18, plant_every=1 << 13),
              W.cfg5(n_patterns=70, hay_bytes=4096)]
    for pats, hay, off in shapes:
        outs = []
        for kind in ("oracle", "reference"):
            d = Driver(kind)
            d.add_php_order(pats)
            d.finalize()
            r = d.search(hay[:int(off[1])])
            outs.append((r["n_events"], r["hash"], r["pos"].tolist(), r["pat"].tolist()))
            d.release()
        assert outs[0] == outs[1]
        assert outs[0][0] > 0


@pytest.mark.parametrize("kind", KINDS)
def test_statuses_duplicates_and_keep(kind):
    d = Driver(kind)
    assert d.add(b"abc") == 0 and d.add(b"abc") == 1 and d.add(b"") == 3
    assert d.add(b"x" * 1025) == 2 and d.add(b"x" * 1024) == 0
    assert d.search(b"abc")["rc"] == -1
    d.finalize()
    assert d.add(b"q") == 4
    a = d.search(b"zzab")
    b = d.search(b"czz", keep=True)          # 'abc' straddles the two chunks
    assert a["n_hits"] == 0 and b["pos"].tolist() == [5] and b["pat"].tolist() == [0]
    assert d.search(b"czz")["n_hits"] == 0   # keep=0 starts over (tests/test6.phpt)
    d.release()
    # duplicates inside one PHP call: the LAST array element wins (reverse insertion, first add wins)
    d = Driver(kind)
    d.add_php_order([b"dup", b"x", b"dup"])
    d.finalize()
    assert d.search(b"dup")["pat"].tolist() == [2]
    d.release()
    # across calls the earlier call wins
    d = Driver(kind)
    d.add_php_order([b"dup"])
    d.add_php_order([b"dup"])
    d.finalize()
    assert d.search(b"dup")["pat"].tolist() == [0]
    d.release()
