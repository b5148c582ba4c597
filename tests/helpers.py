"""Shared helpers for the parity tests (TEST code: may use oracle/)."""
from __future__ import annotations

import numpy as np

from oracle import pydriver
from oracle.pydriver import Driver


def checker_kind(patterns_calls):
    """The reference's own compiled code (oracle/_ref) wherever its finalize is cheap — it is quadratic in the depth
    per node and quartic on nested chains — else the C restatement, which tests/test_oracle.py pins to it."""
    if not pydriver.available("reference"):
        return "oracle"
    n = sum(len(c) for c in patterns_calls)
    longest = max((len(p) for c in patterns_calls for p in c), default=0)
    return "reference" if n <= 5000 and longest <= 128 else "oracle"


def oracle_hits(patterns_calls, haystacks, first_only=False, kind=None):
    """patterns_calls: list of pattern lists, one per init()/add_patterns() call.
    -> list per haystack of (pos[], ordinal[], n_events, hash)"""
    d = Driver(kind or checker_kind(patterns_calls))
    for call in patterns_calls:
        d.add_php_order(call)
    d.finalize()
    out = []
    for h in haystacks:
        r = d.search(h, first_only=first_only)
        out.append((r["pos"].copy(), r["pat"].copy(), r["n_events"], r["hash"]))
    d.release()
    return out


def split(flat: np.ndarray, offsets: np.ndarray):
    return [flat[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]


def gpu_hits_by_haystack(aut, events, n_hay):
    ti, pos, pat, _ln = aut.expand(events)
    out = []
    for h in range(n_hay):
        m = ti == h
        out.append((pos[m], pat[m], int((events["text_idx"] == h).sum())))
    return out


def assert_same(aut, events, n_hay, expected):
    got = gpu_hits_by_haystack(aut, events, n_hay)
    for h in range(n_hay):
        epos, epat, enev, _ = expected[h]
        gpos, gpat, gnev = got[h]
        assert gnev == enev, f"haystack {h}: {gnev} events, oracle {enev}"
        assert np.array_equal(gpos, epos), f"haystack {h}: positions differ"
        assert np.array_equal(gpat, epat), f"haystack {h}: pattern order differs"


# ---- golden fixtures (tests/golden/phpt_golden.json, generated from the reference's tests/*.phpt) ----
import json
import os

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phpt_golden.json")


def load_golden():
    with open(GOLDEN_PATH, encoding="utf-8") as f:
        return json.load(f)


def golden_record(rec):
    """golden var_dump array -> (ordered key list, plain dict)"""
    order = rec["__order__"]
    return order, {k: rec[k] for k in order}


def record_for(spec: dict, pos: int):
    """What php_ahocorasick_match_handler (src/php_ahocorasick.c:542-589) builds for pattern `spec` ending at pos."""
    d = {"pos": pos}
    if "key" in spec:
        d["key"] = spec["key"]
    elif "id" in spec:
        d["keyIdx"] = spec["id"]
    if "aux" in spec:
        d["aux"] = spec["aux"]
    d["start_postion"] = pos - len(spec["value"].encode("utf-8"))
    d["value"] = spec["value"]
    return d


def case_calls(case):
    """pattern-array calls of a golden case: init(...) then each add_patterns(...)"""
    return [case["init"]] + list(case.get("add_patterns", []))
