"""Save / load of a finalized automaton (acb200_save / acb200_load).  The CPU half checks the file
round trip (flat description, output lists, pattern ids); the GPU half that a loaded automaton matches like the original."""
import ctypes as C
import os

import numpy as np
import pytest

from php_aho_corasick_b200 import native, workloads as W
from php_aho_corasick_b200.native import Automaton


def _finalize_anyhow(a):
    a.L.ac_trie_finalize(a.h)          # no GPU here: the host part of finalize still runs


def test_blob_round_trip_on_the_host(tmp_path):
    pats = [b"alfa", b"beta", b"gamma", b"lfa", b"a", b"\x00\xff\x80", b"alfabet"]
    a = Automaton()
    a.add_php_order(pats)
    _finalize_anyhow(a)
    path = str(tmp_path / "dict.acb")
    a.save(path)
    b = Automaton.load(path, require_device=False)
    ia, ib = a.info(), b.info()
    assert (ia.n_patterns, ia.n_states, ib.finalized) == (ib.n_patterns, ib.n_states, 1) and ib.n_patterns == len(pats)
    for state in range(1, 12):
        assert a.state_patterns(state) == b.state_patterns(state)
    # ids and bytes survive (this binding adds numeric ids)
    for i in range(len(pats)):
        pa, pb = a.L.acb200_pattern(a.h, i).contents, b.L.acb200_pattern(b.h, i).contents
        assert C.string_at(pa.ptext.astring, pa.ptext.length) == C.string_at(pb.ptext.astring, pb.ptext.length)
        assert (pa.id.type, pa.id.u.number, pa.aux) == (pb.id.type, pb.id.u.number, pb.aux)
    assert b.add(b"zzz") == 4                       # a loaded automaton is closed (ACERR_TRIE_CLOSED)


def test_loaded_blob_rebuilds_the_gram_table(tmp_path):
    """The exact gram table (direct verification of flagged words) is derived from the flat description: a loaded
    automaton must have the same table — same verdict, end and state for every word of a text."""
    import random
    rng = random.Random(11)
    pats = [bytes(rng.choice(b"abcdef") for _ in range(rng.choice([16, 17, 24, 40]))) for _ in range(200)]
    pats += [b"x" + pats[0], pats[1][4:] + b"yyyy"]            # a failure-target pattern and an overlapping one
    a = Automaton()
    a.add_php_order(pats)
    _finalize_anyhow(a)
    path = str(tmp_path / "grams.acb")
    a.save(path)
    b = Automaton.load(path, require_device=False)
    ia, ib = a.info(), b.info()
    assert ia.direct_keys > 0 and (ia.direct_keys, ia.direct_walk_keys, ia.filter_word) == (ib.direct_keys, ib.direct_walk_keys, ib.filter_word)
    text = bytearray(rng.choice(b"abcdef") for _ in range(4000))
    for i in range(60):
        p = pats[rng.randrange(len(pats))]
        at = rng.randint(0, len(text) - len(p))
        text[at:at + len(p)] = p
    for at in (100, 1003, 2501):                               # the failure-target pattern: its words must be "undecided"
        text[at:at + len(pats[0])] = pats[0]
    text = bytes(text)
    verdicts = [a.direct_probe(text, k) for k in range(len(text) // 8)]
    assert verdicts == [b.direct_probe(text, k) for k in range(len(text) // 8)]
    assert sum(1 for v in verdicts if v[0] == 1) >= 20 and any(v[0] == 2 for v in verdicts)


def test_blob_rejects_garbage_and_unfinalized(tmp_path):
    a = Automaton()
    a.add(b"abc")
    with pytest.raises(native.AcError):
        a.save(str(tmp_path / "x.acb"))             # not finalized
    bad = tmp_path / "bad.acb"
    bad.write_bytes(b"ACB200v2" + os.urandom(100))
    with pytest.raises(native.AcError):
        Automaton.load(str(bad), require_device=False)
    with pytest.raises(native.AcError):
        Automaton.load(str(tmp_path / "missing.acb"), require_device=False)
    _finalize_anyhow(a)
    good = tmp_path / "good.acb"
    a.save(str(good))
    data = good.read_bytes()
    (tmp_path / "cut.acb").write_bytes(data[: len(data) // 2])
    with pytest.raises(native.AcError):
        Automaton.load(str(tmp_path / "cut.acb"), require_device=False)


def _blob_layout(data: bytes):
    """file offsets of the blob's fields (csrc/blob.cpp save_flat) -> dict name -> (offset, element size, count)"""
    import struct
    pos = 8
    out = {}
    for name in ("n_states", "n_rows", "n_classes", "final_bound", "root", "max_pattern_len", "n_used_bytes"):
        out[name] = (pos, 4, 1); pos += 4
    out["cls_map"] = (pos, 1, 256); pos += 256
    out["range_map"] = (pos, 4, 1); pos += 4
    out["range_lo"] = (pos, 4, 1); pos += 4
    for name, size in (("bfs_order", 4), ("level_off", 4), ("fail", 4), ("edge_src", 4), ("edge_dst", 4), ("edge_cls", 2),
                       ("level_edge_off", 4), ("out_off", 8), ("out_idx", 4)):
        (n,) = struct.unpack_from("<Q", data, pos)
        out[name] = (pos + 8, size, n); pos += 8 + size * n
    for name in ("min_pattern_len", "filter_w", "l1_bits"):
        out[name] = (pos, 4, 1); pos += 4
    (n,) = struct.unpack_from("<Q", data, pos)
    out["l1"] = (pos + 8, 4, n); pos += 8 + 4 * n
    out["l2_log2"] = (pos, 4, 1); pos += 4
    return out


def test_blob_with_corrupt_fields_fails_to_load(tmp_path):
    """Every index, offset and size the loader (or the kernels after it) would trust is validated: a file with one
    field changed must be refused, not loaded into memory corruption or silently wrong matches."""
    import random
    import struct
    rng = random.Random(5)
    pats = [bytes(rng.choice(b"abcdef") for _ in range(rng.choice([16, 20, 33]))) for _ in range(50)]
    a = Automaton()
    a.add_php_order(pats)
    _finalize_anyhow(a)
    good = tmp_path / "good.acb"
    a.save(str(good))
    data = good.read_bytes()
    lay = _blob_layout(data)
    assert Automaton.load(str(good), require_device=False).info().n_patterns == len(pats)

    def patched(name, index, value):
        off, size, n = lay[name]
        assert index < n
        b = bytearray(data)
        struct.pack_into({1: "<B", 2: "<H", 4: "<I", 8: "<Q"}[size], b, off + size * index, value)
        return bytes(b)

    n_out = lay["out_off"][2]
    cases = {
        "max_pattern_len too small (halo would shrink)": patched("max_pattern_len", 0, 8),
        "min_pattern_len wrong": patched("min_pattern_len", 0, 40),
        "filter word size the kernels do not have": patched("filter_w", 0, 6),
        "level-1 bitmap of another size": patched("l1_bits", 0, 1 << 16),
        "level 2 announced but absent": patched("l2_log2", 0, 24),
        "out_off decreasing": patched("out_off", n_out // 2, 0),
        "out_off beyond out_idx": patched("out_off", n_out - 1, 1 << 40),
        "out_idx names a missing pattern": patched("out_idx", 3, 1 << 20),
        "level_off not ascending": patched("level_off", 2, 0),
        "level_edge_off beyond the edges": patched("level_edge_off", lay["level_edge_off"][2] - 1, 1 << 30),
        "edge into row 0": patched("edge_dst", 5, 0),
        "edge class beyond the table": patched("edge_cls", 5, 200),
        "failure link beyond the table": patched("fail", 7, 1 << 30),
        "byte class beyond the table": patched("cls_map", 0x61, 250),
        "root outside the table": patched("root", 0, 1 << 30),
        "rows x classes beyond 2^32": patched("n_classes", 0, 1 << 31),
    }
    for what, blob in cases.items():
        path = tmp_path / "bad.acb"
        path.write_bytes(blob)
        with pytest.raises(native.AcError):
            Automaton.load(str(path), require_device=False)
            pytest.fail(what)


@pytest.mark.gpu
def test_loaded_automaton_matches_like_the_original(tmp_path):
    needles, hay, off = W.cfg2(n_hay=64, hay_len=8192, planted_per_hay=8, seed=3)
    a = Automaton()
    a.add_php_order(needles)
    a.finalize()
    path = str(tmp_path / "cfg2.acb")
    a.save(path)
    b = Automaton.load(path)
    for mode in (1, -1):
        a.set_filter(mode); b.set_filter(mode)
        assert np.array_equal(a.search_events(hay, off), b.search_events(hay, off))
    ha, hb = a.search_hits(hay, off), b.search_hits(hay, off)
    assert np.array_equal(ha, hb) and len(ha) >= 64 * 7
    assert [a.pattern_ordinal(i) for i in range(50)] == [b.pattern_ordinal(i) for i in range(50)]
    rc, got = b.search_callback(hay[:20000].tobytes())
    rc2, got2 = a.search_callback(hay[:20000].tobytes())
    assert (rc, got) == (rc2, got2) and got
