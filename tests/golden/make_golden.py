"""Generates tests/golden/phpt_golden.json from the reference's own PHPT tests.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

* Expected outputs are PARSED out of the --EXPECT-- sections (PHP var_dump text) of
  /root/reference/tests/test1..6.phpt, byte-exact (string(n) lengths are honoured).
* Inputs (pattern arrays, haystacks, call sequence) are transcribed below from the
  --FILE-- sections; each block cites the lines it was taken from.
The JSON is committed; the GPU box never reads /root/reference.
"""
import json
import os
import re
import sys

REF = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "phpt_golden.json")


class VarDump:
    """Parser for the subset of var_dump() output the tests print."""

    def __init__(self, data: bytes):
        self.b = data
        self.i = 0

    def ws(self):
        while self.i < len(self.b) and self.b[self.i] in b" \n\r\t":
            self.i += 1

    def value(self):
        self.ws()
        b = self.b
        m = re.compile(rb"int\((-?\d+)\)").match(b, self.i)
        if m:
            self.i = m.end()
            return int(m.group(1))
        m = re.compile(rb"bool\((true|false)\)").match(b, self.i)
        if m:
            self.i = m.end()
            return m.group(1) == b"true"
        m = re.compile(rb'string\((\d+)\) "').match(b, self.i)
        if m:
            n = int(m.group(1))
            s = b[m.end():m.end() + n]
            assert b[m.end() + n:m.end() + n + 1] == b'"', "bad string length"
            self.i = m.end() + n + 1
            return s.decode("utf-8")
        m = re.compile(rb"array\((\d+)\) \{").match(b, self.i)
        if m:
            n = int(m.group(1))
            self.i = m.end()
            keys, vals = [], []
            for _ in range(n):
                self.ws()
                km = re.compile(rb'\[(?:"([^"]*)"|(-?\d+))\]=>').match(b, self.i)
                assert km, b[self.i:self.i + 40]
                self.i = km.end()
                keys.append(km.group(1).decode() if km.group(1) is not None else int(km.group(2)))
                vals.append(self.value())
            self.ws()
            assert b[self.i:self.i + 1] == b"}"
            self.i += 1
            if all(isinstance(k, int) for k in keys) and keys == list(range(len(keys))):
                return vals                      # PHP list
            return {"__order__": keys, **{str(k): v for k, v in zip(keys, vals)}}
        raise ValueError(b[self.i:self.i + 60])


def expect_section(name):
    raw = open(os.path.join(REF, name), "rb").read()
    return raw.split(b"--EXPECT--", 1)[1]


def dumps_in(section: bytes):
    """all top-level var_dump values in order of appearance"""
    out = []
    p = VarDump(section)
    pat = re.compile(rb"(array\(\d+\) \{|bool\((?:true|false)\)|string\(\d+\) \")")
    while True:
        m = pat.search(section, p.i)
        if not m:
            break
        p.i = m.start()
        out.append(p.value())
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tests not found at " + REF)
    g = {}

    # tests/test1.phpt:11-29 (patterns, haystack) and :41-52 (UTF-8 check)
    d = dumps_in(expect_section("test1.phpt"))
    g["test1"] = {
        "source": "tests/test1.phpt:11-52 -> :60-157",
        "cases": [
            {"init": [{"key": "ab", "value": "alfa"}, {"key": "ac", "value": "beta"},
                      {"key": "ad", "value": "gamma", "aux": [1]}, {"key": "ae", "value": "delta"},
                      {"id": 0, "value": "zeta"}, {"key": "ag", "value": "omega"}, {"value": "lfa"}],
             "matches": [{"haystack": "alFABETA gamma zetaomegaalfa!", "expect": d[0]}]},
            {"init": [{"value": "你好"}, {"value": "hi"}, {"value": "谢谢"}, {"value": "thanks"}],
             "matches": [{"haystack": "你好，hi，谢谢，thanks", "expect": d[1]}]},
        ],
    }

    # tests/test2.phpt:11-66
    d = dumps_in(expect_section("test2.phpt"))
    aux1, aux2, aux3 = [["helloAuxObject", 41]], 0x42, "simple-aux"
    g["test2"] = {
        "source": "tests/test2.phpt:11-66 -> :71-298",
        "cases": [
            {"init": [{"key": "ab", "value": "alfa"}, {"key": "ac", "value": "beta"},
                      {"key": "ad", "value": "gamma", "aux": aux2}, {"key": "ae", "value": "delta", "aux": aux3},
                      {"key": "af", "value": "zeta"}, {"key": "ag", "value": "omega"}, {"key": "ah", "value": "lfa"},
                      {"id": 42, "value": "pie"}, {"value": "simple"}, {"value": "aux", "aux": aux1},
                      {"value": "aux2", "aux": aux2}, {"value": "aux3", "aux": aux1},
                      {"value": "ščř+éé"}, {"value": "éé"}],
             "matches": [
                 {"haystack": "alFABETA gammadelta delta delta simple pie! aux ssščř+ééžž ččř é é-é éeéee éé aux2 aux3 aux2",
                  "expect": d[0]},
                 {"haystack": "alFABETAABECEDAAAA!", "expect": d[1]},
                 {"haystack": "alFABETAABECEDAAAA!", "findAll": False, "expect": d[2]},
                 {"haystack": "alFABETAABECEDAAAA!", "findAll": True, "expect": d[3]}],
             "lifecycle": {"isValid": d[4], "deinit": d[5], "isValid_after": d[6], "deinit_again": d[7]}},
        ],
    }

    # tests/test3.phpt:12-24
    d = dumps_in(expect_section("test3.phpt"))
    g["test3"] = {
        "source": "tests/test3.phpt:12-32 -> :34-93",
        "cases": [
            {"init": [],
             "add_patterns": [[{"key": "ab", "value": "alfa"}], [{"key": "ac", "value": "beta"}],
                              [{"key": "ad", "value": "gamma", "aux": [1]}], [{"key": "ae", "value": "delta"}],
                              [{"id": 0, "value": "zeta"}, {"key": "ag", "value": "omega"}, {"value": "lfa"}]],
             "matches": [{"haystack": "alFABETA gamma zetaomegaalfa!", "expect": d[0]}]},
        ],
    }

    # tests/test4.phpt:11-27 — 20 x (init, 1000 x match with exactly 4 hits, deinit)
    g["test4"] = {
        "source": "tests/test4.phpt:11-27",
        "init": [{"value": "a5"}], "haystack": "aoeu a5 a5 a5 a5 aoeu", "outer": 20, "inner": 1000, "hits": 4,
    }

    # tests/test5.phpt:11-38 — no hits, no crash
    strings = [x for x in dumps_in(expect_section("test5.phpt")) if isinstance(x, str)]
    g["test5"] = {
        "source": "tests/test5.phpt:11-38",
        "init": [{"key": "熊本県熊本市北区四方寄町", "value": "北区四方寄町"},
                 {"key": "熊本県熊本市北区立福寺町", "value": "北区立福寺町"}],
        "haystacks": strings,
    }
    assert len(strings) == 15

    # tests/test6.phpt:12-29 — no state carried between calls
    d = dumps_in(expect_section("test6.phpt"))
    g["test6"] = {
        "source": "tests/test6.phpt:12-37 -> :40-73",
        "cases": [
            {"init": [{"key": "a", "value": "abcd"}, {"key": "b", "value": "ghij"},
                      {"key": "c", "value": "defg"}, {"key": "d", "value": "defghijkl"}],
             "matches": [{"haystack": "abcde", "expect": d[0]}, {"haystack": "fghij", "expect": d[1]},
                         {"haystack": "klmno", "expect": d[2]}]},
        ],
    }

    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(g, f, ensure_ascii=False, indent=1)
    n = sum(len(m.get("matches", [])) for t in g.values() for m in t.get("cases", []))
    print(f"wrote {OUT}: {n} golden match arrays")


if __name__ == "__main__":
    main()
