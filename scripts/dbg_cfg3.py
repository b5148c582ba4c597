import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
pats, hay, off = W.cfg3(n_patterns=20_000, hay_bytes=8 << 20, plant_every=1 << 16)
a = Automaton(0); a.add_php_order(pats); a.finalize()
inf = a.info(); print("W", inf.filter_word, "l2", inf.filter_l2_log2, "fill", inf.filter_l1_fill, "Lmax", inf.max_pattern_len)
a.set_filter(1); e1 = a.search_events(hay, off); s1 = a.stats()
a.set_filter(-1); e2 = a.search_events(hay, off)
print(len(e1), len(e2), "flagged", s1.flagged_words, "dense", s1.dense_tiles)
s1_ = set((int(x["end"]), int(x["state"])) for x in e1); s2_ = set((int(x["end"]), int(x["state"])) for x in e2)
print("only filter:", sorted(s1_ - s2_)[:5], "only full:", sorted(s2_ - s1_)[:5])
ends = e1["end"]; 
dup = ends[1:][ends[1:] == ends[:-1]]
print("duplicate ends:", dup[:5])
for end, st in sorted(s1_ - s2_)[:3]:
    print("event", end, st, a.state_patterns(st)[:3], "text", hay[end-20:end+4].tobytes().hex(), "end%4", end % 4, "end%512", end % 512, "end%16384", end % 16384)
for d in dup[:3]:
    print("dup at", int(d), "mod4", int(d) % 4, "mod 512", int(d) % 512, [ (int(x["end"]), int(x["state"])) for x in e1 if int(x["end"]) == int(d)])
