"""Per-call latency of ac_trie_search() on 8 KiB haystacks (the literal benchmark.php loop), as bench.py measures it."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton, AcText, MATCH_CB, Tally
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
hay = W.cfg2_stream(0, 0, 1)
L = a.L
cb = ctypes.cast(L.acb200_tally_match_cb, MATCH_CB)
texts = []
for i in range(256):
    t = AcText(); t.astring = hay.ctypes.data + i * 8192; t.length = 8192; texts.append(t)
def loop():
    tally = Tally()
    for t in texts:
        L.ac_trie_search(a.h, ctypes.byref(t), 0, cb, ctypes.cast(ctypes.byref(tally), ctypes.c_void_p))
    return tally
loop()
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); tl = loop(); best = min(best, time.perf_counter() - t0)
print(f"{best / 256 * 1e6:.2f} us per call, events {tl.events}")
