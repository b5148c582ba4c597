"""ac_trie_search_batch() on 1 GiB of separately allocated pageable strings (what ahocorasick_match_batch() calls):
GB/s against the number of gather threads (ACB200_GATHER_THREADS).  One B200."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from php_aho_corasick_b200 import workloads as W          # noqa: E402
from php_aho_corasick_b200.native import Automaton        # noqa: E402

HAY_LEN, BLOCK = 8192, 256
n_blocks = int(os.environ.get("BLOCKS", "512"))
needles, _ = W.cfg2_needles()
aut = Automaton(device=0)
aut.add_php_order(needles)
aut.finalize()
host = W.cfg2_stream(0, 0, n_blocks)
n = n_blocks * BLOCK
strings = [host[i * HAY_LEN:(i + 1) * HAY_LEN].copy() for i in range(n)]
texts = aut.make_texts(strings)
print("cores", os.cpu_count(), "haystacks", n)
ref = None
for nt, thr in [(1, None), (0, None), (1, 4), (0, 4), (1, 8), (0, 8), (1, 12), (0, 12), (1, 16), (0, 16), (1, 24)]:
    os.environ["ACB200_GATHER_NT"] = str(nt)
    if thr is None:
        os.environ.pop("ACB200_GATHER_THREADS", None)
    else:
        os.environ["ACB200_GATHER_THREADS"] = str(thr)
    tb = aut.search_batch_tally(texts=texts)
    best = 1e9
    for _ in range(4):
        t0 = time.time()
        tb = aut.search_batch_tally(texts=texts)
        best = min(best, time.time() - t0)
    if ref is None:
        ref = tb.events
    assert tb.events == ref
    st = aut.stats()
    print(f"nt {nt} gather threads {thr}: {best * 1e3:8.2f} ms  {n * HAY_LEN / best / 1e9:6.1f} GB/s  events {tb.events}  h2d_ms {st.h2d_ms:.2f}")
