"""Host calls on pageable haystacks of growing size: the direct path staged by the handle's helper threads
(ACB200_STAGE_MIN=1) against cudaMemcpyAsync on the pageable pointer (ACB200_STAGE_MIN huge); scattered strings
(ac_trie_search_batch) likewise.  One B200."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from php_aho_corasick_b200 import workloads as W          # noqa: E402
from php_aho_corasick_b200.native import Automaton        # noqa: E402

needles, _ = W.cfg2_needles()
aut = Automaton(device=0)
aut.add_php_order(needles)
aut.finalize()
big = W.cfg2_stream(0, 0, 16, planted_per_hay=1)           # 32 MiB, one needle per 8 KiB (few events: the wrapper's 4,096-event first try fits up to 32 MiB)
print("cores", os.cpu_count())
for kib in (512, 1024, 2048, 4096, 8192, 16384, 32768):
    hay = big[:kib << 10].copy()
    off = np.array([0, hay.size], dtype=np.uint64)
    strings = [hay[i:i + 8192].copy() for i in range(0, hay.size, 8192)]
    texts = aut.make_texts(strings)
    row = []
    for mode, env in (("driver", str(1 << 40)), ("staged", "1")):
        os.environ["ACB200_STAGE_MIN"] = env
        for _ in range(3):
            ev = aut.search_events(hay, off)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            ev = aut.search_events(hay, off)
        dt = (time.perf_counter() - t0) / reps
        st = aut.stats()
        for _ in range(3):
            tb = aut.search_batch_tally(texts=texts)
        t0 = time.perf_counter()
        for _ in range(reps):
            tb = aut.search_batch_tally(texts=texts)
        db = (time.perf_counter() - t0) / reps
        assert tb.events == len(ev)
        row.append(f"{mode}: one text {dt * 1e6:8.1f} us ({hay.size / dt / 1e9:5.1f} GB/s, h2d {st.h2d_ms * 1e3:7.1f} us)  batch {db * 1e6:8.1f} us ({hay.size / db / 1e9:5.1f} GB/s)")
    print(f"{kib:6d} KiB events {len(ev):5d} | " + " | ".join(row), flush=True)
