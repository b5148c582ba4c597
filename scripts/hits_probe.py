"""Timing of the device-side hit expansion on the bench batch (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=8)
a = Automaton(0); a.add_php_order(needles); a.finalize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 128          # x 2 MiB
flat = np.tile(hay, reps)
offsets = W.offsets_uniform(256 * reps, 8192)
for _ in range(3):
    t0 = time.time()
    hits = a.search_hits(flat, offsets)
    dt = time.time() - t0
    st = a.stats()
    print(f"{flat.size >> 20} MiB: {len(hits)} hits, call {dt*1e3:.1f} ms, kernels {st.kernel_ms:.3f} ms, expand {st.expand_ms:.3f} ms "
          f"({len(hits)/st.expand_ms/1e6:.2f} G hits/s), h2d {st.h2d_ms:.1f} ms")
pats, hay5, off5 = W.cfg5(n_patterns=1100, hay_bytes=1 << 16)
b = Automaton(0); b.add_php_order(pats); b.finalize()
for _ in range(2):
    t0 = time.time(); hits = b.search_hits(hay5, off5); dt = time.time() - t0
    st = b.stats()
    print(f"cfg5 64 KiB: {len(hits)} hits, call {dt*1e3:.1f} ms, expand {st.expand_ms:.3f} ms ({len(hits)/st.expand_ms/1e6:.2f} G hits/s)")
