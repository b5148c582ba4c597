"""Full walk A/B: argv[1] = alternative library (relative to the repo root) or nothing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from php_aho_corasick_b200 import native
if len(sys.argv) > 1 and sys.argv[1]:
    native.LIB_PATH = os.path.join(ROOT, sys.argv[1])
import ctypes as C
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
a.set_filter(-1)
for planted in (8, 0):
    d = torch.from_numpy(W.cfg2_stream(0, 0, 512, planted_per_hay=planted)).cuda()
    for tma in (-1, 1):
        a.L.acb200_set_tma(C.c_void_p(a.h), C.c_int(tma))
        best = 1e9
        for _ in range(3):
            _, n = a.search_device_uniform(d.data_ptr(), 512 * 256, 8192)
            best = min(best, a.stats().kernel_ms)
        print(f"{sys.argv[1] if len(sys.argv) > 1 else 'default':32s} planted={planted} tma={tma}: {best:.3f} ms events {n}", flush=True)
    del d
