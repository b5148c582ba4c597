import sys
sys.path.insert(0, "/root/repo")
import ctypes as C
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
d = torch.from_numpy(W.cfg2_stream(0, 0, 512)).cuda()
a.set_filter(-1)
for tma in (-1, 1):
    a.L.acb200_set_tma(C.c_void_p(a.h), C.c_int(tma))
    for _ in range(2):
        _, n = a.search_device_uniform(d.data_ptr(), 512 * 256, 8192)
        print(f"tma={tma}: {a.stats().kernel_ms:.3f} ms events {n}", flush=True)
