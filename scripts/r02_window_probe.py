import os, sys
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
    for smem in (0, 200*1024, 161*1024, 128*1024, 96*1024):
        a.set_tuning(0, smem)
        best = 1e9
        for _ in range(3):
            _, n = a.search_device_uniform(d.data_ptr(), 512*256, 8192)
            best = min(best, a.stats().kernel_ms)
        print(f"tma={tma} window budget {smem>>10:4d} KiB: {best:.3f} ms events {n}", flush=True)
