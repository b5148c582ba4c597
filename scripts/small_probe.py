"""Kernel time of the literal 256 x 8 KiB batch under different slice / window settings (not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
needles, hay, off = W.cfg2()
a = Automaton(0); a.add_php_order(needles); a.finalize()
d = torch.from_numpy(hay).cuda()
for mode in (-1, 1):
    a.set_filter(mode)
    for chunk, smem in ((0, 0), (64, 0), (128, 0), (256, 0), (512, 0), (0, 65536), (0, 32768), (128, 65536), (256, 32768), (256, 16384)):
        a.set_tuning(chunk, smem)
        best = 1e9
        for _ in range(20):
            a.search_device_uniform(d.data_ptr(), 256, 8192)
            best = min(best, a.stats().kernel_ms)
        st = a.stats()
        print(f"filter={mode:2d} chunk={chunk:4d}->{st.chunk_bytes:4d} smem={smem:6d} kernel={best*1e3:7.1f} us launches={st.kernel_launches}")
        if mode == 1 and chunk: break
