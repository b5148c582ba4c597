"""ncu target: two prefilter-path scans of 1 GiB of config 2 (the second is the one to read)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
d = torch.from_numpy(W.cfg2_stream(0, 0, 512)).cuda()
for _ in range(2):
    _, n = a.search_device_uniform(d.data_ptr(), 512 * 256, 8192)
    print(f"events={n} kernel={a.stats().kernel_ms:.3f} ms", flush=True)
