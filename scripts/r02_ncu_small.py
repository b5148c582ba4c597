"""ncu target: mid-size device-resident config-2 batches (0.25 / 4 MiB) through the full walk, with and without needles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
a.set_filter(-1)
host = W.cfg2_stream(0, 0, 2)
d = torch.from_numpy(host).cuda()
rng = np.random.default_rng(7)
d0 = torch.from_numpy(rng.integers(97, 103, size=host.size, dtype=np.uint8)).cuda()
for buf, name in ((d, "needles"), (d0, "no needles")):
    for n_hay in (32, 512):
        for _ in range(3):
            _, n = a.search_device_uniform(buf.data_ptr(), n_hay, 8192)
        print(f"{name} {n_hay * 8192 / 2**20:.2f} MiB events={n} kernel={a.stats().kernel_ms * 1e3:.1f} us chunk={a.stats().chunk_bytes}", flush=True)
