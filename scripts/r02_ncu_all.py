"""ncu target of round 2: the dominant kernels of every BASELINE config, two scans each (the first warms up).
   ncu --set full --clock-control none --import-source on -k regex:^ac_ -o gpurun_out/r02_all python scripts/r02_ncu_all.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

dev = torch.device("cuda:0")
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
n_blocks = 512
d = torch.from_numpy(W.cfg2_stream(0, 0, n_blocks)).to(dev)
for what in ("prefilter", "full walk"):
    a.set_filter(0 if what == "prefilter" else -1)
    for _ in range(2):
        _, n = a.search_device_uniform(d.data_ptr(), n_blocks * 256, 8192)
        st = a.stats()
        print(f"config 2, 1 GiB, {what}: events={n} kernel={st.kernel_ms:.3f} ms", flush=True)
del d
a.release()

pats, hay, off = W.cfg3(hay_bytes=1 << 30)
a = Automaton(0); a.add_php_order(pats); a.finalize()
d = torch.from_numpy(hay).to(dev)
for _ in range(2):
    _, n = a.search_device(d.data_ptr(), off)
    print(f"config 3, 1 GiB, prefilter: events={n} kernel={a.stats().kernel_ms:.3f} ms", flush=True)
del d
a.release()

pats, _, _ = W.cfg5(hay_bytes=16)
a = Automaton(0); a.add_php_order(pats); a.finalize()
n5 = 256 << 20
d = torch.full((n5,), ord("a"), dtype=torch.uint8, device=dev)
for _ in range(2):
    _, n = a.search_device(d.data_ptr(), np.array([0, n5], dtype=np.uint64))
    print(f"config 5, 256 MiB: events={n} kernel={a.stats().kernel_ms:.3f} ms", flush=True)
