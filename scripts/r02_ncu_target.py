"""ncu target: a few prefilter-path scans of 1 GiB of config 2 — argv[1] = set_direct modes, e.g. "1,2"; argv[2] = full for the full walk too."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

modes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2").split(",")]
needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()
n_blocks = int(os.environ.get("BLOCKS", "512"))
d = torch.from_numpy(W.cfg2_stream(0, 0, n_blocks)).cuda()
for mode in modes:
    a.set_direct(mode)
    for _ in range(2):
        _, n = a.search_device_uniform(d.data_ptr(), n_blocks * 256, 8192)
        st = a.stats()
        print(f"direct={mode} events={n} kernel={st.kernel_ms:.3f} filter={st.filter_ms:.3f} verify={st.verify_ms:.3f} reorder={st.reorder_ms:.3f}", flush=True)
if len(sys.argv) > 2 and sys.argv[2] == "full":
    a.set_filter(-1)
    for _ in range(2):
        _, n = a.search_device_uniform(d.data_ptr(), n_blocks * 256, 8192)
        print(f"full walk events={n} kernel={a.stats().kernel_ms:.3f}", flush=True)
