"""ncu target: cfg2 dictionary over N MiB, device resident, prefilter path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
planted = int(sys.argv[4]) if len(sys.argv) > 4 else 8
needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=planted)
a = Automaton(0); a.add_php_order(needles); a.finalize(); a.set_filter(mode)
if len(sys.argv) > 5: a.set_parts(int(sys.argv[5]))
if len(sys.argv) > 6: a.set_direct(int(sys.argv[6]))
k = (mib << 20) // hay.size
big = torch.from_numpy(hay).to("cuda:0").repeat(k)
boff = W.offsets_uniform(k * 256, 8192)
for _ in range(reps):
    _, ne = a.search_device(big.data_ptr(), boff)
    st = a.stats()
    print(f"{mib} MiB events={ne} kernel={st.kernel_ms:.3f} filter={st.filter_ms:.3f} reorder={st.reorder_ms:.3f} verify={st.verify_ms:.3f} items={0}")
