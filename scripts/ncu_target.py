"""Small fixed workload for ncu captures: cfg2 dictionary over N MiB of planted abcdef haystacks, device resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
planted = int(sys.argv[4]) if len(sys.argv) > 4 else 8
needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=planted)
a = Automaton(0); a.add_php_order(needles); a.finalize()
ilp = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a.set_tuning(chunk, 0, ilp)
k = (mib << 20) // hay.size
big = torch.from_numpy(hay).to("cuda:0").repeat(k)
boff = W.offsets_uniform(k * 256, 8192)
for _ in range(reps):
    _, ne = a.search_device(big.data_ptr(), boff)
    st = a.stats()
    print(f"{mib} MiB ilp={st.ilp} chunk={st.chunk_bytes} events={ne} kernel={st.kernel_ms:.3f} ms {big.numel()/st.kernel_ms/1e6:.1f} GB/s")
