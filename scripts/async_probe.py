"""Stage timings of the synchronous and the asynchronous device call on the same 1 GiB batch (not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton

needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=8)
a = Automaton(0); a.add_php_order(needles); a.finalize()
big = torch.from_numpy(hay).to("cuda:0").repeat((1 << 30) // hay.size)
n_hay = big.numel() // 8192
stream = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    _, n = a.search_device_uniform(big.data_ptr(), n_hay, 8192, stream=stream)
buf = torch.zeros((n + 4096, 2), dtype=torch.int32, device="cuda:0")
for mode in ("sync", "async", "sync", "async"):
    torch.cuda.synchronize()
    e0.record()
    if mode == "sync":
        _, n2 = a.search_device_uniform(big.data_ptr(), n_hay, 8192, stream=stream)
    else:
        assert a.search_device_uniform_async(big.data_ptr(), n_hay, 8192, buf.data_ptr(), n + 4095, stream=stream)
    e1.record()
    if mode == "async":
        n2 = int(buf[0, 0].cpu())                     # stream-ordered read, no device-wide wait before it
        a.async_finish(n2, int(buf[0, 1].cpu()))
    torch.cuda.synchronize()
    st = a.stats()
    print(f"{mode:5s} events={n2} wall(cuda events)={e0.elapsed_time(e1):.3f} ms  kernel={st.kernel_ms:.3f} filter={st.filter_ms:.3f} verify={st.verify_ms:.3f} reorder={st.reorder_ms:.3f}", flush=True)
