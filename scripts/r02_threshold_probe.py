"""Where the gram prefilter starts to pay: device-resident config-2 batches of 0.25 .. 64 MiB through the full walk and
through the prefilter (forced), wall time per synchronous call (CUDA events around it) and kernel time.  One B200."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from php_aho_corasick_b200 import workloads as W          # noqa: E402
from php_aho_corasick_b200.native import Automaton        # noqa: E402

HAY_LEN, BLOCK = 8192, 256
needles, _ = W.cfg2_needles()
aut = Automaton(device=0)
aut.add_php_order(needles)
aut.finalize()
dev = torch.device("cuda", 0)
host = W.cfg2_stream(0, 0, 32)                              # 64 MiB
resident = torch.from_numpy(host).to(dev)
stream = torch.cuda.current_stream().cuda_stream
for planted in (8, 0):
    if planted == 0:
        rng = np.random.default_rng(7)
        resident = torch.from_numpy(rng.integers(97, 103, size=host.size, dtype=np.uint8)).to(dev)
    for mib4 in (1, 2, 4, 8, 16, 32, 64, 128, 256):            # quarter MiB units
        n_hay = mib4 * 32
        row = []
        for mode in (-1, 1):
            aut.set_filter(mode)
            for _ in range(3):
                n = aut.search_device_uniform(resident.data_ptr(), n_hay, HAY_LEN, stream=stream)[1]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best, kbest = 1e9, 1e9
            for _ in range(5):
                e0.record()
                for _ in range(10):
                    n = aut.search_device_uniform(resident.data_ptr(), n_hay, HAY_LEN, stream=stream)[1]
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 10)
                kbest = min(kbest, aut.stats().kernel_ms)
            row.append((best, kbest, n, aut.stats().filtered))
        (wb, wk, wn, wf), (fb, fk, fn, ff) = row
        assert wn == fn and wf == 0 and ff == 1, row
        print(f"planted {planted}  {mib4 / 4:6.2f} MiB  full walk: call {wb * 1e3:7.1f} us kernel {wk * 1e3:7.1f} us | prefilter: call {fb * 1e3:7.1f} us kernel {fk * 1e3:7.1f} us | events {wn}")
