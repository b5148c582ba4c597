"""Mid-size device-resident config-2 batches through the full walk: slice size (set_tuning) x text path (set_tma).
Per synchronous call and per kernel, best of 5 x 10 calls.  One B200."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from php_aho_corasick_b200 import workloads as W          # noqa: E402
from php_aho_corasick_b200.native import Automaton        # noqa: E402

HAY_LEN = 8192
needles, _ = W.cfg2_needles()
aut = Automaton(device=0)
aut.add_php_order(needles)
aut.finalize()
aut.set_filter(-1)
dev = torch.device("cuda", 0)
host = W.cfg2_stream(0, 0, 4)                               # 8 MiB
bufs = {8: torch.from_numpy(host).to(dev),
        0: torch.from_numpy(np.random.default_rng(7).integers(97, 103, size=host.size, dtype=np.uint8)).to(dev)}
stream = torch.cuda.current_stream().cuda_stream
for planted in (8, 0):
    for mib4 in (1, 4, 16):
        n_hay = mib4 * 32
        for tma in (-1, 1):
            aut.set_tma(tma)
            out = []
            for chunk in (0, 16, 32, 48, 64):
                aut.set_tuning(chunk, 0)
                for _ in range(3):
                    n = aut.search_device_uniform(bufs[planted].data_ptr(), n_hay, HAY_LEN, stream=stream)[1]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best, kbest = 1e9, 1e9
                for _ in range(5):
                    e0.record()
                    for _ in range(10):
                        n = aut.search_device_uniform(bufs[planted].data_ptr(), n_hay, HAY_LEN, stream=stream)[1]
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / 10)
                    kbest = min(kbest, aut.stats().kernel_ms)
                out.append(f"chunk {aut.stats().chunk_bytes:4d}: call {best * 1e3:5.1f} kernel {kbest * 1e3:5.1f}")
            print(f"planted {planted} {mib4 / 4:5.2f} MiB tma {tma:2d} events {n:5d} | " + " | ".join(out), flush=True)
