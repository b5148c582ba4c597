"""Mid-size device-resident config-2 batches through the full walk with the table window capped at N bytes
(set_tuning): does a small input want a small window (less to stage, more L2 trips)?  One B200."""
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
host = W.cfg2_stream(0, 0, 8)                               # 16 MiB
bufs = {8: torch.from_numpy(host).to(dev),
        0: torch.from_numpy(np.random.default_rng(7).integers(97, 103, size=host.size, dtype=np.uint8)).to(dev)}
stream = torch.cuda.current_stream().cuda_stream
for planted in (8, 0):
    for mib4 in (1, 4, 16, 64):
        n_hay = mib4 * 32
        out = []
        for smem in (0, 16 << 10, 32 << 10, 64 << 10, 128 << 10):
            aut.set_tuning(0, smem)
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
            out.append(f"{smem >> 10:3d}K: call {best * 1e3:5.1f} kernel {kbest * 1e3:5.1f}")
        print(f"planted {planted} {mib4 / 4:5.2f} MiB events {n:6d} | " + " | ".join(out), flush=True)

# host calls: ac_trie_search-style calls on pageable haystacks of growing size (prefilter automatic), per call
import time
aut.set_tuning(0, 0)
aut.set_filter(0)
big = W.cfg2_stream(0, 0, 64)                               # 128 MiB, 8 needles per 8 KiB
for kib in (8, 32, 64, 256, 1024, 4096, 16384, 32768, 131072):
    hay = big[:kib << 10].copy()
    off = np.array([0, hay.size], dtype=np.uint64)
    for _ in range(3):
        ev = aut.search_events(hay, off)
    reps = 20 if kib <= 4096 else 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ev = aut.search_events(hay, off)
    dt = (time.perf_counter() - t0) / reps
    st = aut.stats()
    print(f"host call {kib:7d} KiB: {dt * 1e6:9.1f} us  {hay.size / dt / 1e9:6.2f} GB/s  events {len(ev):7d}  kernel {st.kernel_ms * 1e3:7.1f} us h2d {st.h2d_ms * 1e3:8.1f} us filtered {st.filtered} devices {st.devices}", flush=True)
