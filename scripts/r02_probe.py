"""Round-2 device-resident probe (not the bench): kernel split of the prefilter path with the flagged words settled
inside the filter pass (default) vs inside the walk kernel (set_direct(2)); the full walk; config 5 slice sweep."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton


def split(name, a, scan, reps=6):
    best = None
    for _ in range(reps):
        n = scan()
        st = a.stats()
        row = (st.kernel_ms, st.filter_ms, st.verify_ms, st.reorder_ms, n, st.flagged_words, st.dense_tiles, st.filtered)
        if best is None or row[0] < best[0]:
            best = row
    k, f, v, r, n, fw, dt, fl = best
    print(f"{name:52s} kernel {k:7.3f} ms = filter {f:6.3f} + verify {v:6.3f} + reorder {r:6.3f}   events {n:9d} flagged {fw:9d} dense {dt:5d} filtered {fl}", flush=True)
    return best


def main():
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["cfg2", "cfg3", "walk", "cfg5"]
    needles, _ = W.cfg2_needles()
    if "cfg2" in which or "walk" in which:
        a = Automaton(0); a.add_php_order(needles); a.finalize()
        n_blocks = 512
        for planted in (8, 0):
            host = W.cfg2_stream(0, 0, n_blocks, planted_per_hay=planted)
            d = torch.from_numpy(host).to(dev)
            n_h = n_blocks * 256
            scan = lambda: a.search_device_uniform(d.data_ptr(), n_h, 8192)[1]
            if "cfg2" in which:
                for mode in (1, -1):
                    a.set_direct(mode)
                    split(f"cfg2 1 GiB planted={planted} direct={mode}", a, scan)
                a.set_direct(0)
                # the same bytes as ONE haystack (matches may straddle former boundaries)
                one = np.array([0, d.numel()], dtype=np.uint64)
                split(f"cfg2 1 GiB planted={planted} one haystack", a, lambda: a.search_device(d.data_ptr(), one)[1])
            if "walk" in which and planted == 8:
                a.set_filter(-1)
                for chunk in (0, 256, 1024):
                    a.set_tuning(chunk, 0)
                    split(f"cfg2 1 GiB full walk chunk={chunk}", a, scan, reps=3)
                a.set_tuning(0, 0)
                import ctypes as C
                for mode in (-1, 1):
                    a.L.acb200_set_tma(C.c_void_p(a.h), C.c_int(mode))
                    for chunk in (0, 256, 1024):
                        a.set_tuning(chunk, 0)
                        split(f"cfg2 1 GiB full walk tma={mode} chunk={chunk}", a, scan, reps=3)
                a.set_tuning(0, 0)
                a.L.acb200_set_tma(C.c_void_p(a.h), C.c_int(0))
                a.set_filter(0)
            del d
    if "cfg3" in which:
        pats, hay, off = W.cfg3(hay_bytes=256 << 20)
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        d = torch.from_numpy(hay).to(dev)
        for mode in (1,):
            a.set_direct(mode)
            split(f"cfg3 256 MiB direct={mode}", a, lambda: a.search_device(d.data_ptr(), off)[1])
        del d
    if "cfg5" in which:
        pats, _, _ = W.cfg5(hay_bytes=16)
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        n = 256 << 20
        d = torch.full((n,), ord("a"), dtype=torch.uint8, device=dev)
        off = np.array([0, n], dtype=np.uint64)
        for chunk in (0, 2048, 8192, 16384):
            a.set_tuning(chunk, 0)
            k = split(f"cfg5 256 MiB chunk={chunk}", a, lambda: a.search_device(d.data_ptr(), off)[1], reps=3)[0]
            print(f"    read {n / k / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
