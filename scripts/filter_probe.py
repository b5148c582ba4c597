"""Device-resident probe of the prefilter path: per-kernel ms and GB/s, filter on vs off (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton


def run(name, aut, dev_buf, offsets, mode, reps=5):
    aut.set_filter(mode)
    best = None
    for _ in range(reps):
        _, ne = aut.search_device(dev_buf.data_ptr(), offsets)
        st = aut.stats()
        if best is None or st.kernel_ms < best[0]:
            best = (st.kernel_ms, st.filter_ms, st.verify_ms, st.flagged_words, st.dense_tiles, st.filtered, ne, st.reorder_ms, 0)
    k, f, v, fl, dt, fi, ne, c, wi = best
    n = dev_buf.numel()
    print(f"{name:28s} filter={mode:2d} used={fi} events={ne:9d} flagged={fl:10d} dense_tiles={dt:7d} "
          f"kernel={k:8.3f} ms ({n/k/1e6:8.1f} GB/s)  filter={f:7.3f} ms ({(n/f/1e6) if f else 0:8.1f} GB/s) reorder={c:6.3f} verify={v:6.3f} ms items={wi}",
          flush=True)


def main():
    torch.cuda.init()
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["cfg2", "cfg3"]
    if "cfg2" in which:
        for planted in (8, 0):
            needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=planted)
            a = Automaton(0); a.add_php_order(needles); a.finalize()
            inf = a.info()
            print(f"cfg2 planted={planted}: states {inf.n_states} W={inf.filter_word} l1_fill={inf.filter_l1_fill:.4f} l2_log2={inf.filter_l2_log2}")
            small = torch.from_numpy(hay).to(dev)
            reps = (1 << 30) // hay.size
            big = small.repeat(reps)
            boff = W.offsets_uniform(reps * 256, 8192)
            run("cfg2 1 GiB", a, big, boff, 1)
            run("cfg2 1 GiB", a, big, boff, -1)
            run("cfg2 2 MiB", a, small, off, 1)
            run("cfg2 2 MiB", a, small, off, -1)
            mid = small.repeat(32)
            run("cfg2 64 MiB", a, mid, W.offsets_uniform(32 * 256, 8192), 1)
            run("cfg2 64 MiB", a, mid, W.offsets_uniform(32 * 256, 8192), -1)
            del big, small, mid
    if "cfg3" in which:
        t0 = time.time()
        pats, hay, off = W.cfg3(hay_bytes=256 << 20)
        t1 = time.time()
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        t2 = time.time()
        inf = a.info()
        print(f"cfg3: states {inf.n_states} table {inf.table_bytes/1e9:.2f} GB W={inf.filter_word} l1_fill={inf.filter_l1_fill:.4f} "
              f"l2_log2={inf.filter_l2_log2}; gen {t1-t0:.1f}s finalize {t2-t1:.1f}s")
        d = torch.from_numpy(hay).to(dev)
        run("cfg3 256 MiB", a, d, off, 1, reps=3)
        run("cfg3 256 MiB", a, d, off, -1, reps=2)
        del d
    if "cfg5" in which:
        pats, hay, off = W.cfg5(hay_bytes=64 << 20)
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        d = torch.from_numpy(hay).to(dev)
        run("cfg5 64 MiB", a, d, off, 0, reps=3)


if __name__ == "__main__":
    main()
