"""Device-resident A/B of the direct verification (gram_table.hpp): per-stage ms with flagged words decided by one
comparison inside the walk kernel (1), through the fused filter + collect pass with staged windows (2), or walked (-1), config 2 at 1 GiB (8 and 0 planted needles) and the config-3 shape at 256 MiB (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton


def run(name, aut, dev_buf, offsets, direct, reps=6):
    aut.set_filter(1)
    aut.set_direct(direct)
    best = None
    for _ in range(reps):
        _, ne = aut.search_device(dev_buf.data_ptr(), offsets)
        st = aut.stats()
        if best is None or st.kernel_ms < best[0]:
            best = (st.kernel_ms, st.filter_ms, st.verify_ms, st.reorder_ms, st.flagged_words, ne)
    k, f, v, r, fl, ne = best
    n = dev_buf.numel()
    print(f"{name:22s} direct={direct:2d} events={ne:9d} flagged={fl:9d} kernels={k:7.3f} ms ({n/k/1e6:7.1f} GB/s) "
          f"filter={f:6.3f} verify={v:6.3f} reorder={r:6.3f}", flush=True)


def main():
    torch.cuda.init()
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["cfg2", "cfg3"]
    if "cfg2" in which:
        for planted in (8, 0):
            needles, hay, off = W.cfg2(n_hay=256, hay_len=8192, planted_per_hay=planted)
            a = Automaton(0); a.add_php_order(needles); a.finalize()
            inf = a.info()
            print(f"cfg2 planted={planted}: direct_keys={inf.direct_keys} walk_keys={inf.direct_walk_keys}")
            big = torch.from_numpy(hay).to(dev).repeat((1 << 30) // hay.size)
            boff = W.offsets_uniform(big.numel() // 8192, 8192)
            for d in (1, 2, -1, 1):
                run("cfg2 1 GiB", a, big, boff, d)
            del big
    if "cfg3" in which:
        pats, hay, off = W.cfg3(hay_bytes=256 << 20)
        t1 = time.time()
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        inf = a.info()
        print(f"cfg3: finalize {time.time()-t1:.1f}s direct_keys={inf.direct_keys} walk_keys={inf.direct_walk_keys}")
        d = torch.from_numpy(hay).to(dev)
        for m in (1, 2, -1):
            run("cfg3 256 MiB", a, d, off, m, reps=3)


if __name__ == "__main__":
    main()
