"""Quick device-resident throughput probe (not the bench): prints scan-kernel ms and GB/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton


def run(name, aut, dev_buf, offsets, reps=5, **tune):
    aut.set_tuning(tune.get("chunk", 0), tune.get("smem", 0))
    best = 1e9
    ne = 0
    for _ in range(reps):
        _, ne = aut.search_device(dev_buf.data_ptr(), offsets)
        best = min(best, aut.stats().kernel_ms)
    st = aut.stats()
    gbs = dev_buf.numel() / best / 1e6
    print(f"{name:34s} chunk={st.chunk_bytes:5d} halo={st.halo_bytes:4d} events={ne:10d} "
          f"kernel={best:8.3f} ms  {gbs:8.1f} GB/s", flush=True)


def main():
    torch.cuda.init()
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]
    if "cfg2" in which:
        needles, hay, off = W.cfg2(n_hay=256, hay_len=8192)
        a = Automaton(0); a.add_php_order(needles); a.finalize()
        inf = a.info()
        print("cfg2 automaton: states", inf.n_states, "classes", inf.n_classes, "entry", inf.entry_bytes, "table", inf.table_bytes)
        small = torch.from_numpy(hay).to(dev)
        run("cfg2 256x8KiB (2 MiB)", a, small, off)
        reps = (1 << 30) // hay.size
        big = small.repeat(reps)
        boff = W.offsets_uniform(reps * 256, 8192)
        for chunk in (0, 128, 256, 512, 1024, 2048, 4096):
            run("cfg2 x512 (1 GiB)", a, big, boff, chunk=chunk)
        for smem in (16384, 65536, 131072):
            run(f"cfg2 1 GiB smem={smem}", a, big, boff, smem=smem)
        del big, small
    if "cfg3" in which:
        t0 = time.time()
        pats, hay, off = W.cfg3(hay_bytes=256 << 20)
        t1 = time.time()
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        t2 = time.time()
        inf = a.info()
        print(f"cfg3 automaton: states {inf.n_states} classes {inf.n_classes} entry {inf.entry_bytes} table {inf.table_bytes/1e9:.2f} GB; gen {t1-t0:.1f}s finalize {t2-t1:.1f}s")
        d = torch.from_numpy(hay).to(dev)
        for chunk in (0, 512, 1024, 4096):
            run("cfg3 256 MiB", a, d, off, chunk=chunk)
        run("cfg3 256 MiB smem=4096", a, d, off, smem=4096)
        del d
    if "cfg5" in which:
        pats, hay, off = W.cfg5(hay_bytes=64 << 20)
        a = Automaton(0); a.add_php_order(pats); a.finalize()
        d = torch.from_numpy(hay).to(dev)
        for chunk in (0, 16384, 65536):
            run("cfg5 64 MiB", a, d, off, reps=3, chunk=chunk)


if __name__ == "__main__":
    main()
