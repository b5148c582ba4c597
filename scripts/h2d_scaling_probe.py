"""Bare host->device copy scaling on this box: n GPUs copy 1 GiB each from pinned host memory AT THE SAME TIME
(one process, one stream per GPU, no kernels).  Names the link that bounds the one-process multi-GPU e2e figure
(bench.py `e2e`): if the aggregate stops growing with n, the limit is on the host side (memory channels / PCIe root
ports shared by several GPUs), not in the library's pipeline.

    python scripts/h2d_scaling_probe.py [GiB per GPU] > profiles/rNN_h2d_scaling.txt
"""
import sys
import time

import torch


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    n_all = torch.cuda.device_count()
    nbytes = int(gib * (1 << 30))
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n_all)]
    for h in host:
        h.fill_(1)
    dev = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_all)]
    streams = [torch.cuda.Stream(device=d) for d in range(n_all)]
    print(f"{n_all} GPU(s), {gib} GiB per GPU per copy, pinned host buffers (one per GPU), best of 3")
    ns = [n for n in (1, 2, 4, 8) if n <= n_all]
    for chunk_mib in (0, 64):
        for n in ns:
            best_wall, best_slowest = 1e9, 1e9
            for _ in range(3):
                evs = []
                for d in range(n):
                    torch.cuda.synchronize(d)
                t0 = time.perf_counter()
                for d in range(n):
                    with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                        e0 = torch.cuda.Event(enable_timing=True)
                        e1 = torch.cuda.Event(enable_timing=True)
                        e0.record()
                        if chunk_mib:
                            step = chunk_mib << 20
                            for o in range(0, nbytes, step):
                                dev[d][o:o + step].copy_(host[d][o:o + step], non_blocking=True)
                        else:
                            dev[d].copy_(host[d], non_blocking=True)
                        e1.record()
                        evs.append((e0, e1))
                for d in range(n):
                    torch.cuda.synchronize(d)
                wall = time.perf_counter() - t0
                slowest = max(a.elapsed_time(b) for a, b in evs) / 1e3
                best_wall = min(best_wall, wall)
                best_slowest = min(best_slowest, slowest)
            tot = n * nbytes / 1e9
            what = f"{chunk_mib} MiB pieces" if chunk_mib else "one copy"
            print(f"n={n} ({what}): slowest GPU {best_slowest * 1e3:7.2f} ms = {nbytes / 1e9 / best_slowest:6.1f} GB/s per GPU; "
                  f"wall {best_wall * 1e3:7.2f} ms = {tot / best_wall:6.1f} GB/s aggregate", flush=True)
    # which GPUs share an upstream link: every pair at the same time
    if n_all >= 2:
        print("pairs (GB/s aggregate, both copying at once):")
        for a in range(min(n_all, 8)):
            row = []
            for b in range(min(n_all, 8)):
                if a == b:
                    row.append("   -  ")
                    continue
                for d in (a, b):
                    torch.cuda.synchronize(d)
                t0 = time.perf_counter()
                for d in (a, b):
                    with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                        dev[d].copy_(host[d], non_blocking=True)
                for d in (a, b):
                    torch.cuda.synchronize(d)
                row.append(f"{2 * nbytes / 1e9 / (time.perf_counter() - t0):6.1f}")
            print(f"  GPU{a}: " + " ".join(row), flush=True)


if __name__ == "__main__":
    main()
