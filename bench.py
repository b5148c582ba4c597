#!/usr/bin/env python
"""bench.py — haystack GB/s matched (bit-exact hits) on B200, next to the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-also]

A step is one pass of the hot path (ahocorasick_match over a batch == ac_trie_search) over the
benchmark.php-shaped batch: the 2,048 x 16-byte `abcdef` dictionary of BASELINE.json config 2 over
8 KiB `abcdef` haystacks with 8 planted needles each, scaled from 256 haystacks (2 MiB, launch-latency
bound) to 131,072 DISTINCT haystacks = 1 GiB per GPU (SURVEY.md §8d "steady-state variant"; block-seeded,
workloads.cfg2_stream — nothing is tiled).

* value      device-timed whole-job throughput, haystacks resident in HBM (CUDA events, max over ranks);
             at N > 1 the step includes the NCCL gather of every rank's events to rank 0
* parity     every haystack's event count + order-sensitive event hash, GPU (at N > 1: the rows rank 0
             gathered) against the reference's own ac_trie_search (oracle/_ref) run on the same bytes; a
             mismatch exits non-zero
* roofline   the device kernels of one step: 1 algorithmic HBM byte per haystack byte / their summed duration
             (library's own CUDA events on the launching stream), against MEASURED_PEAKS.json hbm_gbs
* e2e        ONE process (rank 0) calling ac_trie_search_flat() on a pinned HOST buffer holding the haystacks
             of all N GPUs: slabs -> per-GPU H2D pipelines -> kernels -> D2H events -> callback replay, all
             inside the timed region; e2e_batch is the same through ac_trie_search_batch() on separately
             allocated pageable strings (the PHP extension's ahocorasick_match_batch)
* also       the other BASELINE configs, each with its own kernel time, roofline fraction and CPU figure:
             config 3 (100 k signatures / 1 GiB), config 5 (adversarial, event-write bound), the literal
             benchmark.php loop (256 sequential calls of 8 KiB), config 4 (65,536 x 64 KiB over the N GPUs)
* cpu_baseline / --impl reference   the reference's own ac_trie_search (oracle/_ref, compiled from
             /root/reference) or, where that is absent, the C restatement (oracle/), on host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "haystack GB/s matched (bit-exact hits)"
HAY_LEN = 8192
BLOCK_HAYS = 256                    # the literal benchmark.php batch; the unit the stream is seeded by
HAYS_PER_GPU = 131072               # 1 GiB per GPU
CPU_SAMPLE_BLOCKS = 128             # 256 MiB: the bounded sample of the CPU legs (blocks 0..127 of rank 0's stream)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload_key):
    """dram bytes per step from the committed ncu capture, if one matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload_key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax, "reasons": ["no sample inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU legs ----

def cpu_kind(prefer_reference=True):
    from oracle import pydriver
    if prefer_reference and pydriver.available("reference"):
        return "reference"
    if not pydriver.available("oracle"):
        pydriver.build(("liboracle_driver.so",))
    return "oracle"


def cpu_leg(patterns, flat, off, threads, reps, kind, halo=0, digest=False, what=""):
    """Times the CPU ac_trie_search over the given bytes (finalize excluded). -> dict (+ counts/hashes with digest)"""
    from oracle import pydriver
    sec, events, counts, hashes = pydriver.bench_digest(kind, patterns, flat, off, threads, reps, halo=halo, digest=digest)
    n = len(off) - 1
    shape = (f"{n} haystacks" if n > 1 else f"one haystack in {threads} slices with a {halo}-byte halo" if threads > 1
             else "one haystack")
    return {"value": flat.size / sec / 1e9, "unit": "GB/s", "cores": threads,
            "kind": "reference" if kind == "reference" else "port",
            "sample": f"{what}{shape}, {flat.size >> 20} MiB, best of {reps}, ac_trie_search only (finalize excluded), "
                      "one private trie per thread",
            "events": events, "seconds": sec, "counts": counts, "hashes": hashes}


def public(leg):
    return {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}


def workload_config(n_gpus, bytes_per_gpu):
    return {
        "workload": "BASELINE config 2 (benchmark.php shape) at steady state: 2048 x 16 B needles, alphabet 'abcdef', "
                    f"{bytes_per_gpu // HAY_LEN} distinct haystacks x {HAY_LEN} B per GPU with 8 planted needles per haystack",
        "haystack_bytes_per_gpu": bytes_per_gpu,
        "patterns": 2048, "pattern_len": 16, "alphabet": "abcdef", "planted_per_haystack": 8,
        "l2_policy": "inputs larger than L2 (1 GiB per GPU vs 126 MB)",
        "sharding": f"{n_gpus} x independent haystack blocks, automaton replicated, events gathered to rank 0",
    }


def run_reference(args):
    """The reference arm: the reference's own ac_trie_search on all host cores, each step the same bounded sample
    (the first 256 MiB of rank 0's stream) of the GPU arm's workload."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from php_aho_corasick_b200 import workloads as W
    needles, _ = W.cfg2_needles()
    blocks = min(CPU_SAMPLE_BLOCKS, max(1, args.hays_per_gpu // BLOCK_HAYS))
    flat = W.cfg2_stream(0, 0, blocks)
    off = W.offsets_uniform(blocks * BLOCK_HAYS, HAY_LEN)
    threads = os.cpu_count() or 1
    kind = cpu_kind()
    t0 = time.time()
    per_step, leg = [], None
    for s in range(args.warmup + max(1, args.steps)):
        leg = cpu_leg(needles, flat, off, threads, 1, kind, what="blocks 0..%d of rank 0's stream: " % (blocks - 1))
        if s >= args.warmup:
            per_step.append(leg["seconds"])
        if time.time() - t0 > 240 and per_step:
            break
    sec = float(np.mean(per_step))
    val = flat.size / sec / 1e9
    nbytes = max(1, args.hays_per_gpu // BLOCK_HAYS) * BLOCK_HAYS * HAY_LEN
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": len(per_step), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.gpus, nbytes),
        "cpu_baseline": dict(public(leg), value=val),
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------- GPU helpers ----

def packed_to_events(aut, torch, ptr_n, n_hay, hay_len=None, offsets=None):
    """the library's device event buffer of the last scan -> host structured events (end, state, text_idx)"""
    from php_aho_corasick_b200.native import EVENT_DTYPE
    _, n = ptr_n
    dev = torch.device("cuda", torch.cuda.current_device())
    buf = torch.empty((max(n, 1), 2), dtype=torch.int32, device=dev)
    aut.copy_events(buf.data_ptr(), n, stream=1)           # 1 = cudaStreamLegacy: ordered with torch's default stream
    torch.cuda.synchronize()
    a = buf[:n].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    out = np.empty(n, dtype=EVENT_DTYPE)
    if hay_len:
        h = (a[:, 0] - 1) // hay_len
        out["end"] = a[:, 0] - h * hay_len
    else:
        off = np.asarray(offsets, dtype=np.int64)
        h = np.searchsorted(off, a[:, 0], side="left") - 1
        out["end"] = a[:, 0] - off[h]
    out["state"] = a[:, 1]
    out["text_idx"] = h
    return out


def kernel_split(st_sum, steps, nbytes, peak):
    per = {}
    for name, ms in (("ac_filter_kernel", st_sum["filter_ms"]), ("ac_collect_kernel + ac_walk_kernel", st_sum["verify_ms"]),
                     ("ac_offsets_kernel + ac_emit_kernel", st_sum["reorder_ms"])):
        m = ms / steps
        per[name] = {"ms": m, "GBps": nbytes / (m * 1e-3) / 1e9 if m else None,
                     "frac_of_peak": nbytes / (m * 1e-3) / 1e9 / peak if m else None}
    return per


def timed_device_scans(aut, torch, scan, steps, warmup=3):
    """K device-resident scans: CUDA events around the loop + the library's per-kernel event times. -> dict"""
    for _ in range(warmup):
        scan()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {"kernel_ms": 0.0, "filter_ms": 0.0, "verify_ms": 0.0, "reorder_ms": 0.0, "launches": 0, "filtered": 0}
    e0.record()
    n = 0
    for _ in range(steps):
        n = scan()
        st = aut.stats()
        acc["kernel_ms"] += st.kernel_ms
        acc["filter_ms"] += st.filter_ms
        acc["verify_ms"] += st.verify_ms
        acc["reorder_ms"] += st.reorder_ms
        acc["launches"] += st.kernel_launches
        acc["filtered"] += st.filtered
    e1.record()
    torch.cuda.synchronize()
    acc["ms_per_step"] = e0.elapsed_time(e1) / steps
    acc["events"] = n
    return acc


def also_record(name, fn):
    t0 = time.time()
    try:
        rec = fn()
    except Exception as e:                                   # a side record must never take the headline line down
        rec = {"error": f"{type(e).__name__}: {e}"}
    rec["wall_s"] = round(time.time() - t0, 2)
    return name, rec


# ------------------------------------------------------------------------------------------- main ----

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--hays-per-gpu", type=int, default=HAYS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-filter", action="store_true", help="force the full automaton walk (ac_scan_kernel)")
    ap.add_argument("--skip-also", action="store_true", help="only the headline workload (configs 3, 4, 5 and the literal loop are skipped)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.native import Automaton, EVENT_DTYPE
    from php_aho_corasick_b200.dist import MailboxGatherer, ShardedMatcher, globalize

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")          # waits that must not put a spinning kernel on a GPU

    def host_barrier():
        if host_group is not None:
            dist.barrier(group=host_group)

    needles, _ = W.cfg2_needles()
    aut = Automaton(device=local_rank)
    aut.add_php_order(needles)
    t0 = time.time()
    aut.finalize()
    finalize_s = time.time() - t0
    inf = aut.info()
    if args.no_filter:
        aut.set_filter(-1)

    n_blocks = max(1, args.hays_per_gpu // BLOCK_HAYS)
    n_hays = n_blocks * BLOCK_HAYS
    nbytes = n_hays * HAY_LEN
    offsets = W.offsets_uniform(n_hays, HAY_LEN)
    host_stream = W.cfg2_stream(rank, 0, n_blocks)          # this rank's distinct haystacks
    resident = torch.from_numpy(host_stream).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    sm = ShardedMatcher(aut)
    peak, peak_src = hbm_peak()

    mg = None
    if world > 1:
        # One synchronous step through the NCCL all_gather path sizes the mailbox gather (dist.MailboxGatherer): from then
        # on every rank copies exactly its own rows into rank 0's IPC-mapped buffer with the copy engines, behind its scan.
        n0, _ = sm.scan_and_gather(resident, offsets, 0, stream=stream, uniform_len=HAY_LEN)
        most = torch.tensor([n0], dtype=torch.int64, device=dev)
        dist.all_reduce(most, op=dist.ReduceOp.MAX)
        try:
            mg = MailboxGatherer(aut, cap_rows=2 * int(most.item()) + 4096)
        except Exception as e:                   # (e.g. CUDA IPC not permitted in this container)
            print(f"rank {rank}: mailbox gather unavailable ({e}); using the all_gather path", file=sys.stderr)
        ok = torch.tensor([1 if mg is not None else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and mg is not None:
            mg.close(collective=False)
            mg = None

    def step_resident():
        if world > 1 and mg is None:
            return sm.scan_and_gather(resident, offsets, 0, stream=stream, uniform_len=HAY_LEN)
        if world > 1:
            n = mg.scan_and_send(resident, n_hays, HAY_LEN, stream=stream)
            if mg.step >= 2:
                mg.result(mg.step - 2)           # rank 0: the rows of the step before — they landed while this one ran
            return n, None
        # one GPU: the C-ABI call itself; the sorted events stay in the library's device buffer (its contract)
        return aut.search_device_uniform(resident.data_ptr(), n_hays, HAY_LEN, stream=stream)[1], None

    def drain():
        """the rows of the last step are on rank 0 too before the clock stops"""
        if mg is not None:
            mg.drain(stream)
            return mg.result(mg.step - 1)
        return None

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps: kernel-resident leg ("value")
    for _ in range(args.warmup):
        n_events, _ = step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    acc = {"kernel_ms": 0.0, "filter_ms": 0.0, "verify_ms": 0.0, "reorder_ms": 0.0}
    launches, filtered_steps = 0, 0
    for _ in range(args.steps):
        n_events, _ = step_resident()
        st = aut.stats()
        for k in acc:
            acc[k] += getattr(st, k)
        launches += st.kernel_launches
        filtered_steps += st.filtered
    drain()
    e1.record()
    sync()
    wall1 = time.time()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(wall0, wall1)

    # ---- parity: per-haystack event count + order-sensitive hash, GPU vs the CPU reference on the same bytes.
    #      At N > 1 what is checked is what rank 0 GATHERED (one more step, outside the timed region).
    kind = cpu_kind()
    threads_total = os.cpu_count() or 1
    cpu_full = cpu_leg(needles, host_stream, offsets, max(1, threads_total // world), 1, kind, digest=True)
    n_own, got = step_resident()
    if mg is not None:
        got = drain()
    if world > 1:
        exp = torch.from_numpy(np.stack([cpu_full["counts"], cpu_full["hashes"]]).view(np.int64)).to(dev)
        all_exp = [torch.empty_like(exp) for _ in range(world)]
        dist.all_gather(all_exp, exp)
        parity = None
        if rank == 0:
            goff = W.offsets_uniform(world * n_hays, HAY_LEN)
            ranges = [(r * n_hays, (r + 1) * n_hays) for r in range(world)]
            ev = globalize(got, ranges, goff)
            counts, hashes = aut.event_digest(ev, world * n_hays)
            e_counts = np.concatenate([x[0].cpu().numpy().view(np.uint64) for x in all_exp])
            e_hashes = np.concatenate([x[1].cpu().numpy().view(np.uint64) for x in all_exp])
            bad = int(np.count_nonzero((counts != e_counts) | (hashes != e_hashes)))
            ordered = bool(np.all(np.diff(ev["text_idx"].astype(np.int64)) >= 0))
            parity = {"checked": True, "events": int(ev.size), "haystacks": int(world * n_hays), "hash_ok": bad == 0 and ordered,
                      "mismatching_haystacks": bad, "rows_in_global_order": ordered,
                      "what": ("rows gathered on rank 0 (copy engines over NVLink into its IPC-mapped buffer, dist.MailboxGatherer) "
                               if mg is not None else "rows gathered on rank 0 (NCCL all_gather, dist.ShardedMatcher) ")
                              + f"from {world} ranks vs {cpu_full['kind']} ac_trie_search, every haystack"}
    else:
        ev = packed_to_events(aut, torch, (None, n_own), n_hays, hay_len=HAY_LEN)
        counts, hashes = aut.event_digest(ev, n_hays)
        bad = int(np.count_nonzero((counts != cpu_full["counts"]) | (hashes != cpu_full["hashes"])))
        parity = {"checked": True, "events": int(ev.size), "haystacks": int(n_hays), "hash_ok": bad == 0,
                  "mismatching_haystacks": bad,
                  "what": f"device event list vs {cpu_full['kind']} ac_trie_search, every haystack"}

    # ---- e2e: ONE process drives all N GPUs through the C-ABI from a pinned host buffer (rank 0; the other ranks wait)
    e2e = None
    gathered = None
    if world > 1:                                            # every rank's haystacks -> rank 0 (NVLink), then its host buffer
        gathered = [torch.empty_like(resident) for _ in range(world)] if rank == 0 else None
        dist.gather(resident, gathered, dst=0)
        torch.cuda.synchronize()
    host_barrier()
    if rank == 0:
        L = aut.L
        total_bytes = world * nbytes
        host_ptr = L.acb200_host_alloc(total_bytes)
        if not host_ptr:
            raise RuntimeError("pinned allocation failed")
        host = np.ctypeslib.as_array((ctypes.c_uint8 * total_bytes).from_address(host_ptr))
        host[:nbytes] = host_stream
        for r in range(1, world):
            torch.from_numpy(host[r * nbytes:(r + 1) * nbytes]).copy_(gathered[r])
        gathered = None
        goff = W.offsets_uniform(world * n_hays, HAY_LEN)
        if world > 1:
            aut.set_devices(list(range(world)))
        e2e_steps = max(3, min(args.steps, 8))
        for _ in range(2):
            tally = aut.search_flat_tally(host_ptr, goff)
        t0 = time.time()
        h2d_ms = d2h_ms = 0.0
        for _ in range(e2e_steps):
            tally = aut.search_flat_tally(host_ptr, goff)
            st = aut.stats()
            h2d_ms += st.h2d_ms
            d2h_ms += st.d2h_ms
        e2e_s = (time.time() - t0) / e2e_steps
        st_e2e = aut.stats()
        # the same call at event level, once: the events of all N GPUs, per haystack, against the CPU reference
        ev = aut.search_events(host, goff)
        counts, hashes = aut.event_digest(ev, world * n_hays)
        e2e_ok = None
        if world == 1:
            e2e_ok = bool(np.array_equal(counts, cpu_full["counts"]) and np.array_equal(hashes, cpu_full["hashes"]))
        else:
            e2e_ok = bool(np.array_equal(counts, e_counts) and np.array_equal(hashes, e_hashes))
        assert tally.events == int(counts.sum()), (tally.events, int(counts.sum()))
        # ... and through the scattered-strings entry (ahocorasick_match_batch): pageable, separately allocated haystacks
        n_b = min(world * n_hays, 4 * 32768)                 # 1 GiB of them at most
        strings = [host[i * HAY_LEN:(i + 1) * HAY_LEN].copy() for i in range(n_b)]
        texts = aut.make_texts(strings)
        tb = aut.search_batch_tally(texts=texts)
        t0 = time.time()
        for _ in range(3):
            tb = aut.search_batch_tally(texts=texts)
        batch_s = (time.time() - t0) / 3
        assert tb.events == int(counts[:n_b].sum())
        e2e = {"value": total_bytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(total_bytes),
               "d2h_bytes_per_step": int(tally.events * 8 + 16 * st_e2e.devices), "ms_per_step": e2e_s * 1e3,
               "h2d_ms_slowest_gpu": h2d_ms / e2e_steps, "d2h_ms_slowest_gpu": d2h_ms / e2e_steps, "steps": e2e_steps,
               "gpus_driven_by_this_process": int(st_e2e.devices),
               "api": "ac_trie_search_flat(pinned host buffer) + acb200_tally_cb replay, one process, "
                      f"{st_e2e.devices} GPU pipeline(s) (csrc/shard.hpp)",
               "events": int(tally.events), "hits": int(tally.hits), "parity_vs_cpu_reference": e2e_ok,
               "batch_api": {"value": n_b * HAY_LEN / batch_s / 1e9, "unit": "GB/s", "ms": batch_s * 1e3, "haystacks": n_b,
                             "api": "ac_trie_search_batch(separately allocated pageable strings): gather into pinned "
                                    "staging + the same pipelines — what ahocorasick_match_batch() calls"}}
        if world > 1:
            aut.set_devices([])
        L.acb200_host_free(host_ptr)
        del strings, texts
    host_barrier()
    sync()

    # ---- max over ranks
    t = torch.tensor([elapsed_ms, acc["kernel_ms"], float(launches), acc["filter_ms"], acc["verify_ms"], acc["reorder_ms"]],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0])
        acc["kernel_ms"], acc["filter_ms"], acc["verify_ms"], acc["reorder_ms"] = (float(tmax[i]) for i in (1, 3, 4, 5))
        launches = int(tsum[2])
        # per rank, for the record: the step and kernel times the maxima above were taken from, and every GPU's clock
        mine = torch.tensor([float(t[0]), float(t[1]), float(t[3]), float(t[4]), float(t[5]), clocks.get("sm_mhz") or 0.0,
                             1.0 if clocks.get("reasons") else 0.0], dtype=torch.float64, device=dev)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {"ms_per_step": [round(float(x[0]) / args.steps, 4) for x in every],
                    "kernel_ms": [round(float(x[1]) / args.steps, 4) for x in every],
                    "filter_ms": [round(float(x[2]) / args.steps, 4) for x in every],
                    "collect_walk_ms": [round(float(x[3]) / args.steps, 4) for x in every],
                    "offsets_emit_ms": [round(float(x[4]) / args.steps, 4) for x in every],
                    "sm_mhz": [float(x[5]) for x in every], "throttle_flagged": [bool(x[6]) for x in every]}
    else:
        per_rank = None
    ms_per_step = elapsed_ms / args.steps
    value = world * nbytes / (ms_per_step * 1e-3) / 1e9
    k_ms = acc["kernel_ms"] / args.steps
    achieved = nbytes / (k_ms * 1e-3) / 1e9

    filtered = filtered_steps == args.steps
    if filtered:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic("cfg2_1GiB_filtered"), "peak_source": peak_src,
                    "kernel": "ac_filter_kernel + ac_collect_kernel + ac_walk_kernel + ac_offsets_kernel + ac_emit_kernel "
                              "(the device kernels of one step)",
                    "kernel_ms": k_ms, "algorithmic_bytes_per_launch": nbytes,
                    "per_kernel": kernel_split(acc, args.steps, nbytes, peak),
                    "note": "1 HBM byte per haystack byte over the summed duration of the step's kernels. "
                            "ac_filter_kernel is the only one that streams the haystack (HBM-bound, per_kernel "
                            "shows its own fraction) — see DESIGN.md"}
    else:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic("cfg2_1GiB"), "peak_source": peak_src,
                    "kernel": "ac_scan_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": nbytes,
                    "note": "1 HBM byte per haystack byte; the kernel is bound by dependent shared-memory table "
                            "lookups, not by HBM — see DESIGN.md"}

    # ---- the other BASELINE configs
    also = {}
    if not args.skip_also:
        if world == 1:
            for name, fn in (("config2_full_walk_1GiB", lambda: also_cfg2_full_walk(torch, dev, aut, resident, n_hays, n_events, peak, filtered)),
                             ("config3_signatures_1GiB", lambda: also_cfg3(torch, dev, peak)),
                             ("config5_adversarial_256MiB", lambda: also_cfg5(torch, dev, peak)),
                             ("literal_benchmark_php_loop", lambda: also_literal(torch, dev, aut, needles))):
                k, rec = also_record(name, fn)
                also[k] = rec
        k, rec = also_record("config4_batched_65536x64KiB", lambda: also_cfg4(torch, dist, dev, aut, needles, world, rank, peak))
        if rank == 0:
            also[k] = rec

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": dict(workload_config(world, nbytes),
                       path="gram prefilter + verify" if filtered else "full automaton walk",
                       automaton={"states": int(inf.n_states), "classes": int(inf.n_classes),
                                  "prefilter_word": int(inf.filter_word), "prefilter_l1_fill": float(inf.filter_l1_fill),
                                  "direct_keys": int(inf.direct_keys), "direct_walk_keys": int(inf.direct_walk_keys),
                                  "entry_bytes": int(inf.entry_bytes), "table_bytes": int(inf.table_bytes),
                                  "finalize_s": round(finalize_s, 4)},
                       events_per_step_per_gpu=int(n_events)),
        "parity": parity,
        "per_rank": per_rank,
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
        "also": also,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference arm's sample (blocks 0..127 of this stream), all cores and one core
        nb = min(CPU_SAMPLE_BLOCKS, n_blocks)
        s_flat, s_off = host_stream[: nb * BLOCK_HAYS * HAY_LEN], W.offsets_uniform(nb * BLOCK_HAYS, HAY_LEN)
        cb = cpu_leg(needles, s_flat, s_off, threads_total, 2, kind, what=f"blocks 0..{nb - 1} of rank 0's stream: ")
        nb1 = max(1, nb // 16)
        cb1 = cpu_leg(needles, host_stream[: nb1 * BLOCK_HAYS * HAY_LEN], W.offsets_uniform(nb1 * BLOCK_HAYS, HAY_LEN), 1, 1, kind)
        line["cpu_baseline"] = public(cb)
        line["cpu_baseline"]["single_core_GBps"] = cb1["value"]
        line["cpu_baseline"]["whole_batch_GBps"] = cpu_full["value"]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        if mg is not None:
            mg.close()
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity and not parity["hash_ok"]:
        print("PARITY FAILURE: " + json.dumps(parity), file=sys.stderr)
        return 3
    if rank == 0 and e2e and e2e.get("parity_vs_cpu_reference") is False:
        print("PARITY FAILURE in the e2e path", file=sys.stderr)
        return 3
    return 0


# --------------------------------------------------------------------------------- the other configs ----

def _roof(nbytes, kernel_ms, peak):
    ach = nbytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms else 0.0
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None}


def also_cfg2_full_walk(torch, dev, aut, resident, n_hays, n_events, peak, filtered, steps=5):
    """The headline batch through the FULL automaton walk (ac_scan_kernel / ac_scan_tma_kernel, what dictionaries with a
    pattern shorter than 8 bytes get), and a batch of the same shape without planted needles (text that matches
    nothing keeps every lane in the shared-memory window)."""
    import ctypes as C
    from php_aho_corasick_b200 import workloads as W
    nbytes = n_hays * HAY_LEN
    rec = {"workload": f"the headline batch ({n_hays} x {HAY_LEN} B, 8 planted needles per haystack) with the prefilter switched off",
           "bytes": nbytes}
    aut.set_filter(-1)
    try:
        scan = lambda: aut.search_device_uniform(resident.data_ptr(), n_hays, HAY_LEN)[1]
        r = timed_device_scans(aut, torch, scan, steps)
        k_ms = r["kernel_ms"] / steps
        rec.update({"kernel_ms": k_ms, "ms_per_step": r["ms_per_step"], "GBps": nbytes / (r["ms_per_step"] * 1e-3) / 1e9,
                    "roofline": _roof(nbytes, k_ms, peak), "events": int(r["events"]),
                    "parity": {"checked": True, "ok": int(r["events"]) == int(n_events),
                               "what": "event count equals the prefilter path's, which is checked per haystack against the CPU "
                                       "reference above (event-level comparison of the two paths: tests/test_gpu_filter.py)"}})
        clean = torch.from_numpy(W.cfg2_stream(7, 0, n_hays // BLOCK_HAYS, planted_per_hay=0)).to(dev)
        for name, mode in (("text_by_tma", 1), ("text_by_ldg", -1)):
            aut.L.acb200_set_tma(C.c_void_p(aut.h), C.c_int(mode))
            r0 = timed_device_scans(aut, torch, lambda: aut.search_device_uniform(clean.data_ptr(), n_hays, HAY_LEN)[1], steps)
            rec.setdefault("no_needles", {})[name] = {"kernel_ms": r0["kernel_ms"] / steps, "events": int(r0["events"]),
                                                        "roofline_frac": _roof(nbytes, r0["kernel_ms"] / steps, peak)["frac"]}
        del clean
    finally:
        aut.L.acb200_set_tma(C.c_void_p(aut.h), C.c_int(0))
        aut.set_filter(0 if filtered else -1)
    return rec


def also_cfg3(torch, dev, peak, hay_bytes=1 << 30, cpu_bytes=64 << 20, steps=5):
    """BASELINE config 3: 100,000 binary signatures of 8..64 B over ONE 1 GiB binary haystack, 1 planted per MiB."""
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.native import Automaton
    pats, hay, off = W.cfg3(hay_bytes=hay_bytes)
    a = Automaton(device=dev.index)
    a.add_php_order(pats)
    t0 = time.time()
    a.finalize()
    fin = time.time() - t0
    inf = a.info()
    text = torch.from_numpy(hay).to(dev)
    scan = lambda: a.search_device(text.data_ptr(), off)[1]
    r = timed_device_scans(a, torch, scan, steps)
    k_ms = r["kernel_ms"] / steps
    rec = {"workload": f"{len(pats)} binary signatures of 8..64 B, one {hay_bytes >> 20} MiB binary haystack, 1 planted signature per MiB",
           "bytes": hay_bytes, "ms_per_step": r["ms_per_step"], "GBps": hay_bytes / (r["ms_per_step"] * 1e-3) / 1e9,
           "kernel_ms": k_ms, "roofline": _roof(hay_bytes, k_ms, peak),
           "per_kernel": kernel_split(r, steps, hay_bytes, peak) if r["filtered"] == steps else None,
           "path": "gram prefilter (W = %d, level-2 bitmap 2^%d bits) + verify" % (inf.filter_word, inf.filter_l2_log2) if r["filtered"] == steps else "full walk",
           "events": int(r["events"]), "automaton": {"states": int(inf.n_states), "table_bytes": int(inf.table_bytes), "finalize_s": round(fin, 2)}}
    # the full walk of the same haystack, for the record
    a.set_filter(-1)
    r2 = timed_device_scans(a, torch, scan, 2, warmup=1)
    rec["full_walk"] = {"kernel_ms": r2["kernel_ms"] / 2, "roofline_frac": _roof(hay_bytes, r2["kernel_ms"] / 2, peak)["frac"],
                        "events": int(r2["events"])}
    a.set_filter(0)
    # CPU: the reference's own code on a prefix, cut into per-thread slices with an (Lmax-1)-byte halo
    kind = cpu_kind()
    threads = min(8, os.cpu_count() or 1)        # every thread builds a private 3.45 M-node trie (~0.6 GB, 16 s): bounded
    cb = cpu_leg(pats, hay[:cpu_bytes], np.array([0, cpu_bytes], dtype=np.uint64), threads, 1, kind,
                 halo=int(inf.max_pattern_len) - 1, what="prefix of the same haystack: ")
    rec["cpu_baseline"] = public(cb)
    # parity of the prefix: the events that end inside it
    ev = packed_to_events(a, torch, (None, scan()), 1, offsets=off)
    rec["parity"] = {"checked": True, "events_in_cpu_prefix_gpu": int(np.count_nonzero(ev["end"] <= cpu_bytes)),
                     "events_in_cpu_prefix_cpu": int(cb["events"]),
                     "ok": int(np.count_nonzero(ev["end"] <= cpu_bytes)) == int(cb["events"]) and r2["events"] == r["events"],
                     "what": "event count of the prefix vs the CPU reference; full-size prefilter path vs full walk event count "
                             "(event-level comparison: tests/test_gpu_filter.py)"}
    a.release()
    return rec


def also_cfg5(torch, dev, peak, hay_bytes=256 << 20, steps=3):
    """BASELINE config 5: a^1..a^4096 submitted (a^1..a^1024 accepted) over 256 MiB of 'a': one event per byte."""
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.native import Automaton
    pats, _, off = W.cfg5(hay_bytes=16)
    off = np.array([0, hay_bytes], dtype=np.uint64)
    a = Automaton(device=dev.index)
    a.add_php_order(pats)
    a.finalize()
    inf = a.info()
    text = torch.full((hay_bytes,), ord("a"), dtype=torch.uint8, device=dev)
    scan = lambda: a.search_device(text.data_ptr(), off)[1]
    r = timed_device_scans(a, torch, scan, steps, warmup=2)
    k_ms = r["kernel_ms"] / steps
    events, hits = W.cfg5_expected(hay_bytes)
    rec = {"workload": f"4096 nested patterns a^1..a^4096 ({inf.n_patterns} accepted) over {hay_bytes >> 20} MiB of 'a': one event per byte",
           "bytes": hay_bytes, "ms_per_step": r["ms_per_step"], "read_GBps": hay_bytes / (r["ms_per_step"] * 1e-3) / 1e9,
           "kernel_ms": k_ms, "events": int(r["events"]),
           "read_plus_16B_per_event_GBps": (hay_bytes + 16 * r["events"]) / (k_ms * 1e-3) / 1e9,
           "read_plus_written_GBps": (hay_bytes + 8 * r["events"]) / (k_ms * 1e-3) / 1e9,
           "roofline": dict(_roof(hay_bytes + 8 * r["events"], k_ms, peak),
                            note="algorithmic bytes = 1 read per haystack byte + 8 written per event (the packed {end, state} "
                                 "record this library emits; SURVEY 8d's 16-byte record figure is read_plus_16B_per_event_GBps)"),
           "path": "full walk (shortest pattern is 1 byte: no prefilter)",
           "parity": {"checked": True, "ok": int(r["events"]) == events,
                      "what": "closed form: one event per byte (event-level check at full size: tests/test_gpu_filter.py)"}}
    # CPU: the C restatement (the reference's own finalize needs 113 s for this dictionary — SURVEY 3.2)
    kind = cpu_kind(prefer_reference=False)
    threads = os.cpu_count() or 1
    cpu_bytes = min(hay_bytes, 64 << 20)
    hay = np.full(cpu_bytes, ord("a"), dtype=np.uint8)
    cb = cpu_leg(pats, hay, np.array([0, cpu_bytes], dtype=np.uint64), threads, 1, kind, halo=int(inf.max_pattern_len) - 1,
                 what="prefix of the same haystack (restatement: the reference's finalize takes 113 s here): ")
    rec["cpu_baseline"] = public(cb)
    rec["parity"]["cpu_events_ok"] = int(cb["events"]) == cpu_bytes
    a.release()
    return rec


def also_literal(torch, dev, aut, needles, calls=256):
    """examples/benchmark.php:55-76 taken literally: 256 SEQUENTIAL ahocorasick_match() calls on 8 KiB haystacks —
    each call is H2D + kernel(s) + D2H + the callback replay, nothing batched."""
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.native import AcText, MATCH_CB, Tally
    hay = W.cfg2_stream(0, 0, 1)
    L = aut.L
    cb = ctypes.cast(L.acb200_tally_match_cb, MATCH_CB)
    texts = []
    for i in range(calls):
        t = AcText()
        t.astring = hay.ctypes.data + i * HAY_LEN
        t.length = HAY_LEN
        texts.append(t)

    def loop():
        tally = Tally()
        for t in texts:
            L.ac_trie_search(aut.h, ctypes.byref(t), 0, cb, ctypes.cast(ctypes.byref(tally), ctypes.c_void_p))
        return tally

    loop()
    best = 1e9
    for _ in range(5):
        t0 = time.time()
        tally = loop()
        best = min(best, time.time() - t0)
    # the same 2 MiB as ONE batched call
    off = W.offsets_uniform(calls, HAY_LEN)
    aut.search_flat_tally(hay.ctypes.data, off)
    tb = 1e9
    for _ in range(5):
        t0 = time.time()
        t_batch = aut.search_flat_tally(hay.ctypes.data, off)
        tb = min(tb, time.time() - t0)
    kind = cpu_kind()
    c1 = cpu_leg(needles, hay, off, 1, 3, kind, digest=True, what="the same 256 calls, one core (the reference is single-threaded): ")
    return {"workload": "256 sequential ac_trie_search() calls, one 8 KiB haystack each (examples/benchmark.php:55-76), pageable host strings",
            "us_per_call": best / calls * 1e6, "GBps": calls * HAY_LEN / best / 1e9, "events": int(tally.events),
            "one_batched_call": {"us": tb * 1e6, "GBps": calls * HAY_LEN / tb / 1e9, "events": int(t_batch.events)},
            "cpu_baseline": dict(public(c1), us_per_call=c1["seconds"] / calls * 1e6),
            "parity": {"checked": True, "ok": int(tally.events) == int(c1["events"]) == int(t_batch.events)}}


def also_cfg4(torch, dist, dev, aut, needles, world, rank, peak, n_hay_total=65536, hay_len=65536, steps=5):
    """BASELINE config 4: 65,536 haystacks x 64 KiB (4 GiB) with the config-2 dictionary, sharded over the N ranks
    (contiguous blocks), events gathered to rank 0.  Device-resident; a seeded 64 MiB block (1,024 haystacks, CPU-checked
    haystack by haystack) rolled by block index fills each shard."""
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.dist import ShardedMatcher
    blk_hays = 1024
    blk = W.cfg2_stream(1000, 0, blk_hays * hay_len // (BLOCK_HAYS * HAY_LEN))        # 64 MiB of config-2 text + needles every 1 KiB
    per_rank = n_hay_total // world
    reps = per_rank // blk_hays
    b = torch.from_numpy(blk).to(dev).view(blk_hays, hay_len)
    shard = torch.cat([torch.roll(b, shifts=(rank * reps + i) % blk_hays, dims=0) for i in range(reps)]).reshape(-1)
    nbytes = per_rank * hay_len
    off = W.offsets_uniform(per_rank, hay_len)
    stream = torch.cuda.current_stream().cuda_stream
    sm = ShardedMatcher(aut)
    parts = -(-nbytes // (2 << 30))              # a device-resident call addresses its stream with 32 bits: 2 GiB per call
    assert world == 1 or parts == 1
    part_hays = per_rank // parts
    k_acc = [0.0]

    def scan():
        if world > 1:
            n = sm.scan_and_gather(shard, off, 0, stream=stream, uniform_len=hay_len)[0]
            k_acc[0] += aut.stats().kernel_ms
            return n
        n = 0
        for p in range(parts):
            n += aut.search_device_uniform(shard.data_ptr() + p * part_hays * hay_len, part_hays, hay_len, stream=stream)[1]
            k_acc[0] += aut.stats().kernel_ms
        return n

    for _ in range(3):
        scan()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k_acc[0] = 0.0
    for _ in range(steps):
        n = scan()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps, k_acc[0] / steps, float(n)], dtype=torch.float64, device=dev)
    tot = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, kms = float(t[0]), float(t[1])
    rec = {"workload": f"{n_hay_total} haystacks x {hay_len} B = {n_hay_total * hay_len >> 30} GiB, config-2 dictionary, "
                       f"{world} rank(s) x {per_rank} haystacks, events gathered to rank 0",
           "bytes": n_hay_total * hay_len, "ms_per_step": ms, "GBps": n_hay_total * hay_len / (ms * 1e-3) / 1e9,
           "kernel_ms_slowest_rank": kms, "roofline": _roof(nbytes, kms, peak), "events": int(tot[2])}
    if rank == 0:
        # parity of the block every shard is made of: per haystack against the CPU reference
        kind = cpu_kind()
        boff = W.offsets_uniform(blk_hays, hay_len)
        cb = cpu_leg(needles, blk, boff, os.cpu_count() or 1, 1, kind, digest=True, what="the 1,024-haystack block: ")
        ev = aut.search_events(blk, boff)
        counts, hashes = aut.event_digest(ev, blk_hays)
        ok = bool(np.array_equal(counts, cb["counts"]) and np.array_equal(hashes, cb["hashes"]))
        rec["cpu_baseline"] = public(cb)
        rec["parity"] = {"checked": True, "ok": ok and int(tot[2]) == int(counts.sum()) * reps * world,
                         "what": "the block's events per haystack vs the CPU reference; all ranks' event total = block total x repetitions"}
    return rec


if __name__ == "__main__":
    sys.exit(main())
