"""Host logic of the multi-GPU path on CPU: byte-balanced sharding, the variable-length event gather
(gloo, world_size 2) and the conversion back to per-haystack coordinates."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from php_aho_corasick_b200.dist import EventGatherer, gather_packed_events, globalize, shard_ranges


def test_shard_ranges_cover_and_balance():
    off = np.concatenate([[0], np.cumsum([10, 0, 5, 100, 1, 1, 50, 0, 33])]).astype(np.uint64)
    for world in (1, 2, 3, 4, 8, 16):
        r = shard_ranges(off, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == 9
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    off = np.arange(0, 65537, dtype=np.uint64) * 65536          # config 4 shape: 65536 x 64 KiB
    r = shard_ranges(off, 8)
    assert [b - a for a, b in r] == [8192] * 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # batch of 6 haystacks; rank 0 owns [0,3), rank 1 owns [3,6)
    off = np.array([0, 100, 100, 250, 300, 420, 500], dtype=np.uint64)
    ranges = shard_ranges(off, world)
    if rank == 0:       # events (end offset in own stream, state)
        local = torch.tensor([[5, 7], [100, 9], [101, 3], [250, 4]], dtype=torch.int32)
    else:
        base = int(off[ranges[1][0]])
        local = torch.tensor([[300 - base, 11], [301 - base, 12], [500 - base, 13]], dtype=torch.int32) if ranges[1][1] > ranges[1][0] else torch.zeros((0, 2), dtype=torch.int32)
    got = gather_packed_events(local, 0)
    if rank == 0:
        ev = globalize(got, ranges, off)
        q.put((ranges, ev["text_idx"].tolist(), ev["end"].tolist(), ev["state"].tolist()))
    else:
        assert got is None
    # second round with an empty contribution from rank 1
    got = gather_packed_events(local if rank == 0 else torch.zeros((0, 2), dtype=torch.int32), 0)
    if rank == 0:
        q.put(sum(int(g.shape[0]) for g in got))
    # third round: rank 1 outgrows the agreed capacity -> every rank regrows and the gather is repeated
    big = torch.arange(2 * 5000, dtype=torch.int32).reshape(5000, 2)
    got = gather_packed_events(local if rank == 0 else big, 0)
    if rank == 0:
        q.put((int(got[0].shape[0]), int(got[1].shape[0]), got[1][-1].tolist(), got[0].tolist() == local.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_rank_gather_and_globalize():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranges, tidx, end, state = q.get(timeout=120)
    n_second = q.get(timeout=120)
    third = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ranges == [(0, 3), (3, 6)]
    # haystack 0 = (0,100], 2 = (100,250], 3 = (250,300], 4 = (300,420], 5 = (420,500]
    assert tidx == [0, 0, 2, 2, 3, 4, 5]
    assert end == [5, 100, 1, 150, 50, 1, 80]
    assert state == [7, 9, 3, 4, 11, 12, 13]
    assert n_second == 4
    assert third == (4, 5000, [9998, 9999], True)


def _chained_worker(rank, world, port, q):
    """The chained scan -> gather step (ShardedMatcher.scan_and_gather, equal-length batches): the library writes
    {count, events} straight into the send buffer and EventGatherer.exchange() sends it on.  Here a stand-in writes
    the rows; rank 1 takes the synchronous gather() in the second step (a rank whose batch needs the full walk),
    and outgrows the agreed rows in the third — all ranks must stay in step."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = EventGatherer()
    like = torch.zeros((0, 2), dtype=torch.int32)

    def library_writes(events, rows):                    # what acb200_search_device_uniform_async does with d_rows
        g._ensure(rows, world, like)
        g.send[0, 0] = events.shape[0]
        m = min(events.shape[0], rows)
        g.send[1:1 + m] = events[:m]

    def chained(events):
        while True:
            rows = g.rows
            library_writes(events, rows)
            sizes = g.exchange(rows, world)
            if max(sizes) <= rows:
                return sizes, g.views(rows, world, sizes, 0)

    mine = torch.arange(2 * (10 + rank), dtype=torch.int32).reshape(-1, 2) + 1000 * rank
    got = g.gather(mine, 0)                              # first step: the synchronous form agrees on the rows
    assert g.rows >= 1023
    out = []
    sizes, got = chained(mine)                           # second step: every rank chained
    out.append((sizes, None if got is None else [x.tolist() for x in got]))
    if rank == 0:                                        # third step: rank 0 chained, rank 1 synchronous
        sizes, got = chained(mine)
        out.append((sizes, [x.tolist() for x in got]))
    else:
        assert g.gather(mine, 0) is None
    big = torch.arange(2 * 3000, dtype=torch.int32).reshape(3000, 2)
    sizes, got = chained(big if rank == 1 else mine)     # fourth step: rank 1 outgrows the rows -> both repeat
    out.append((sizes, None if got is None else (int(got[0].shape[0]), int(got[1].shape[0]), got[1][-1].tolist())))
    q.put((rank, out, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_chained_scan_and_gather_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chained_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        rank, out, mine = q.get(timeout=120)
        res[rank] = (out, mine)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    out0, mine0 = res[0]
    out1, mine1 = res[1]
    assert out0[0] == ([10, 11], [mine0, mine1]) and out1[0] == ([10, 11], None)
    assert out0[1] == ([10, 11], [mine0, mine1])
    assert out0[2] == ([10, 3000], (10, 3000, [5998, 5999])) and out1[1] == ([10, 3000], None)


def _nccl_worker(rank, world, port, q, same_gpu=False):
    """Every rank scans its shard of ONE seeded batch (chained scan -> all_gather), rank 0 digests the gathered rows per
    haystack.  Two real GPUs over NCCL — or, on a one-GPU box, two processes on GPU 0 over gloo (NCCL refuses two
    ranks on one device): the library calls, buffers and stream ordering are the same."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if same_gpu else rank
    torch.cuda.set_device(dev)
    if same_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.dist import ShardedMatcher
    from php_aho_corasick_b200.native import Automaton
    needles, _ = W.cfg2_needles()
    hay_len, blocks = 8192, 8                                # 16 MiB per rank: the prefilter path (>= 8 MiB)
    flat = W.cfg2_stream(0, 0, blocks * world)               # the whole batch, identical on every rank
    n_hays = blocks * 256
    off = W.offsets_uniform(n_hays, hay_len)
    a = Automaton(device=dev)
    a.add_php_order(needles)
    a.finalize()
    shard = torch.from_numpy(flat[rank * n_hays * hay_len:(rank + 1) * n_hays * hay_len]).cuda()
    sm = ShardedMatcher(a)
    res = []
    goff = W.offsets_uniform(world * n_hays, hay_len)
    for step in range(3):                                    # step 0 synchronous, steps 1.. the chained form
        n, got = sm.scan_and_gather(shard, off, 0, stream=torch.cuda.current_stream().cuda_stream, uniform_len=hay_len)
        if rank == 0:
            ev = globalize(got, [(r * n_hays, (r + 1) * n_hays) for r in range(world)], goff)
            counts, hashes = a.event_digest(ev, world * n_hays)
            ordered = bool(np.all(np.diff(ev["text_idx"].astype(np.int64)) >= 0))
            if not ordered:                                  # what a race between a rank's scan / copy and the collective looks like
                for r, g_r in enumerate(got):
                    e = g_r.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
                    down = np.flatnonzero(np.diff(e[:, 0]) <= 0)
                    print(f"step {step}: rank {r} sent {len(e)} rows, {len(down)} not ascending, first at {down[:3].tolist()}, "
                          f"zero rows {int(np.sum(e[:, 0] == 0))}", flush=True)
            res.append((counts.tolist(), hashes.tolist(), ordered))
    # the same steps through the mailbox gather (copy engines into rank 0's IPC-mapped buffer, pipelined by one step):
    # several steps in a row, so that slots are reused and the acknowledgement flow control is exercised
    from php_aho_corasick_b200.dist import MailboxGatherer
    mg = MailboxGatherer(a, cap_rows=2 * n + 1024)
    for step in range(5):
        mg.scan_and_send(shard, n_hays, hay_len, stream=torch.cuda.current_stream().cuda_stream)
        if step >= 1 and rank == 0:
            got = [g.clone() for g in mg.result(step - 1)]
            ev = globalize(got, [(r * n_hays, (r + 1) * n_hays) for r in range(world)], goff)
            counts, hashes = a.event_digest(ev, world * n_hays)
            res.append((counts.tolist(), hashes.tolist(), bool(np.all(np.diff(ev["text_idx"].astype(np.int64)) >= 0))))
    mg.drain()
    if rank == 0:
        ev = globalize(mg.result(4), [(r * n_hays, (r + 1) * n_hays) for r in range(world)], goff)
        counts, hashes = a.event_digest(ev, world * n_hays)
        res.append((counts.tolist(), hashes.tolist(), True))
    mg.close()
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_gathered_rows_equal_the_cpu_reference_per_haystack():
    from oracle import pydriver
    from php_aho_corasick_b200 import workloads as W
    same_gpu = torch.cuda.device_count() < 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, same_gpu)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    needles, _ = W.cfg2_needles()
    flat = W.cfg2_stream(0, 0, 16)
    off = W.offsets_uniform(16 * 256, 8192)
    kind = "reference" if pydriver.available("reference") else "oracle"
    _, _, counts, hashes = pydriver.bench_digest(kind, needles, flat, off, 4)
    assert len(res) == 8                                     # 3 steps through all_gather, 5 results through the mailboxes
    bad = [(i, ordered, int(np.sum(np.array(c, dtype=np.uint64) != counts)), int(np.sum(np.array(h, dtype=np.uint64) != hashes)), sum(c), int(counts.sum()))
           for i, (c, h, ordered) in enumerate(res)
           if not ordered or c != counts.tolist() or h != hashes.tolist()]
    assert not bad, bad


def _mailbox_worker(rank, world, port, q, same_gpu):
    """MailboxGatherer between separate PROCESSES: rank 0's buffer mapped through CUDA IPC, rows copied in by the copy
    engines, mailboxes, acknowledgement flow control.  With `same_gpu` both ranks sit on GPU 0 (IPC works between
    processes on one device; the process group is gloo because NCCL refuses two ranks on one GPU) — what a one-GPU
    box can run; else every rank takes its own GPU."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if same_gpu else rank
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from php_aho_corasick_b200 import workloads as W
    from php_aho_corasick_b200.dist import MailboxGatherer
    from php_aho_corasick_b200.native import Automaton
    needles, _ = W.cfg2_needles()
    hay_len, blocks = 8192, 8
    flat = W.cfg2_stream(0, 0, blocks * world)
    n_hays = blocks * 256
    goff = W.offsets_uniform(world * n_hays, hay_len)
    ranges = [(r * n_hays, (r + 1) * n_hays) for r in range(world)]
    a = Automaton(device=dev)
    a.add_php_order(needles)
    a.finalize()
    shard = torch.from_numpy(flat[rank * n_hays * hay_len:(rank + 1) * n_hays * hay_len]).cuda()
    mg = MailboxGatherer(a, cap_rows=3 * n_hays * 8)
    res = []
    steps = 6
    for step in range(steps):
        if rank == 1 and step == 3:
            import time
            time.sleep(0.3)                                  # a slow sender: rank 0 must wait for its mailbox
        if rank == 0 and step == 4:
            import time
            time.sleep(0.3)                                  # a slow collector: the sender must wait for the acknowledgement
        mg.scan_and_send(shard, n_hays, hay_len, stream=torch.cuda.current_stream().cuda_stream)
        if step >= 1 and rank == 0:
            got = [g.clone() for g in mg.result(step - 1)]
            ev = globalize(got, ranges, goff)
            counts, hashes = a.event_digest(ev, world * n_hays)
            res.append((counts.tolist(), hashes.tolist()))
    mg.drain()
    if rank == 0:
        ev = globalize(mg.result(steps - 1), ranges, goff)
        counts, hashes = a.event_digest(ev, world * n_hays)
        res.append((counts.tolist(), hashes.tolist()))
        q.put(res)
    mg.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_mailbox_gather_between_processes_equals_the_cpu_reference_per_haystack():
    from oracle import pydriver
    from php_aho_corasick_b200 import workloads as W
    same_gpu = torch.cuda.device_count() < 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mailbox_worker, args=(r, 2, port, q, same_gpu)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    needles, _ = W.cfg2_needles()
    flat = W.cfg2_stream(0, 0, 16)
    off = W.offsets_uniform(16 * 256, 8192)
    kind = "reference" if pydriver.available("reference") else "oracle"
    _, _, counts, hashes = pydriver.bench_digest(kind, needles, flat, off, 4)
    assert len(res) == 6
    for c, h in res:
        assert c == counts.tolist() and h == hashes.tolist()
