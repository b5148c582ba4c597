"""Per-phase timing of ShardedMatcher.scan_and_gather on N GPUs (torchrun); not the bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
from php_aho_corasick_b200.dist import ShardedMatcher, EventGatherer

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
needles, hay, off = W.cfg2()
a = Automaton(device=lr); a.add_php_order(needles); a.finalize()
reps = 512
res = torch.from_numpy(hay).to(dev).repeat(reps)
offsets = W.offsets_uniform(256 * reps, 8192)
sm = ShardedMatcher(a)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    sm.scan_and_gather(res, offsets, 0, stream=stream, uniform_len=8192)
g = sm._gatherer
def T():
    torch.cuda.synchronize(); return time.perf_counter()
acc = {}
def add(k, dt): acc[k] = acc.get(k, 0.0) + dt
N = 30
for _ in range(N):
    dist.barrier(); t0 = T()
    _, n = a.search_device_uniform(res.data_ptr(), 256 * reps, 8192, stream=stream); t1 = T(); add("search", t1 - t0)
    g._launch_counts(n, world); sizes = g.counts.cpu().tolist(); t2 = T(); add("counts all_gather + .cpu()", t2 - t1)
    a.copy_events(g.send.data_ptr(), n, stream=stream); t3 = T(); add("copy_events", t3 - t2)
    dist.gather(g.send[:g.rows], [g.recv[r, :g.rows] for r in range(world)] if rank == 0 else None, dst=0); t4 = T(); add("gather", t4 - t3)
    t0 = T(); sm.scan_and_gather(res, offsets, 0, stream=stream, uniform_len=8192); t1 = T(); add("scan_and_gather (whole)", t1 - t0)
if rank == 0:
    for k, v in acc.items(): print(f"{k:32s} {v / N * 1e6:8.1f} us")
    print("rows", g.rows, "events", n)
dist.barrier(); dist.destroy_process_group()
