"""Batched matching sharded over the GPUs of one box: one process per GPU (torch.distributed).

The path shards by independent haystacks: every rank holds a replica of the
automaton and scans a contiguous block of the batch, balanced by bytes; there is no exchange
while scanning.  The only collective is the gather of the compact event lists to rank 0
(NCCL over NVLink on GPUs; gloo in the CPU tests of this file's host logic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .native import EVENT_DTYPE


def shard_ranges(offsets, world: int):
    """Contiguous haystack blocks [h0, h1) per rank, balanced by bytes. offsets: uint64[n+1]."""
    off = np.asarray(offsets, dtype=np.uint64)
    n = off.size - 1
    total = int(off[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        h = int(np.searchsorted(off, target, side="left"))
        h = min(max(h, cuts[-1]), n)
        cuts.append(h)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_packed_events(local: torch.Tensor, dst: int = 0, group=None):
    """local: int32/uint32-as-int32 tensor [n, 2] = {end offset in this rank's stream, state}.
    Returns on `dst` the list of per-rank tensors (rank order), elsewhere None.
    Variable lengths are handled by a size all_gather followed by a padded all_gather."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(1, max(sizes))
    padded = torch.zeros((cap, 2), dtype=local.dtype, device=local.device)
    if local.shape[0]:
        padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if dist.get_rank(group) != dst:
        return None
    return [bufs[r][: sizes[r]] for r in range(world)]


def globalize(per_rank_events, ranges, offsets):
    """Packed per-rank events -> one structured array (end, state, text_idx) ordered by (text_idx, end)."""
    off = np.asarray(offsets, dtype=np.uint64)
    parts = []
    for ev, (h0, h1) in zip(per_rank_events, ranges):
        a = ev.detach().cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        if a.shape[0] == 0:
            continue
        base = int(off[h0])
        g = a[:, 0].astype(np.uint64) + np.uint64(base)              # end offset in the whole batch stream
        h = np.searchsorted(off, g, side="left") - 1                    # off[h] < g <= off[h+1]
        out = np.empty(a.shape[0], dtype=EVENT_DTYPE)
        out["end"] = g - off[h]
        out["state"] = a[:, 1].astype(np.uint32)
        out["text_idx"] = h.astype(np.uint32)
        parts.append(out)
    if not parts:
        return np.empty(0, dtype=EVENT_DTYPE)
    return np.concatenate(parts)


class ShardedMatcher:
    """ahocorasick_match_batch() over all ranks.  Every rank calls match() with the SAME batch
    description; rank 0 receives the events of the whole batch in (text_idx, end) order."""

    def __init__(self, automaton, group=None):
        self.aut = automaton
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._ev = None

    def scan_local_device(self, dev_tensor: torch.Tensor, local_offsets, first_only=False, stream=0, uniform_len=0):
        """Scans this rank's shard (uint8 CUDA tensor, haystacks end to end). -> int32 CUDA tensor [n,2]"""
        if uniform_len:                          # equal-length batch: no offsets array on any side
            n_hay = len(local_offsets) - 1
            _, n = self.aut.search_device_uniform(dev_tensor.data_ptr(), n_hay, int(uniform_len),
                                                  first_only=first_only, stream=stream)
        else:
            _, n = self.aut.search_device(dev_tensor.data_ptr(), local_offsets, first_only=first_only, stream=stream)
        if self._ev is None or self._ev.shape[0] < max(n, 1):
            self._ev = torch.empty((max(n, 1024), 2), dtype=torch.int32, device=dev_tensor.device)
        self.aut.copy_events(self._ev.data_ptr(), n, stream=stream)
        return self._ev[:n]

    def match(self, flat: np.ndarray, offsets, first_only=False):
        """flat: host uint8 array of the whole batch; offsets uint64[n+1]."""
        off = np.asarray(offsets, dtype=np.uint64)
        ranges = shard_ranges(off, self.world)
        h0, h1 = ranges[self.rank]
        lo, hi = int(off[h0]), int(off[h1])
        local_off = off[h0:h1 + 1] - off[h0]
        dev = torch.device("cuda", torch.cuda.current_device())
        shard = torch.from_numpy(np.ascontiguousarray(flat[lo:hi])).to(dev) if hi > lo else torch.empty(0, dtype=torch.uint8, device=dev)
        ev = self.scan_local_device(shard, local_off, first_only=first_only,
                                    stream=torch.cuda.current_stream().cuda_stream)
        if first_only and ev.shape[0]:
            ev = _first_per_haystack(ev, local_off)
        if self.world == 1:
            return globalize([ev], ranges, off)
        got = gather_packed_events(ev.contiguous(), 0, self.group)
        if got is None:
            return None
        return globalize(got, ranges, off)


def _first_per_haystack(ev: torch.Tensor, local_off) -> torch.Tensor:
    """Device-side events of a first_only scan may hold one candidate per slice; keep the earliest per haystack."""
    a = ev.detach().cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    off = np.asarray(local_off, dtype=np.uint64)
    h = np.searchsorted(off, a[:, 0].astype(np.uint64), side="left") - 1
    keep = np.ones(a.shape[0], dtype=bool)
    keep[1:] = h[1:] != h[:-1]
    return ev[torch.from_numpy(keep).to(ev.device)]
