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

LEGACY_STREAM = 1        # cudaStreamLegacy: the explicit handle of the default stream torch uses (its cuda_stream is 0)


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


class EventGatherer:
    """Variable-length gather of packed event lists with persistent buffers: ONE collective and one host wait
    per call.

    Every rank contributes `rows + 1` rows: row 0 carries its event count, rows 1.. its events; an all_gather
    brings all of it to every rank (a collective costs ~0.1 ms of launch latency here, the extra NVLink traffic
    of all_gather over gather far less).  `rows` is the same on all ranks by construction — a deterministic
    function of the previous call's counts, which every rank saw, and a little above them; if this call's
    counts outgrew it, every rank sees that too and the gather is repeated with more rows.  The tensors
    returned on `dst` are views into the receive buffer, valid until the next call."""

    def __init__(self, group=None):
        self.group = group
        self.rows = 0            # event rows per rank in the collective (agreed)
        self.send = None
        self.recv = None
        self.row0_extra = []

    def _ensure(self, rows, world, like):
        if self.send is None or self.send.shape[0] < rows + 1 or self.send.device != like.device or self.send.dtype != like.dtype:
            cap = max(rows + rows // 4, 1024) + 1
            self.send = torch.zeros((cap, 2), dtype=like.dtype, device=like.device)
            self.recv = torch.zeros((world * cap, 2), dtype=like.dtype, device=like.device)

    @staticmethod
    def _rows_for(max_count):
        return max(1023, (max_count + max_count // 8 + 4095) // 4096 * 4096 - 1)

    def exchange(self, rows: int, world: int):
        """all_gather of the first rows + 1 rows of the send buffer (row 0 = this rank's count, already in place);
        -> every rank's count (one host wait); agrees on the rows of the next call."""
        out = self.recv[: world * (rows + 1)]
        try:
            dist.all_gather_into_tensor(out, self.send[: rows + 1], group=self.group)
        except (RuntimeError, NotImplementedError, AttributeError):
            dist.all_gather([out[r * (rows + 1):(r + 1) * (rows + 1)] for r in range(world)], self.send[: rows + 1],
                            group=self.group)
        row0 = out.view(world, rows + 1, 2)[:, 0, :].cpu().tolist()      # the one host wait of the step
        sizes = [int(r[0]) & 0xFFFFFFFF for r in row0]
        self.row0_extra = [int(r[1]) & 0xFFFFFFFF for r in row0]           # second word of row 0 (the library: dense tiles)
        self.rows = self._rows_for(max(sizes))
        return sizes

    def views(self, rows: int, world: int, sizes, dst: int):
        if dist.get_rank(self.group) != dst:
            return None
        got = self.recv[: world * (rows + 1)].view(world, rows + 1, 2)
        return [got[r, 1:1 + sizes[r]] for r in range(world)]

    def gather(self, local, dst: int = 0, n: int | None = None, fill=None, like: torch.Tensor | None = None):
        """local: [n, 2] tensor of this rank's events — or, to save a device copy, `n` plus `fill(rows)`, a callable
        that writes the first len(rows) events into the given rows of the send buffer (`like` gives device/dtype)."""
        world = dist.get_world_size(self.group)
        if local is not None:
            n = int(local.shape[0])
            like = local
        if self.rows == 0:                       # first call: agree on a size from the counts alone
            n_local = torch.tensor([n], dtype=torch.int64, device=like.device)
            sizes = [torch.zeros_like(n_local) for _ in range(world)]
            dist.all_gather(sizes, n_local, group=self.group)
            self.rows = self._rows_for(max(int(x.item()) for x in sizes))
        while True:
            rows = self.rows
            self._ensure(rows, world, like)
            self.send[0].fill_(n)                # row 0 = the count (a fill kernel, no host-to-device copy)
            m = min(n, rows)
            if m:
                if fill is not None:
                    fill(self.send[1:1 + m])
                else:
                    self.send[1:1 + m].copy_(local[:m])
            out = self.recv[: world * (rows + 1)]
            try:
                dist.all_gather_into_tensor(out, self.send[: rows + 1], group=self.group)
            except (RuntimeError, NotImplementedError, AttributeError):
                dist.all_gather([out[r * (rows + 1):(r + 1) * (rows + 1)] for r in range(world)], self.send[: rows + 1],
                                group=self.group)
            got = out.view(world, rows + 1, 2)
            sizes = [int(x) for x in got[:, 0, 0].cpu().tolist()]     # the one host wait of the call
            self.rows = self._rows_for(max(sizes))                    # same decision on every rank
            if max(sizes) <= rows:
                break
        if dist.get_rank(self.group) != dst:
            return None
        return [got[r, 1:1 + sizes[r]] for r in range(world)]


_gatherers = {}


def gather_packed_events(local: torch.Tensor, dst: int = 0, group=None):
    """local: int32/uint32-as-int32 tensor [n, 2] = {end offset in this rank's stream, state}.
    Returns on `dst` the list of per-rank tensors (rank order; views valid until the next call), elsewhere None."""
    g = _gatherers.get(id(group))
    if g is None:
        g = _gatherers[id(group)] = EventGatherer(group)
    return g.gather(local, dst)


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
        self._gatherer = None
        self._like = None

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
        # (a NULL handle would mean the library's private stream: the copy has to be ordered with the caller's stream)
        self.aut.copy_events(self._ev.data_ptr(), n, stream=stream or LEGACY_STREAM)
        return self._ev[:n]

    def scan_and_gather(self, dev_tensor: torch.Tensor, local_offsets, dst: int = 0, stream=0, uniform_len=0):
        """Scans this rank's shard and gathers every rank's packed events on `dst` (list in rank order, views valid
        until the next call; None elsewhere).  The events go from the library's buffer straight into the send buffer.

        Equal-length batches after the first call take the chained form: the library enqueues its kernels with the
        send buffer as their output (row 0 = the count) and returns at once, the all_gather is enqueued behind them,
        and the one host wait of the step is the read of the gathered counts — no host round trip between scan and
        collective.  If a rank's events outgrow the agreed rows, every rank sees it in the counts and repeats."""
        if self._gatherer is None:
            self._gatherer = EventGatherer(self.group)
        g = self._gatherer
        if uniform_len and g.rows and dev_tensor.is_cuda:
            n_hay = len(local_offsets) - 1
            while True:
                rows = g.rows
                if self._like is None:
                    self._like = torch.empty((0, 2), dtype=torch.int32, device=dev_tensor.device)
                g._ensure(rows, self.world, self._like)
                if not self.aut.search_device_uniform_async(dev_tensor.data_ptr(), n_hay, int(uniform_len),
                                                            g.send.data_ptr(), rows, stream=stream):
                    break                                    # this batch needs the synchronous call (full walk)
                sizes = g.exchange(rows, self.world)         # all_gather + the host wait
                n = sizes[self.rank]
                self.aut.async_finish(n, g.row0_extra[self.rank])
                if max(sizes) <= rows:
                    return n, g.views(rows, self.world, sizes, dst)
        if uniform_len:
            _, n = self.aut.search_device_uniform(dev_tensor.data_ptr(), len(local_offsets) - 1, int(uniform_len), stream=stream)
        else:
            _, n = self.aut.search_device(dev_tensor.data_ptr(), local_offsets, stream=stream)
        like = torch.empty((0, 2), dtype=torch.int32, device=dev_tensor.device)
        got = self._gatherer.gather(None, dst, n=n, like=like,
                                    fill=lambda rows: self.aut.copy_events(rows.data_ptr(), rows.shape[0], stream=stream or LEGACY_STREAM))
        return n, got

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
