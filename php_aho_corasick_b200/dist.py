"""Batched matching sharded over the GPUs of one box: one process per GPU (torch.distributed).

The path shards by independent haystacks: every rank holds a replica of the
automaton and scans a contiguous block of the batch, balanced by bytes; there is no exchange
while scanning.  The only collective is the gather of the compact event lists to rank 0
(NCCL over NVLink on GPUs; gloo in the CPU tests of this file's host logic).
"""
from __future__ import annotations

import ctypes as C

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


class _DeviceMemory:
    """raw device memory as a __cuda_array_interface__ object (torch.as_tensor maps it without a copy)"""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


class MailboxGatherer:
    """Every rank's event rows -> rank `dst`, with no collective on the critical path.

    all_gather costs every step a kernel that competes with the scan for SMs (the scan's kernels are persistent and
    fill the GPU), an inbound transfer of world x padded rows on EVERY rank, and a host wait that serialises scan and
    exchange: measured 0.80 / 0.74 / 0.63 scaling efficiency at 2 / 4 / 8 B200.  Here rank `dst` owns a buffer that
    the other processes map through CUDA IPC; after its scan a rank copies exactly its own rows into its slot with
    the copy engines (a side stream: NVLink, no SM), then writes {step, count} into its mailbox behind them.  The
    transfer of step k overlaps the scan of step k+1; `dst` looks at step k's mailboxes while it scans step k+1.
    Rows and mailboxes are double-buffered by step parity, an acknowledgement word per parity keeps a fast sender
    from overwriting rows `dst` still holds.  The step itself is ONE library call (acb200_mailbox_step, csrc/cabi.cpp):
    driving the same CUDA calls from here cost 0.1 ms of interpreter time per step.  This class only sets up: buffers,
    IPC handles through the process group, tensor views of the rows.  A rank whose events outgrow `cap_rows` raises:
    size it from a first synchronous step (ShardedMatcher.scan_and_gather, which also remains the general path for
    ragged batches)."""

    MBOX_WORDS = 4          # {step number, event count, densely flagged tiles, -}; 16 bytes apart

    def __init__(self, automaton, cap_rows: int, group=None, dst: int = 0):
        self.aut = automaton
        self.L = automaton.L
        self.group = group
        self.dst = dst
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.cap = int(cap_rows)
        self.dev = torch.cuda.current_device()
        self.step = 0
        self.rows_bytes = self.mbox_bytes = 0
        handles = [None, None, self.cap]          # the collector's capacity is everyone's: slot offsets must agree
        self._own = None
        if self.rank == dst:
            self.rows_bytes = 2 * self.world * self.cap * 8
            self.mbox_bytes = 2 * self.world * self.MBOX_WORDS * 4 + 2 * 4     # mailboxes + one acknowledgement word per parity
            self._own = (self.L.acb200_device_alloc(self.dev, self.rows_bytes), self.L.acb200_device_alloc(self.dev, self.mbox_bytes))
            if not self._own[0] or not self._own[1]:
                raise RuntimeError("acb200_device_alloc failed: " + self.L.acb200_last_error().decode())
            handles = []
            for p in self._own:
                h = C.create_string_buffer(64)
                if self.L.acb200_ipc_export(C.c_void_p(p), h) != 0:
                    raise RuntimeError("acb200_ipc_export failed: " + self.L.acb200_last_error().decode())
                handles.append(h.raw)
            handles.append(self.cap)
        dist.broadcast_object_list(handles, src=dst, group=group)
        self.cap = int(handles[2])
        if self.rank == dst:
            self.rows_ptr, self.mbox_ptr = self._own
        else:
            self.rows_ptr = self.L.acb200_ipc_open(self.dev, handles[0])
            self.mbox_ptr = self.L.acb200_ipc_open(self.dev, handles[1])
            if not self.rows_ptr or not self.mbox_ptr:
                raise RuntimeError("acb200_ipc_open failed: " + self.L.acb200_last_error().decode())
        self.h = self.L.acb200_mailbox_create(self.aut.h, self.rank, self.world, int(self.rank == dst), self.cap,
                                              C.c_void_p(self.rows_ptr), C.c_void_p(self.mbox_ptr))
        if not self.h:
            raise RuntimeError("acb200_mailbox_create failed: " + self.L.acb200_last_error().decode())
        if self.rank == dst:
            d = torch.device("cuda", self.dev)
            self.rows = torch.as_tensor(_DeviceMemory(self.rows_ptr, self.rows_bytes // 4), device=d).view(2, self.world, self.cap, 2)
            self._counts = (C.c_uint32 * self.world)()

    def scan_and_send(self, dev_tensor: torch.Tensor, n_hay: int, hay_len: int, stream=0) -> int:
        """One step: scans this rank's equal-length batch and sends its rows on their way.  -> this rank's event count"""
        n = self.L.acb200_mailbox_step(self.h, C.c_void_p(dev_tensor.data_ptr()), int(n_hay), int(hay_len), C.c_void_p(stream or LEGACY_STREAM))
        if n < 0:
            raise RuntimeError(self.L.acb200_last_error().decode())
        self.step += 1
        return int(n)

    def result(self, k: int):
        """On `dst`: the rows of step k from every rank (views, valid until step k+2 is sent), once they have all
        landed; elsewhere None.  Call it while a later step is in flight — the wait is then already over."""
        if self.rank != self.dst:
            return None
        if self.L.acb200_mailbox_result(self.h, int(k), self._counts) != 0:
            raise RuntimeError(self.L.acb200_last_error().decode())
        p = k & 1
        return [self.rows[p, r, :int(self._counts[r])] for r in range(self.world)]

    def drain(self, stream=0):
        """makes `stream` wait for this rank's copies still in flight (before a clock stops, before buffers go away)"""
        self.L.acb200_mailbox_drain(self.h, C.c_void_p(stream or LEGACY_STREAM))

    def close(self, collective: bool = True):
        """collective: every rank closes now (a barrier keeps the collector's buffers alive until all copies are over)"""
        torch.cuda.synchronize()
        self.L.acb200_mailbox_free(self.h)
        self.h = None
        if collective:
            dist.barrier(group=self.group)
        if self.rank == self.dst:
            self.rows = None
            self.L.acb200_device_free(self.dev, self._own[0])
            self.L.acb200_device_free(self.dev, self._own[1])
        else:
            self.L.acb200_ipc_close(self.dev, self.rows_ptr)
            self.L.acb200_ipc_close(self.dev, self.mbox_ptr)
