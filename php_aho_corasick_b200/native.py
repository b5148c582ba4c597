"""ctypes binding of libacb200.so (the C-ABI in include/acb200.h).

There is no fallback: if the CUDA library is missing or no B200 is visible the
calls raise.  Nothing here imports or calls anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACB200_LIB") or os.path.join(_HERE, "libacb200.so")
_lib = None


class AcText(C.Structure):
    _fields_ = [("astring", C.c_void_p), ("length", C.c_size_t)]


class _PattIdU(C.Union):
    _fields_ = [("stringy", C.c_char_p), ("number", C.c_long)]


class AcPattId(C.Structure):
    _fields_ = [("u", _PattIdU), ("type", C.c_int)]


class AcPattern(C.Structure):
    _fields_ = [("ptext", AcText), ("rtext", AcText), ("id", AcPattId), ("aux", C.c_void_p)]


class AcMatch(C.Structure):
    _fields_ = [("patterns", C.POINTER(AcPattern)), ("size", C.c_size_t), ("position", C.c_size_t)]


class Event(C.Structure):
    _fields_ = [("end", C.c_uint64), ("state", C.c_uint32), ("text_idx", C.c_uint32)]


HIT_DTYPE = np.dtype([("text_idx", np.uint32), ("end", np.uint32), ("start", np.uint32), ("pattern", np.uint32)])
EVENT_DTYPE = np.dtype([("end", np.uint64), ("state", np.uint32), ("text_idx", np.uint32)])
PACKED_EVENT_DTYPE = np.dtype([("end", np.uint32), ("state", np.uint32)])


class Slab(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("halo", C.c_uint32), ("device_slot", C.c_uint32),
                ("first_text", C.c_uint64), ("end_text", C.c_uint64)]


class Tally(C.Structure):
    _fields_ = [("events", C.c_uint64), ("hits", C.c_uint64), ("hash", C.c_uint64)]


class Info(C.Structure):
    _fields_ = [("n_patterns", C.c_uint64), ("n_states", C.c_uint64), ("n_classes", C.c_uint32),
                ("entry_bytes", C.c_uint32), ("max_pattern_len", C.c_uint32), ("final_bound", C.c_uint32),
                ("root", C.c_uint32), ("table_bytes", C.c_uint64), ("device", C.c_int32),
                ("finalized", C.c_int32), ("filter_word", C.c_int32), ("min_pattern_len", C.c_uint32),
                ("filter_l1_fill", C.c_float), ("filter_l2_log2", C.c_uint32), ("reserved_", C.c_uint32),
                ("direct_keys", C.c_uint32), ("direct_walk_keys", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("events", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("chunk_bytes", C.c_uint32), ("halo_bytes", C.c_uint32), ("kernel_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("devices", C.c_uint32), ("filtered", C.c_uint32),
                ("filter_ms", C.c_float), ("verify_ms", C.c_float), ("flagged_words", C.c_uint64),
                ("dense_tiles", C.c_uint64), ("reorder_ms", C.c_float), ("expand_ms", C.c_float)]


MATCH_CB = C.CFUNCTYPE(C.c_int, C.POINTER(AcMatch), C.c_void_p)
BATCH_CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.POINTER(AcMatch), C.c_void_p)

# every symbol include/acb200.h declares
EXPORTS = [
    "ac_trie_create", "ac_trie_add", "ac_trie_finalize", "ac_trie_search", "ac_trie_release",
    "ac_trie_search_batch", "ac_trie_search_flat", "acb200_search_events", "acb200_search_device",
    "acb200_state_patterns", "acb200_info", "acb200_last_stats", "acb200_last_error",
    "acb200_set_device", "acb200_device_count", "acb200_host_alloc", "acb200_host_free",
    "acb200_set_tuning", "acb200_version", "acb200_copy_events", "acb200_tally_cb", "acb200_tally_match_cb",
    "acb200_set_filter", "acb200_search_device_uniform", "acb200_search_hits", "acb200_pattern", "acb200_save", "acb200_load", "acb200_filter_probe",
    "acb200_set_direct", "acb200_set_tma", "acb200_direct_probe", "acb200_search_device_uniform_async", "acb200_async_finish",
    "acb200_set_devices", "acb200_set_slab_bytes", "acb200_plan_slabs", "acb200_event_digest",
    "acb200_device_alloc", "acb200_device_free", "acb200_ipc_export", "acb200_ipc_open", "acb200_ipc_close",
    "acb200_copy_async", "acb200_mailbox_wait_async", "acb200_mailbox_create", "acb200_mailbox_step", "acb200_mailbox_result",
    "acb200_mailbox_drain", "acb200_mailbox_free", "acb200_last_hits",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.ac_trie_create.restype = C.c_void_p
    L.ac_trie_add.argtypes = [C.c_void_p, C.POINTER(AcPattern), C.c_int]
    L.ac_trie_add.restype = C.c_int
    L.ac_trie_finalize.argtypes = [C.c_void_p]
    L.ac_trie_finalize.restype = None
    L.ac_trie_search.argtypes = [C.c_void_p, C.POINTER(AcText), C.c_int, MATCH_CB, C.c_void_p]
    L.ac_trie_search.restype = C.c_int
    L.ac_trie_release.argtypes = [C.c_void_p]
    L.ac_trie_release.restype = None
    L.ac_trie_search_batch.argtypes = [C.c_void_p, C.POINTER(AcText), C.c_size_t, C.c_int, BATCH_CB, C.c_void_p]
    L.ac_trie_search_batch.restype = C.c_int
    L.ac_trie_search_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, BATCH_CB, C.c_void_p]
    L.ac_trie_search_flat.restype = C.c_int
    L.acb200_search_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.acb200_search_events.restype = C.c_int
    L.acb200_search_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.acb200_search_device.restype = C.c_int
    L.acb200_search_device_uniform.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.acb200_search_device_uniform.restype = C.c_int
    L.acb200_state_patterns.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(AcPattern))]
    L.acb200_state_patterns.restype = C.c_size_t
    L.acb200_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.acb200_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.acb200_last_error.restype = C.c_char_p
    L.acb200_set_device.argtypes = [C.c_int]
    L.acb200_device_count.restype = C.c_int
    L.acb200_host_alloc.argtypes = [C.c_size_t]
    L.acb200_host_alloc.restype = C.c_void_p
    L.acb200_host_free.argtypes = [C.c_void_p]
    L.acb200_host_free.restype = None
    L.acb200_set_tuning.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.acb200_version.restype = C.c_char_p
    L.acb200_set_filter.argtypes = [C.c_void_p, C.c_int]
    L.acb200_search_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.acb200_search_hits.restype = C.c_int
    L.acb200_pattern.argtypes = [C.c_void_p, C.c_size_t]
    L.acb200_pattern.restype = C.POINTER(AcPattern)
    L.acb200_filter_probe.argtypes = [C.c_void_p, C.c_uint64, C.c_uint]
    L.acb200_filter_probe.restype = C.c_int
    L.acb200_search_device_uniform_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.acb200_search_device_uniform_async.restype = C.c_int
    L.acb200_async_finish.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    L.acb200_async_finish.restype = C.c_int
    L.acb200_set_direct.argtypes = [C.c_void_p, C.c_int]
    L.acb200_set_tma.argtypes = [C.c_void_p, C.c_int]
    L.acb200_direct_probe.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.acb200_direct_probe.restype = C.c_int
    L.acb200_save.argtypes = [C.c_void_p, C.c_char_p]
    L.acb200_save.restype = C.c_int
    L.acb200_load.argtypes = [C.c_char_p]
    L.acb200_load.restype = C.c_void_p
    L.acb200_set_devices.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_size_t]
    L.acb200_set_devices.restype = C.c_int
    L.acb200_set_slab_bytes.argtypes = [C.c_void_p, C.c_uint64]
    L.acb200_set_slab_bytes.restype = C.c_int
    L.acb200_plan_slabs.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint64, C.POINTER(Slab), C.c_size_t,
                                    C.POINTER(C.c_size_t)]
    L.acb200_plan_slabs.restype = C.c_int
    L.acb200_event_digest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
    L.acb200_event_digest.restype = C.c_int
    L.acb200_device_alloc.argtypes = [C.c_int, C.c_size_t]
    L.acb200_device_alloc.restype = C.c_void_p
    L.acb200_device_free.argtypes = [C.c_int, C.c_void_p]
    L.acb200_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
    L.acb200_ipc_open.argtypes = [C.c_int, C.c_char_p]
    L.acb200_ipc_open.restype = C.c_void_p
    L.acb200_ipc_close.argtypes = [C.c_int, C.c_void_p]
    L.acb200_copy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.acb200_mailbox_wait_async.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.acb200_mailbox_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    L.acb200_mailbox_create.restype = C.c_void_p
    L.acb200_mailbox_step.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.acb200_mailbox_step.restype = C.c_long
    L.acb200_mailbox_result.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.acb200_mailbox_drain.argtypes = [C.c_void_p, C.c_void_p]
    L.acb200_mailbox_free.argtypes = [C.c_void_p]
    L.acb200_mailbox_free.restype = None
    L.acb200_copy_events.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.acb200_copy_events.restype = C.c_long
    _lib = L
    return L


class AcError(RuntimeError):
    pass


def last_error() -> str:
    return lib().acb200_last_error().decode("utf-8", "replace")


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(buf), dtype=np.uint8)
    return np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1)


def plan_slabs(offsets, halo_max: int, n_devices: int, slab_bytes: int = 0):
    """the slab plan of a host call (csrc/shard.hpp), evaluated on the host -> list of dicts"""
    L = lib()
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = C.c_size_t(0)
    L.acb200_plan_slabs(off.ctypes.data, off.size - 1, int(halo_max), int(n_devices), int(slab_bytes), None, 0, C.byref(n))
    arr = (Slab * max(1, n.value))()
    L.acb200_plan_slabs(off.ctypes.data, off.size - 1, int(halo_max), int(n_devices), int(slab_bytes), arr, n.value, C.byref(n))
    return [{"begin": int(x.begin), "end": int(x.end), "halo": int(x.halo), "device_slot": int(x.device_slot),
             "first_text": int(x.first_text), "end_text": int(x.end_text)} for x in arr[: n.value]]


class Automaton:
    """Thin object over an AC_TRIE_t*.  Pattern ordinals are the order of add()."""

    def __init__(self, device: int | None = None):
        self.L = lib()
        if device is not None:
            self.L.acb200_set_device(int(device))
        self.h = self.L.ac_trie_create()
        self.n_added = 0
        self._keep = []   # keeps pattern bytes alive for copy=0 callers (unused: we copy)

    # -- build ------------------------------------------------------------
    def _add(self, pattern: bytes, ordinal: int) -> int:
        p = AcPattern()
        buf = C.create_string_buffer(pattern, len(pattern))
        p.ptext.astring = C.cast(buf, C.c_void_p)
        p.ptext.length = len(pattern)
        p.rtext.astring = None
        p.rtext.length = 0
        p.id.type = 1
        p.id.u.number = ordinal
        p.aux = ordinal + 1
        return self.L.ac_trie_add(self.h, C.byref(p), 1)

    def add(self, pattern: bytes) -> int:
        o = self.n_added
        self.n_added += 1
        return self._add(pattern, o)

    def add_php_order(self, patterns) -> None:
        """One init()/add_patterns() call: ordinals follow array order, insertion is last-first
        (reference src/php_ahocorasick.c:410-421, 457-486); statuses ignored."""
        patterns = list(patterns)
        base = self.n_added
        for k in range(len(patterns) - 1, -1, -1):
            self._add(patterns[k], base + k)
        self.n_added += len(patterns)

    def finalize(self) -> None:
        self.L.ac_trie_finalize(self.h)
        inf = self.info()
        if inf.device < 0:
            raise AcError("finalize did not reach the GPU: " + last_error())

    def set_tuning(self, chunk_bytes: int = 0, smem_table_bytes: int = 0) -> None:
        self.L.acb200_set_tuning(self.h, int(chunk_bytes), int(smem_table_bytes))

    def save(self, path: str) -> None:
        if self.L.acb200_save(self.h, os.fsencode(path)) != 0:
            raise AcError(last_error())

    @classmethod
    def load(cls, path: str, device: int | None = None, require_device: bool = True) -> "Automaton":
        self = cls.__new__(cls)
        self.L = lib()
        if device is not None:
            self.L.acb200_set_device(int(device))
        self.h = self.L.acb200_load(os.fsencode(path))
        self.n_added = 0
        self._keep = []
        if not self.h:
            raise AcError(last_error())
        if require_device and self.info().device < 0:
            raise AcError("load did not reach the GPU: " + last_error())
        return self

    def filter_probe(self, word: int, next_byte: int) -> int:
        """host-side evaluation of the prefilter decision for one aligned word (diagnostic, see acb200.h)"""
        return int(self.L.acb200_filter_probe(self.h, int(word), int(next_byte)))

    def set_direct(self, mode: int) -> None:
        """0 automatic (= 1), 1 flagged words settled by one comparison inside the walk kernel, -1 every flagged word is walked"""
        self.L.acb200_set_direct(self.h, int(mode))

    def set_tma(self, mode: int) -> None:
        """full walk: 0 automatic, 1 haystack text staged by the TMA unit wherever the shape allows, -1 plain loads"""
        self.L.acb200_set_tma(self.h, int(mode))

    def direct_probe(self, text: bytes, word_index: int, hay_begin: int = 0):
        """host-side evaluation of the direct verification of one aligned word -> (verdict, end, state)"""
        e, s = C.c_uint32(0), C.c_uint32(0)
        v = int(self.L.acb200_direct_probe(self.h, text, len(text), int(hay_begin), int(word_index), C.byref(e), C.byref(s)))
        return v, int(e.value), int(s.value)

    def set_devices(self, devices) -> None:
        """GPUs one host call may use (the automaton is replicated onto them; see acb200.h)"""
        arr = (C.c_int * max(1, len(devices)))(*[int(d) for d in devices])
        if self.L.acb200_set_devices(self.h, arr, len(devices)) != 0:
            raise AcError(last_error())

    def set_slab_bytes(self, nbytes: int) -> None:
        if self.L.acb200_set_slab_bytes(self.h, int(nbytes)) != 0:
            raise AcError(last_error())

    def set_filter(self, mode: int) -> None:
        """0 automatic, 1 prefilter whenever the dictionary allows, -1 always the full automaton walk"""
        self.L.acb200_set_filter(self.h, int(mode))

    # -- search -----------------------------------------------------------
    def search_events(self, flat, offsets=None, first_only: bool = False) -> np.ndarray:
        """Haystacks laid end to end in `flat`; -> structured array (end, state, text_idx)."""
        buf = _as_u8(flat)
        if offsets is None:
            offsets = np.array([0, buf.size], dtype=np.uint64)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        cap = max(1 << 12, buf.size >> 10)       # (a second call only where events are denser than one per KiB)
        while True:
            ev = np.empty(cap, dtype=EVENT_DTYPE)
            ne = C.c_size_t(0)
            rc = self.L.acb200_search_events(self.h, buf.ctypes.data if buf.size else None, off.ctypes.data, n,
                                             int(first_only), ev.ctypes.data, cap, C.byref(ne))
            if rc != 0:
                raise AcError(last_error())
            if ne.value <= cap:
                return ev[:ne.value]
            cap = int(ne.value)

    def search_hits(self, flat, offsets=None) -> np.ndarray:
        """Hit-level search with device-side expansion -> structured array (text_idx, end, start, pattern);
        `pattern` indexes the accepted patterns (see pattern_ordinal)."""
        buf = _as_u8(flat)
        if offsets is None:
            offsets = np.array([0, buf.size], dtype=np.uint64)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        cap = 1 << 12
        while True:
            hits = np.empty(cap, dtype=HIT_DTYPE)
            nh = C.c_size_t(0)
            rc = self.L.acb200_search_hits(self.h, buf.ctypes.data if buf.size else None, off.ctypes.data, off.size - 1,
                                           hits.ctypes.data, cap, C.byref(nh))
            if rc != 0:
                raise AcError(last_error())
            if nh.value <= cap:
                return hits[:nh.value]
            cap = int(nh.value)

    def pattern_ordinal(self, index: int) -> int:
        """add() ordinal of accepted pattern `index` (this binding stores ordinal+1 in aux)"""
        p = self.L.acb200_pattern(self.h, int(index))
        return int(p.contents.aux) - 1 if p else -1

    def search_device(self, dev_ptr: int, offsets, first_only: bool = False, stream: int = 0):
        """Haystack stream already in HBM. -> (device pointer of packed events, n_events)"""
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = C.c_void_p(0)
        ne = C.c_size_t(0)
        rc = self.L.acb200_search_device(self.h, C.c_void_p(dev_ptr), off.ctypes.data, off.size - 1, int(first_only),
                                         C.c_void_p(stream), C.byref(out), C.byref(ne))
        if rc != 0:
            raise AcError(last_error())
        return out.value, int(ne.value)

    def search_device_uniform(self, dev_ptr: int, n_hay: int, hay_len: int, first_only: bool = False, stream: int = 0):
        """n_hay haystacks of hay_len bytes each, end to end in HBM. -> (device pointer of packed events, n_events)"""
        out = C.c_void_p(0)
        ne = C.c_size_t(0)
        rc = self.L.acb200_search_device_uniform(self.h, C.c_void_p(dev_ptr), int(n_hay), int(hay_len), int(first_only),
                                                 C.c_void_p(stream), C.byref(out), C.byref(ne))
        if rc != 0:
            raise AcError(last_error())
        return out.value, int(ne.value)

    def search_device_uniform_async(self, dev_ptr: int, n_hay: int, hay_len: int, rows_ptr: int, max_events: int, stream: int = 0) -> bool:
        """Enqueues the scan and returns at once: row 0 of the device buffer at rows_ptr receives the event count, rows 1..
        the events.  False if only the synchronous call can serve this batch (see acb200.h)."""
        return self.L.acb200_search_device_uniform_async(self.h, C.c_void_p(dev_ptr), int(n_hay), int(hay_len),
                                                         C.c_void_p(rows_ptr), int(max_events), C.c_void_p(stream)) == 0

    def async_finish(self, n_events: int, dense_tiles: int = 0) -> None:
        """after the caller has waited for the stream: the two values it read from row 0"""
        self.L.acb200_async_finish(self.h, int(n_events), int(dense_tiles))

    def search_flat_tally(self, host_ptr: int, offsets, first_only: bool = False) -> Tally:
        """ac_trie_search_flat() on a HOST buffer (e.g. pinned) with the library's tally callback:
        H2D copy, scan, D2H of the events and the host-side replay all happen inside this call."""
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        t = Tally()
        cb = C.cast(self.L.acb200_tally_cb, BATCH_CB)
        rc = self.L.ac_trie_search_flat(self.h, C.c_void_p(host_ptr), off.ctypes.data, off.size - 1, int(first_only),
                                        cb, C.cast(C.byref(t), C.c_void_p))
        if rc != 0:
            raise AcError(last_error())
        return t

    @staticmethod
    def make_texts(haystacks):
        """list of bytes / uint8 arrays -> (AC_TEXT_t array, n, the arrays that own the bytes)"""
        n = len(haystacks)
        texts = (AcText * max(1, n))()
        keep = []
        for i, h in enumerate(haystacks):
            a = _as_u8(h)
            keep.append(a)
            texts[i].astring = a.ctypes.data if a.size else None
            texts[i].length = a.size
        return texts, n, keep

    def search_batch_tally(self, haystacks=None, first_only: bool = False, texts=None) -> Tally:
        """ac_trie_search_batch() over separately allocated host strings (what the PHP extension holds) with the
        library's tally callback.  `haystacks`: list of bytes / uint8 arrays, or `texts` = make_texts(...) built once."""
        texts, n, _keep = texts if texts is not None else self.make_texts(haystacks)
        t = Tally()
        cb = C.cast(self.L.acb200_tally_cb, BATCH_CB)
        rc = self.L.ac_trie_search_batch(self.h, texts, n, int(first_only), cb, C.cast(C.byref(t), C.c_void_p))
        if rc != 0:
            raise AcError(last_error())
        return t

    def event_digest(self, events: np.ndarray, n_texts: int):
        """per-haystack (event count, order-sensitive event hash) of an event list -> (counts[u64], hashes[u64])"""
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        counts = np.zeros(n_texts, dtype=np.uint64)
        hashes = np.zeros(n_texts, dtype=np.uint64)
        if self.L.acb200_event_digest(self.h, ev.ctypes.data if ev.size else None, ev.size, int(n_texts),
                                      counts.ctypes.data, hashes.ctypes.data) != 0:
            raise AcError(last_error())
        return counts, hashes

    def copy_events(self, dev_ptr: int, max_events: int, stream: int = 0) -> int:
        n = self.L.acb200_copy_events(self.h, C.c_void_p(dev_ptr), int(max_events), C.c_void_p(stream))
        if n < 0:
            raise AcError(last_error())
        return int(n)

    def search_callback(self, text: bytes, keep: bool = False, stop_after_first: bool = False):
        """ac_trie_search() with a recording callback. -> (rc, [(position, [ordinals...]), ...])"""
        got = []

        def cb(mp, _user):
            m = mp.contents
            got.append((int(m.position), [int(m.patterns[j].aux) - 1 for j in range(m.size)]))
            return 1 if stop_after_first else 0

        t = AcText()
        buf = C.create_string_buffer(text, len(text))
        t.astring = C.cast(buf, C.c_void_p)
        t.length = len(text)
        rc = self.L.ac_trie_search(self.h, C.byref(t), int(keep), MATCH_CB(cb), None)
        return rc, got

    def state_patterns(self, state: int):
        """-> list of (ordinal, length) reported by `state`, longest first"""
        pp = C.POINTER(AcPattern)()
        n = self.L.acb200_state_patterns(self.h, int(state), C.byref(pp))
        return [(int(pp[j].aux) - 1, int(pp[j].ptext.length)) for j in range(n)]

    def expand(self, events: np.ndarray):
        """events -> hits in callback order: (text_idx[], pos[], ordinal[], length[])"""
        cache = {}
        ti, pos, pat, ln = [], [], [], []
        for e in events:
            st = int(e["state"])
            lst = cache.get(st)
            if lst is None:
                lst = cache[st] = self.state_patterns(st)
            for (o, l) in lst:
                ti.append(int(e["text_idx"])); pos.append(int(e["end"])); pat.append(o); ln.append(l)
        return (np.array(ti, dtype=np.uint32), np.array(pos, dtype=np.uint64),
                np.array(pat, dtype=np.uint32), np.array(ln, dtype=np.uint32))

    # -- facts ------------------------------------------------------------
    def info(self) -> Info:
        i = Info()
        self.L.acb200_info(self.h, C.byref(i))
        return i

    def stats(self) -> Stats:
        s = Stats()
        self.L.acb200_last_stats(self.h, C.byref(s))
        return s

    def release(self) -> None:
        if self.h:
            self.L.ac_trie_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
