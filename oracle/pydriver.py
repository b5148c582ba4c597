"""ctypes front end of oracle/driver.c — TEST INFRASTRUCTURE ONLY.

Loads one of the three builds of the same driver (see oracle/driver.c):

    Driver("oracle")     CPU restatement            oracle/liboracle_driver.so
    Driver("reference")  reference's own C sources  oracle/_ref/libref_driver.so
    Driver("gpu")        the CUDA product through its C-ABI   oracle/libgpu_driver.so

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "oracle": os.path.join(HERE, "liboracle_driver.so"),
    "reference": os.path.join(HERE, "_ref", "libref_driver.so"),
    "gpu": os.path.join(HERE, "libgpu_driver.so"),
}
_LIBS: dict[str, C.CDLL] = {}


def build(targets=("liboracle_driver.so", "ref", "gpu")) -> None:
    subprocess.run(["make", "-s", "-C", HERE, *targets], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def _load(kind: str) -> C.CDLL:
    if kind in _LIBS:
        return _LIBS[kind]
    lib = C.CDLL(_PATHS[kind])
    lib.drv_create.restype = C.c_void_p
    lib.drv_add.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.drv_add.restype = C.c_int
    lib.drv_add_php_order.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    lib.drv_add_php_order.restype = None
    lib.drv_finalize.argtypes = [C.c_void_p]
    lib.drv_finalize.restype = None
    lib.drv_search.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.drv_search.restype = C.c_long
    lib.drv_last_rc.argtypes = [C.c_void_p]
    lib.drv_last_rc.restype = C.c_int
    lib.drv_release.argtypes = [C.c_void_p]
    lib.drv_release.restype = None
    lib.drv_bench.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                              C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.drv_bench.restype = C.c_double
    lib.drv_bench2.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
    lib.drv_bench2.restype = C.c_double
    _LIBS[kind] = lib
    return lib


def flatten(items) -> tuple[bytes, np.ndarray]:
    """list of bytes -> (concatenation, uint64 offsets[n+1])"""
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    if len(items):
        off[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    return b"".join(items), off


class Driver:
    """Pattern ordinals are 0-based positions in the order patterns were handed in."""

    def __init__(self, kind: str = "oracle"):
        self.kind = kind
        self.lib = _load(kind)
        self.h = self.lib.drv_create()
        if not self.h:
            raise RuntimeError("drv_create failed")

    def add(self, pattern: bytes) -> int:
        return self.lib.drv_add(self.h, pattern, len(pattern))

    def add_php_order(self, patterns) -> None:
        """One ahocorasick_init()/add_patterns() call worth of patterns (array order)."""
        flat, off = flatten(list(patterns))
        self.lib.drv_add_php_order(self.h, flat, off.ctypes.data, len(off) - 1)

    def finalize(self) -> None:
        self.lib.drv_finalize(self.h)

    def search(self, text, keep: bool = False, first_only: bool = False, cap: int | None = None):
        """-> dict(rc, n_hits, n_events, hash, pos[u64], pat[u32], len[u32]) ; hits in callback order"""
        if isinstance(text, (bytes, bytearray)):
            buf = np.frombuffer(bytes(text), dtype=np.uint8)
        else:
            buf = np.ascontiguousarray(text, dtype=np.uint8)
        n = int(buf.size)
        if cap is None:
            cap = 1 << 16
        while True:
            pos = np.empty(cap, dtype=np.uint64)
            pat = np.empty(cap, dtype=np.uint32)
            ln = np.empty(cap, dtype=np.uint32)
            ne = C.c_uint64(0)
            hs = C.c_uint64(0)
            ptr = buf.ctypes.data if n else None
            nh = self.lib.drv_search(self.h, ptr, n, int(keep), int(first_only),
                                     pos.ctypes.data, pat.ctypes.data, ln.ctypes.data, cap,
                                     C.byref(ne), C.byref(hs))
            if nh <= cap or keep:
                break
            cap = int(nh)
        k = min(nh, cap)
        return dict(rc=self.lib.drv_last_rc(self.h), n_hits=int(nh), n_events=int(ne.value),
                    hash=int(hs.value), pos=pos[:k], pat=pat[:k], len=ln[:k])

    def release(self) -> None:
        if self.h:
            self.lib.drv_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def bench(kind: str, patterns, haystacks_flat: np.ndarray, hay_off: np.ndarray, threads: int, reps: int):
    """Times ac_trie_search over the batch. -> (seconds of best repetition, events per pass)"""
    lib = _load(kind)
    pflat, poff = flatten(list(patterns))
    ev = C.c_uint64(0)
    hay = np.ascontiguousarray(haystacks_flat, dtype=np.uint8)
    off = np.ascontiguousarray(hay_off, dtype=np.uint64)
    sec = lib.drv_bench(pflat, poff.ctypes.data, len(poff) - 1, hay.ctypes.data, off.ctypes.data,
                        len(off) - 1, int(threads), int(reps), C.byref(ev))
    return float(sec), int(ev.value)


def bench_digest(kind: str, patterns, haystacks_flat: np.ndarray, hay_off: np.ndarray, threads: int, reps: int = 1,
                 halo: int = 0, digest: bool = True):
    """drv_bench2: like bench(); ONE haystack is cut into per-thread slices with a `halo`-byte warm-up (halo = Lmax-1);
    for batches `digest` also returns every haystack's event count and order-sensitive event hash.
    -> (seconds, events, counts[u64] | None, hashes[u64] | None)"""
    lib = _load(kind)
    pflat, poff = flatten(list(patterns))
    ev = C.c_uint64(0)
    hay = np.ascontiguousarray(haystacks_flat, dtype=np.uint8)
    off = np.ascontiguousarray(hay_off, dtype=np.uint64)
    n = len(off) - 1
    want = digest and n > 1
    counts = np.zeros(n, dtype=np.uint64) if want else None
    hashes = np.zeros(n, dtype=np.uint64) if want else None
    sec = lib.drv_bench2(pflat, poff.ctypes.data, len(poff) - 1, hay.ctypes.data, off.ctypes.data, n, int(threads),
                         int(reps), int(halo), C.byref(ev), counts.ctypes.data if want else None,
                         hashes.ctypes.data if want else None)
    return float(sec), int(ev.value), counts, hashes
