"""Seeded synthetic workloads of BASELINE.json's five configs.

The reference's benchmark (examples/benchmark.php:13-31) draws from an unseeded
rand(); these generators keep its shapes but use fixed numpy seeds so that the
CPU oracle and the GPU see identical bytes.  Random `abcdef` data almost never
contains a 16-byte needle, so needles are planted at seeded offsets (including
offset 0 and the last possible offset).
"""
from __future__ import annotations

import numpy as np

SEED = 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF

# config 1 — README.md:72-94 / tests/test1.phpt:11-25
CFG1_PATTERNS = [
    {"key": "ab", "value": "alfa"},
    {"key": "ac", "value": "beta"},
    {"key": "ad", "value": "gamma", "aux": [1]},
    {"key": "ae", "value": "delta"},
    {"id": 0, "value": "zeta"},
    {"key": "ag", "value": "omega"},
    {"value": "lfa"},
]
CFG1_HAYSTACK = b"alFABETA gamma zetaomegaalfa!"


def _alphabet_bytes(rng, n, alphabet: bytes) -> np.ndarray:
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    idx = rng.integers(0, len(alphabet), size=n, dtype=np.uint8)
    if len(alphabet) > 1 and bytes(range(alphabet[0], alphabet[0] + len(alphabet))) == alphabet:
        idx += np.uint8(alphabet[0])          # a contiguous alphabet needs no table (same bytes, 25x cheaper)
        return idx
    return lut[idx]


def offsets_uniform(n: int, length: int) -> np.ndarray:
    return np.arange(n + 1, dtype=np.uint64) * np.uint64(length)


def cfg2(n_hay: int = 256, hay_len: int = 8192, n_needles: int = 2048, needle_len: int = 16,
         planted_per_hay: int = 8, alphabet: bytes = b"abcdef", seed: int = SEED):
    """benchmark.php shape. -> (needles: list[bytes], hay: uint8[n_hay*hay_len], offsets: uint64[n_hay+1])"""
    rng = np.random.default_rng(seed)
    needles_arr = _alphabet_bytes(rng, n_needles * needle_len, alphabet).reshape(n_needles, needle_len)
    hay = _alphabet_bytes(rng, n_hay * hay_len, alphabet).reshape(n_hay, hay_len)
    if planted_per_hay and hay_len >= needle_len:
        k = planted_per_hay
        which = rng.integers(0, n_needles, size=(n_hay, k))
        pos = rng.integers(0, hay_len - needle_len + 1, size=(n_hay, k))
        pos[:, 0] = 0
        if k > 1:
            pos[:, 1] = hay_len - needle_len
        cols = np.arange(needle_len)
        for j in range(k):
            idx = pos[:, j][:, None] + cols[None, :]
            hay[np.arange(n_hay)[:, None], idx] = needles_arr[which[:, j]]
    needles = [needles_arr[i].tobytes() for i in range(n_needles)]
    return needles, hay.reshape(-1), offsets_uniform(n_hay, hay_len)


def cfg2_needles(n_needles: int = 2048, needle_len: int = 16, alphabet: bytes = b"abcdef", seed: int = SEED):
    """the dictionary of cfg2() alone -> (needles: list[bytes], needles_arr: uint8[n_needles, needle_len])"""
    rng = np.random.default_rng(seed)
    arr = _alphabet_bytes(rng, n_needles * needle_len, alphabet).reshape(n_needles, needle_len)
    return [arr[i].tobytes() for i in range(n_needles)], arr


def cfg2_stream(stream_seed: int, first_block: int, n_blocks: int, block_hays: int = 256, hay_len: int = 8192,
                planted_per_hay: int = 8, alphabet: bytes = b"abcdef", seed: int = SEED):
    """config 2 at steady state as a stream of independent 256-haystack blocks: block b of stream `stream_seed`
    (one stream per GPU rank) is seeded by (seed, stream_seed, b) alone, so any range of blocks can be
    regenerated anywhere — the GPU arm, the CPU reference arm's bounded sample, the parity check — and all
    see the same bytes.  Every block is distinct (nothing is tiled).  The dictionary is cfg2_needles(seed).
    -> uint8[n_blocks * block_hays * hay_len]"""
    _, needles_arr = cfg2_needles(alphabet=alphabet, seed=seed)
    n_needles, needle_len = needles_arr.shape
    out = np.empty((n_blocks * block_hays, hay_len), dtype=np.uint8)
    cols = np.arange(needle_len)
    rows = np.arange(block_hays)[:, None]

    def make(b):
        rng = np.random.default_rng([seed & 0xFFFFFFFF, stream_seed, first_block + b])
        hay = _alphabet_bytes(rng, block_hays * hay_len, alphabet).reshape(block_hays, hay_len)
        k = planted_per_hay
        if k and hay_len >= needle_len:
            which = rng.integers(0, n_needles, size=(block_hays, k))
            pos = rng.integers(0, hay_len - needle_len + 1, size=(block_hays, k))
            pos[:, 0] = 0
            if k > 1:
                pos[:, 1] = hay_len - needle_len
            for j in range(k):
                hay[rows, pos[:, j][:, None] + cols[None, :]] = needles_arr[which[:, j]]
        out[b * block_hays:(b + 1) * block_hays] = hay

    if n_blocks >= 8:                      # numpy's generators release the GIL: blocks are independent
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            list(pool.map(make, range(n_blocks)))
    else:
        for b in range(n_blocks):
            make(b)
    return out.reshape(-1)


def cfg3(n_patterns: int = 100_000, min_len: int = 8, max_len: int = 64, hay_bytes: int = 1 << 30,
         plant_every: int = 1 << 20, seed: int = SEED + 3):
    """virus-signature shape: binary patterns of 8..64 B over one binary haystack, one planted
    signature per `plant_every` bytes (some straddling multiples of 4 KiB)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n_patterns)
    blob = rng.integers(0, 256, size=int(lens.sum()), dtype=np.uint8)
    offs = np.zeros(n_patterns + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    patterns = [blob[offs[i]:offs[i + 1]].tobytes() for i in range(n_patterns)]
    hay = rng.integers(0, 256, size=hay_bytes, dtype=np.uint8)
    n_plant = max(1, hay_bytes // plant_every)
    for j in range(n_plant):
        p = patterns[int(rng.integers(0, n_patterns))]
        base = j * plant_every
        if j % 4 == 1:
            at = base + 4096 * int(rng.integers(1, max(2, plant_every // 4096))) - len(p) // 2
        else:
            at = base + int(rng.integers(0, max(1, plant_every - len(p))))
        at = max(0, min(at, hay_bytes - len(p)))
        if at + len(p) <= hay_bytes:
            hay[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return patterns, hay, np.array([0, hay_bytes], dtype=np.uint64)


def cfg5(n_patterns: int = 4096, hay_bytes: int = 256 << 20):
    """adversarial: 'a', 'aa', ... (only the first 1024 are accepted — AC_PATTRN_MAX_LENGTH) over 'aaaa…'."""
    patterns = [b"a" * (i + 1) for i in range(n_patterns)]
    hay = np.full(hay_bytes, ord("a"), dtype=np.uint8)
    return patterns, hay, np.array([0, hay_bytes], dtype=np.uint64)


def cfg5_expected(hay_bytes: int, n_patterns: int = 4096, max_len: int = 1024):
    """closed form for cfg5: (events, hits)"""
    m = min(n_patterns, max_len)
    head = min(hay_bytes, m)
    hits = head * (head + 1) // 2 + max(0, hay_bytes - m) * m
    return hay_bytes, hits
