"""Seeded synthetic nanopore reads.

Signal model: piecewise-constant level ~ U(60,120) pA, a new level with
probability 0.1 per sample, plus N(0, 2 pA) noise; digitised with
digitisation 8192, range 1402.882324, offset = read_index % 53:
    raw = rint(pA * digitisation / range - offset), clipped to int16.
numpy version (CPU tests, CPU baseline samples) -- bench.py has the torch/CUDA
twin that generates the same model on the device.
"""
from __future__ import annotations

import numpy as np

DIGITISATION = 8192.0
RANGE = 1402.882324
SEED = 20260001


def read_lengths(n_reads: int, mean: float = 40000.0, sigma: float = 0.6, seed: int = SEED,
                 lo: int = 2000, hi: int = 4_000_000) -> np.ndarray:
    """lognormal lengths with the given mean: len = clamp(round(exp(N(mu, sigma^2))), lo, hi)."""
    rng = np.random.default_rng(seed)
    mu = np.log(mean) - 0.5 * sigma * sigma
    ln = np.rint(np.exp(rng.normal(mu, sigma, size=n_reads)))
    return np.clip(ln, lo, hi).astype(np.int64)


def make_read(index: int, length: int, seed: int = SEED, level_lo: float = 60.0, level_hi: float = 120.0,
              p_change: float = 0.1, noise: float = 2.0):
    """One read: (raw int16[length], digitisation, offset, range)."""
    rng = np.random.default_rng([seed, index])
    change = rng.random(length) < p_change
    change[0] = True
    seg = np.cumsum(change) - 1
    levels = rng.uniform(level_lo, level_hi, size=int(seg[-1]) + 1)
    pa = levels[seg] + rng.normal(0.0, noise, size=length)
    offset = float(index % 53)
    raw = np.rint(pa * (DIGITISATION / RANGE) - offset)
    raw = np.clip(raw, -32768, 32767).astype(np.int16)
    return raw, DIGITISATION, offset, RANGE


def make_reads(n_reads: int, mean: float = 40000.0, sigma: float = 0.6, seed: int = SEED, lo: int = 2000,
               hi: int = 4_000_000, rna: bool = False):
    """A list of reads with lognormal lengths. RNA-like reads dwell ~4x longer per level."""
    lens = read_lengths(n_reads, mean, sigma, seed, lo, hi)
    p = 0.025 if rna else 0.1
    return [make_read(i, int(lens[i]), seed, p_change=p) for i in range(n_reads)]


def svbzd_encode(raw: np.ndarray) -> np.ndarray:
    """The svb-zd stream slow5lib stores for `raw` in a BLOW5 record (SLOW5_COMPRESS_SVB_ZD; slow5_press.c:1055-1089,
    streamvbyte_encode.c, streamvbyte_zigzag.c:5-25): uint32 count | 2-bit length codes, four per key byte | the
    zigzag-coded deltas in 1..4 little-endian bytes. Vectorised numpy; used to make synthetic COMPRESSED input for the
    benchmarks and the CLI tools (the decoder under test is the CUDA one; tests check this encoder against the
    reference's own)."""
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    n = raw.shape[0]
    cur = raw.astype(np.int32)
    prev = np.concatenate([np.zeros(1, np.int32), cur[:-1]]) if n else cur
    d = (cur - prev).astype(np.int32)
    z = ((d.astype(np.uint32) << np.uint32(1)) ^ (d >> 31).astype(np.uint32)).astype(np.uint32)
    code = (z >= 1 << 8).astype(np.uint8) + (z >= 1 << 16).astype(np.uint8) + (z >= 1 << 24).astype(np.uint8)
    pad = (-n) % 4
    c4 = np.concatenate([code, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    keys = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    zb = z.astype("<u4").view(np.uint8).reshape(n, 4)
    data = zb[np.arange(4, dtype=np.uint8)[None, :] <= code[:, None]]
    return np.concatenate([np.array([n], dtype="<u4").view(np.uint8), keys, data])
