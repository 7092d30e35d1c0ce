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


# ---- the counter-based generator of the named benchmark set ----------------------------------------------------------
# Every sample is a pure function of (seed, read index, sample index): any rank, GPU or CPU regenerates any read of
# the 1,000,000-read set bit for bit ("counter-based ... so any GPU/CPU regenerates any read").
# Only integer arithmetic and exactly rounded float64 operations are used -- no libm call whose last bit could differ
# between numpy and CUDA: the N(0, 2 pA) noise is the sum of twelve 16-bit uniforms (Irwin-Hall: mean 0, standard
# deviation 2 pA, support +-12 pA), the level changes and levels come from 64-bit hashes.
# bench.py's device generator (torch, int64 with wrap-around) is the same arithmetic: tests/test_synth.py.
_M1, _M2, _M3 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0x9E3779B97F4A7C15
_K_SAMPLE, _K_STREAM = 0xD1B54A32D192ED03, 0x8CB92BA72F3D8DD7
_U64 = (1 << 64) - 1


def _mix_np(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 (wrap-around arithmetic)"""
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(_M1)
    x ^= x >> np.uint64(27)
    x *= np.uint64(_M2)
    x ^= x >> np.uint64(31)
    return x


def read_key(seed: int, index) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix_np(np.uint64((seed * _M3) & _U64) + np.asarray(index, dtype=np.uint64))


def _hash_np(key, i: np.ndarray, stream: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix_np(key ^ (i.astype(np.uint64) * np.uint64(_K_SAMPLE) + np.uint64((stream * _K_STREAM) & _U64)))


def make_read_cb(index: int, length: int, seed: int = SEED, level_lo: float = 60.0, level_hi: float = 120.0,
                 p_change: float = 0.1, noise: float = 2.0):
    """One read of the counter-based set: (raw int16[length], digitisation, offset, range)."""
    with np.errstate(over="ignore"):
        key = read_key(seed, index)
        i = np.arange(length, dtype=np.uint64)
        change = (_hash_np(key, i, 1) >> np.uint64(40)) < np.uint64(int(p_change * (1 << 24)))
        if length:
            change[0] = True
        start = np.maximum.accumulate(np.where(change, i.astype(np.int64), -1))
        u = (_hash_np(key, start.astype(np.uint64), 2) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        level = level_lo + (level_hi - level_lo) * u
        tot = np.zeros(length, dtype=np.int64)
        for s in (3, 4, 5):
            h = _hash_np(key, i, s)
            for k in range(4):
                tot += ((h >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.int64)
        pa = level + (tot - 393210).astype(np.float64) * (noise / 65536.0)
    offset = float(index % 53)
    raw = np.rint(pa * (DIGITISATION / RANGE) - offset)
    return np.clip(raw, -32768, 32767).astype(np.int16), DIGITISATION, offset, RANGE


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
