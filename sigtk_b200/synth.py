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
