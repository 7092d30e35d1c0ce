"""CPU tests of the chunk walker logic (sigtk_b200/csrc/walk_core.cuh compiled for the host by tests/_hostwalk.py)
against the oracle: event boundaries and pA bit-exact for every chunk length / warm-up / bitmap shift, and the
soundness of the boundary-state check (a read without a mismatch is exact even with a far too short warm-up)."""
import numpy as np
import pytest

import _hostwalk
from _oracle import Oracle
from sigtk_b200 import synth


@pytest.fixture(scope="module")
def lib():
    return _hostwalk.load()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def check(lib, orc, rd, rna, L, W, sh, must_verify=True):
    raw, dig, off, rng = rd
    st = orc.events(raw, dig, off, rng, rna=rna)[0].astype(np.int64)
    mism, starts, pa, mm = _hostwalk.walk(lib, raw, dig, off, rng, rna, L, W, sh)
    if must_verify:
        assert mism == 0
    if mism == 0:
        assert np.array_equal(starts, st), f"n={len(raw)} L={L} W={W} sh={sh}"
    assert np.array_equal(pa.view(np.uint32), orc.pa(raw, dig, off, rng).view(np.uint32))
    assert mm == (int(raw.min()), int(raw.max()))
    return mism


@pytest.mark.parametrize("rna,L,W", [(0, 128, 64), (0, 256, 64), (0, 1024, 64), (1, 512, 384), (1, 1024, 384)])
def test_walker_equals_oracle(lib, orc, rna, L, W):
    reads = synth.make_reads(6, mean=9000.0, seed=41 + rna, rna=bool(rna))
    for k, rd in enumerate(reads):
        check(lib, orc, rd, rna, L, W, (0, 8, 16, 24)[k % 4])


@pytest.mark.parametrize("rna,L,W", [(0, 128, 64), (1, 512, 384)])
def test_ragged_lengths(lib, orc, rna, L, W):
    """lengths around the window sizes, the block size and the chunk length (single chunk, 2 chunks, short last chunk)"""
    base = synth.make_read(3, 3 * L + 40, seed=9)
    for n in [1, 2, 5, 6, 7, 8, 9, 12, 13, 14, 15, 16, 17, 27, 28, 29, 31, 33, 63, 65, 199, 200, L - 1, L, L + 1, L + 7,
              L + 8, L + 9, 2 * L - 1, 2 * L, 2 * L + 1, 3 * L, 3 * L + 5]:
        check(lib, orc, (base[0][:n].copy(), base[1], base[2], base[3]), rna, L, W, 8)


def test_short_warmup_is_detected_not_trusted(lib, orc):
    """with an 8-sample warm-up many chunks start from a wrong state: the boundary check must say so, and every
    read it passes must still be exact"""
    reads = synth.make_reads(12, mean=15000.0, seed=5)
    flagged = sum(check(lib, orc, rd, 0, 128, 8, 0, must_verify=False) > 0 for rd in reads)
    assert flagged > 0


def test_flat_stretch(lib, orc):
    """a long constant stretch: t-statistics exactly 0 (delta == 0 with the variance floor); SURVEY 7.3(2)"""
    rd = synth.make_read(11, 20000, seed=3)
    raw = rd[0].copy()
    raw[6000:14000] = raw[5999]
    for rna, L, W in ((0, 256, 64), (1, 512, 384)):
        check(lib, orc, (raw, rd[1], rd[2], rd[3]), rna, L, W, 0, must_verify=False)


def test_exact_recompute_path(orc):
    """the rare path (a t-statistic next to a float rounding midpoint: the block is recomputed with the reference's
    own operations) forced on every third block, first / interior / last chunks, DNA and RNA"""
    lib = _hostwalk.load(force_redo=True)
    for rna, L, W in ((0, 128, 64), (0, 1024, 64), (1, 512, 384)):
        reads = synth.make_reads(4, mean=7000.0, seed=77 + rna, rna=bool(rna))
        for k, rd in enumerate(reads):
            check(lib, orc, rd, rna, L, W, (0, 8, 16, 24)[k % 4])
    base = synth.make_read(3, 600, seed=10)
    for n in (5, 13, 29, 199, 200, 257, 600):
        check(lib, orc, (base[0][:n].copy(), base[1], base[2], base[3]), 0, 128, 64, 8)


def test_old_peaks_take_the_exact_path(orc):
    """a peak older than the block's register mask (forced here: older than the earliest possible emission) makes
    the block record its peaks one by one; events must not change"""
    lib = _hostwalk.load(force_far=True)
    for rna, L, W in ((0, 128, 64), (0, 1024, 64), (1, 512, 384)):
        reads = synth.make_reads(4, mean=7000.0, seed=99 + rna, rna=bool(rna))
        for k, rd in enumerate(reads):
            check(lib, orc, rd, rna, L, W, (0, 8, 16, 24)[k % 4])


def ramp_read(n=12000, seed=3):
    """linear ramps give a t-statistic that stays within peak_height of its maximum for as long as the ramp lasts:
    the peak is emitted dozens of samples after its position (older than the walker's per-block register mask)"""
    rd = synth.make_read(5, n, seed=seed)
    raw = rd[0].copy()
    for start, step, ln in ((1300, 20, 150), (2400, -2, 180), (3500, 7, 90), (4500, 11, 40), (7000, 3, 60), (9000, 30, 100)):
        raw[start:start + ln] = raw[start - 1] + step * np.arange(1, ln + 1)
    assert raw.min() > 0
    return raw, rd[1], rd[2], rd[3]


def test_plateaus_emit_old_peaks(lib, orc):
    rd = ramp_read()
    for rna, L, W in ((0, 4096, 64), (0, 2048, 64), (1, 4096, 384)):
        for sh in (0, 8, 24):
            check(lib, orc, rd, rna, L, W, sh)


def glitch_read(n=9000, seed=12, where=(0, 1, 7, 8, 100, 127, 128, 129, 1023, 1024, 1031, 2000, 2001, 2002, 2003, 4000)):
    """samples whose pA is zero or negative (raw <= -offset): about 3 reads in 100 of real R9.4 data have some"""
    rd = synth.make_read(7, n, seed=seed)
    raw = rd[0].copy()
    off = int(rd[2])
    vals = [-off, -off - 1, -off - 300, -32768, -off, 0 - off - 7]
    for k, p in enumerate(list(where) + [n - 1, n - 2, n - 9]):
        if p < n:
            raw[p] = vals[k % len(vals)]
    return raw, rd[1], rd[2], rd[3]


@pytest.mark.parametrize("rna,L,W", [(0, 128, 64), (0, 1024, 64), (1, 512, 384), (1, 4096, 384)])
def test_non_positive_samples_stay_on_the_walker(lib, orc, rna, L, W):
    for seed in (12, 13):
        rd = glitch_read(9000, seed)
        assert (orc.pa(*rd) <= 0).sum() >= 10
        for sh in (0, 16):
            check(lib, orc, rd, rna, L, W, sh)
    # a read that is non-positive throughout, and one with a single zero
    rd = synth.make_read(8, 3000, seed=5)
    check(lib, orc, ((-rd[0] - 100).astype(np.int16), rd[1], rd[2], rd[3]), rna, L, W, 8, must_verify=False)
    raw = rd[0].copy(); raw[1500] = -int(rd[2])
    check(lib, orc, (raw, rd[1], rd[2], rd[3]), rna, L, W, 8)


# ---- the long detector: not stepped on the fast path, its lives replayed as jobs (walk_core.cuh) ------------------
def check_thr(lib, orc, rd, rna, L, W, sh, thr, jobs):
    raw, dig, off, rng = rd
    st = orc.event_starts(raw, dig, off, rng, rna=rna, thr_long=thr)
    mism, starts, _, _ = _hostwalk.walk(lib, raw, dig, off, rng, rna, L, W, sh, thr_long=thr, jobs_out=jobs)
    if mism == 0:
        assert np.array_equal(starts, st), f"n={len(raw)} L={L} W={W} sh={sh} thr={thr}"
    return mism, st


def test_event_starts_helper_equals_event_read(orc):
    rd = synth.make_read(2, 9000, seed=4)
    assert np.array_equal(orc.event_starts(*rd), orc.events(*rd)[0].astype(np.int64))


@pytest.mark.parametrize("rna,L,W", [(0, 128, 64), (0, 256, 64), (0, 1024, 64), (1, 512, 384), (1, 2048, 384)])
@pytest.mark.parametrize("thr", [0.3, 1.0, 2.5, 5.0])
def test_long_detector_lives_are_replayed(lib, orc, rna, L, W, thr):
    """with the reference's threshold (9.0) the long detector never emits on these reads; lowered, it emits all the
    time: lives inside a chunk, lives that began in an earlier chunk (LS_PRED), lives that run past the chunk (LS_CONT)"""
    reads = synth.make_reads(5, mean=9000.0, seed=141 + rna, rna=bool(rna))
    jobs, extra = [], 0
    for k, rd in enumerate(reads):
        mism, st = check_thr(lib, orc, rd, rna, L, W, (0, 8, 16, 24)[k % 4], thr, jobs)
        assert mism == 0
        extra += len(np.setdiff1d(st, orc.events(*rd, rna=rna)[0].astype(np.int64)))
    assert sum(jobs) > 0
    if thr <= 2.5:
        assert extra > 0   # the long detector did emit peaks the short one does not


def test_long_detector_quiet_stretches(lib, orc):
    """long lives (no short peak for thousands of samples: a flat stretch, a slow ramp) across many chunk boundaries"""
    rd = synth.make_read(11, 30000, seed=3)
    raw = rd[0].copy()
    raw[6000:14000] = raw[5999]
    raw[14000:20000] = raw[5999] + (np.arange(6000) // 40)
    jobs = []
    for rna, L, W in ((0, 128, 64), (0, 1024, 64), (1, 512, 384)):
        for thr in (0.05, 0.5, 9.0):
            check_thr(lib, orc, (raw, rd[1], rd[2], rd[3]), rna, L, W, 8, thr, jobs)


def test_long_detector_ragged_and_edges(lib, orc):
    base = synth.make_read(3, 3 * 128 + 40, seed=9)
    jobs = []
    for n in [1, 5, 13, 14, 27, 28, 29, 33, 127, 128, 129, 135, 136, 137, 255, 256, 257, 3 * 128, 3 * 128 + 5]:
        for thr in (0.3, 1.5):
            for rna in (0, 1):
                check_thr(lib, orc, (base[0][:n].copy(), base[1], base[2], base[3]), rna, 128 if not rna else 512,
                          64 if not rna else 384, 8, thr, jobs)


def test_far_from_the_pivot(lib, orc):
    """samples more than ZMAX raw units from the chunk's pivot (level jumps of thousands of units, spikes): the
    integer sums are clamped there and every window that holds such a sample is a candidate"""
    rd = synth.make_read(21, 12000, seed=8)
    raw = rd[0].astype(np.int32)
    raw[3000:5000] += 3000
    raw[7000] = 30000
    raw[7001] = -20000
    raw[9000:9100] += np.arange(100) * 40
    raw = np.clip(raw, -32768, 32767).astype(np.int16)
    jobs = []
    for rna, L, W in ((0, 128, 64), (0, 1024, 64), (1, 512, 384)):
        for thr in (1.0, 9.0):
            mism, _ = check_thr(lib, orc, (raw, rd[1], rd[2], rd[3]), rna, L, W, 0, thr, jobs)


def test_every_hot_life_is_replayed(lib, orc):
    """the filter that replaces the long detector on the fast path is conservative: the walker replays at least the
    lives in which a stepped position has t2 > thr_long (counted with the oracle, stepping its detector position by
    position) -- at the reference's threshold, where the long detector never emits and parity alone could not tell"""
    reads = [synth.make_read(k, 12000, seed=77) for k in range(3)]
    for thr in (9.0, 6.0):
        jobs, need = [], 0
        for rd in reads:
            need += orc.hot_lives(*rd, rna=0, thr_long=thr)
            check_thr(lib, orc, rd, 0, 32768, 64, 0, thr, jobs)   # one chunk per read: one job per life
        assert need > 0 and sum(jobs) >= need, (thr, need, sum(jobs))
        assert sum(jobs) <= 40 * need + 40                        # ... and not wildly more
