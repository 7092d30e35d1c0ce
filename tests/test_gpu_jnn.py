"""GPU parity tests of `sigtk jnn` (jnn.cu + the clamped-signal moments of stat.cu through the C-ABI, SGPU_WANT_JNN)
against the committed stdout of the compiled reference and against the CPU oracle. Segment boundaries are integers
that depend on two order-dependent float sums (the band); everything is compared exactly."""
import os

import numpy as np
import pytest

import _fmt
from _oracle import Oracle
import sigtk_b200 as sg
from sigtk_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def ctx():
    c = sg.Context(device=0, max_samples=1 << 23, max_reads=8192)
    yield c
    c.close()


@pytest.mark.parametrize("npz,txt,rna_flag", [("sp1_dna.npz", "ref_sp1_jnn", 0), ("synth_rna.npz", "ref_rna_jnn", 1),
                                              ("jnn_stalls_dna.npz", "ref_jnn_stalls_dna", 0),
                                              ("jnn_stalls_rna.npz", "ref_jnn_stalls_rna", 1)])
def test_jnn_text_equals_reference_stdout(ctx, npz, txt, rna_flag):
    reads = _fmt.load_npz(os.path.join(G, npz))
    res = ctx.run([rd for _, rd in reads], rna=rna_flag, want=sg.WANT_JNN)
    for compact, suffix in ((False, ".txt"), (True, "_c.txt")):
        got = _fmt.JNN_HDR + "".join(_fmt.jnn_line(rid, len(rd[0]), res.jnn[r], compact)
                                     for r, (rid, rd) in enumerate(reads))
        assert got == open(os.path.join(G, txt + suffix)).read(), suffix


def _stall_reads(rng, n_reads, rna_flag, scale=1.0):
    win = 1000 if rna_flag else 150
    out = []
    for k in range(n_reads):
        n = int(rng.integers(2 * win, int(40 * win * scale)))
        raw = (int(rng.integers(300, 700)) + rng.integers(20, 120) * rng.standard_normal(n)).astype(np.int64)
        pos = int(rng.integers(0, 2 * win))
        centre = int(np.mean(raw))
        while pos < n:
            stop = min(n, pos + int(rng.integers(win // 6, 5 * win)))
            raw[pos:stop] = centre + rng.integers(-6, 7, stop - pos)
            if k % 2 and stop - pos > 60:
                raw[rng.integers(pos + 3, stop - 3, int(rng.integers(1, 9)))] = centre + 400
            pos = stop + int(rng.integers(3, 90) if k % 3 else rng.integers(30, 8 * win))
        if k % 5 == 0:
            raw[::911] = 9000
            raw[3::1777] = -900
        out.append((np.clip(raw, -32768, 32767).astype(np.int16), 8192.0, float(k % 7), 1400.0))
    return out


@pytest.mark.parametrize("rna_flag,seed", [(0, 1), (0, 2), (1, 3)])
def test_jnn_stall_batches_vs_oracle(ctx, orc, rna_flag, seed):
    """hundreds of seeded reads with stalls, outliers inside them, near-merge gaps and clipped spikes, more reads than
    one wave of threads, ragged lengths, with the other outputs requested in the same batch"""
    rng = np.random.default_rng(seed)
    reads = _stall_reads(rng, 400 if not rna_flag else 60, rna_flag)
    reads[3] = (np.zeros(0, dtype=np.int16), 8192.0, 0.0, 1400.0)
    reads[4] = (np.array([700], dtype=np.int16), 8192.0, 0.0, 1400.0)
    reads[5] = (np.full(3000, 1300, dtype=np.int16), 8192.0, 0.0, 1400.0)
    res = ctx.run(reads, rna=rna_flag, want=sg.WANT_JNN | sg.WANT_STAT | sg.WANT_ENT)
    n_seg = 0
    for r, rd in enumerate(reads):
        exp = orc.jnn(rd[0], rna_flag)
        assert np.array_equal(res.jnn[r], exp), f"read {r}: {res.jnn[r][:4]} vs {exp[:4]}"
        n_seg += len(exp)
    assert n_seg > len(reads)
    assert np.array_equal(res.stat[10].view(np.uint32), orc.stat(*reads[10]).view(np.uint32))


def test_jnn_synthetic_reads_long_read_and_svbzd(ctx, orc):
    reads = synth.make_reads(200, mean=8000.0, seed=21) + _stall_reads(np.random.default_rng(8), 3, 0, scale=40.0)
    res = ctx.run(reads, rna=0, want=sg.WANT_JNN)
    for r, rd in enumerate(reads):
        assert np.array_equal(res.jnn[r], orc.jnn(rd[0], 0)), f"read {r}"
    sres = ctx.run_svbzd([(orc.svbzd_encode(rd[0]), rd[1], rd[2], rd[3]) for rd in reads], rna=0, want=sg.WANT_JNN)
    for r in range(len(reads)):
        assert np.array_equal(sres.jnn[r], res.jnn[r])
