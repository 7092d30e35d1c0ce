"""GPU parity tests of `sigtk prefix` (SGPU_WANT_PREFIX; reference src/cfunc.c:169-234, src/jnn.c:99-188, 352-374):
the CUDA path through the C-ABI against the CPU oracle (orc_adaptor_polya), against the committed stdout of the
compiled reference (`sigtk prefix`, `sigtk prefix --print-stat`) and against the reference's own golden
test/prefix_dna.exp. Positions and all six statistics bit-exact."""
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
    c = sg.Context(device=0, max_samples=1 << 23, max_reads=4096)
    yield c
    c.close()


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def check(orc, res, reads, rna_flag):
    for r, rd in enumerate(reads):
        pos, st = orc.adaptor_polya(*rd, rna=rna_flag)
        got_pos, got_st = res.prefix_pos[r].astype(np.int64), res.prefix_stat[r]
        assert np.array_equal(got_pos, pos), f"read {r}: positions {got_pos} != {pos}"
        if pos[1] > 0:
            assert np.array_equal(bits(got_st[:3]), bits(st[:3])), f"read {r}: adaptor statistics {got_st[:3]} != {st[:3]}"
        if pos[3] > 0:
            assert np.array_equal(bits(got_st[3:]), bits(st[3:])), f"read {r}: poly-A statistics {got_st[3:]} != {st[3:]}"


@pytest.mark.parametrize("npz,txt,rna_flag", [("sp1_dna.npz", "ref_sp1_prefix_stat.txt", 0),
                                              ("synth_rna.npz", "ref_rna_prefix_stat.txt", 1),
                                              ("prefix_adaptor_dna.npz", "ref_prefix_adaptor_dna_stat.txt", 0),
                                              ("prefix_adaptor_rna.npz", "ref_prefix_adaptor_rna_stat.txt", 1)])
def test_prefix_equals_reference_stdout(ctx, orc, npz, txt, rna_flag):
    reads = _fmt.load_npz(os.path.join(G, npz))
    res = ctx.run([rd for _, rd in reads], rna=rna_flag, want=sg.WANT_PREFIX)
    check(orc, res, [rd for _, rd in reads], rna_flag)
    rows = [(rid, len(rd[0]), res.prefix_pos[r], res.prefix_stat[r]) for r, (rid, rd) in enumerate(reads)]
    got = _fmt.PREFIX_HDR + _fmt.PREFIX_HDR_STAT + "\n" + "".join(_fmt.prefix_line(r, n, p, s, True) for r, n, p, s in rows)
    assert got == open(os.path.join(G, txt)).read()
    plain = txt.replace("_stat.txt", ".txt")
    if os.path.exists(os.path.join(G, plain)):
        got = _fmt.PREFIX_HDR + "\n" + "".join(_fmt.prefix_line(r, n, p, s, False) for r, n, p, s in rows)
        assert got == open(os.path.join(G, plain)).read()


def test_prefix_dna_exp(ctx):
    """the reference's own golden test/prefix_dna.exp (scripts/test.sh:54-56)"""
    reads = _fmt.load_npz(os.path.join(G, "sp1_dna.npz"))
    res = ctx.run([rd for _, rd in reads], rna=0, want=sg.WANT_PREFIX)
    got = _fmt.PREFIX_HDR + "\n" + "".join(_fmt.prefix_line(rid, len(rd[0]), res.prefix_pos[r], res.prefix_stat[r])
                                           for r, (rid, rd) in enumerate(reads))
    assert got == open(os.path.join(G, "prefix_dna.exp")).read()


def adaptor_read(k, n, seed, polya=True, rna_like=True):
    """a direct-RNA-like read: a low adaptor stretch, a poly-A plateau about 30 pA above it, then the transcript"""
    rng = np.random.default_rng([seed, k])
    rd = synth.make_read(k, n, seed=seed, p_change=0.025 if rna_like else 0.1)
    raw = rd[0].astype(np.int32)
    a0 = int(rng.integers(500, 3000))
    a1 = a0 + int(rng.integers(2500, 9000))
    scale = synth.DIGITISATION / synth.RANGE
    if a1 + 4000 < n:
        lvl = 55.0 + 10.0 * rng.random()
        raw[a0:a1] = np.rint((lvl + rng.normal(0.0, 1.5, a1 - a0)) * scale - rd[2])
        if polya:
            p1 = a1 + int(rng.integers(300, 2500))
            seg = np.rint((lvl + 30.0 + rng.normal(0.0, 2.5, p1 - a1)) * scale - rd[2])
            spikes = rng.random(p1 - a1) < 0.03          # tolerated outliers inside the stretch
            seg[spikes] += 400
            raw[a1:p1] = seg
    return np.clip(raw, -32768, 32767).astype(np.int16), rd[1], rd[2], rd[3]


@pytest.mark.parametrize("rna_flag", [0, 1])
def test_prefix_seeded_reads(ctx, orc, rna_flag):
    """adaptor and poly-A like stretches of random position and length, reads around the window length, reads
    without an adaptor; every rolling-mean / band decision and every float statistic bit-exact"""
    reads = [adaptor_read(k, int(n), seed=17) for k, n in enumerate([30000, 45000, 60000, 90000, 25000, 150000, 33333])]
    reads += [adaptor_read(50 + k, 40000, seed=18, polya=False) for k in range(3)]
    base = synth.make_read(3, 9000, seed=9)
    reads += [(base[0][:n].copy(), base[1], base[2], base[3]) for n in (1, 7, 1999, 2000, 2001, 2002, 2033, 3024, 4001, 9000)]
    flat = (np.full(20000, 500, np.int16), base[1], base[2], base[3])      # constant signal: every rolling mean equal
    reads.append(flat)
    res = ctx.run(reads, rna=rna_flag, want=sg.WANT_PREFIX)
    check(orc, res, reads, rna_flag)
    found = sum(1 for r in range(len(reads)) if res.prefix_pos[r][1] > 0)
    assert found >= 5
    if rna_flag:
        assert sum(1 for r in range(len(reads)) if res.prefix_pos[r][3] > 0) >= 3


def test_prefix_with_everything_else(ctx, orc):
    """SGPU_WANT_PREFIX next to the other outputs of the same batch"""
    reads = [adaptor_read(k, 40000, seed=21) for k in range(4)]
    res = ctx.run(reads, rna=1, want=sg.WANT_PREFIX | sg.WANT_EVENTS | sg.WANT_STAT)
    check(orc, res, reads, 1)
    for r, rd in enumerate(reads):
        assert np.array_equal(res.events(r).start, orc.events(*rd, rna=1)[0])
        assert np.array_equal(res.stat[r].view(np.uint32), orc.stat(*rd).view(np.uint32))
