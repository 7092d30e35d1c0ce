"""GPU parity tests of `sigtk ent` (ent.cu through the C-ABI, SGPU_WANT_ENT) against the committed stdout of the
compiled reference and against the CPU oracle. Histogram counts are integers (exact) and the terms are summed in the
reference's order; log2 is CUDA's, so doubles are compared to 1e-12 absolute (entropies are <= 16 bits) and the "%f"
text byte for byte."""
import os

import numpy as np
import pytest

import _fmt
from _oracle import Oracle
import sigtk_b200 as sg
from sigtk_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def ctx():
    c = sg.Context(device=0, max_samples=1 << 23, max_reads=8192)
    yield c
    c.close()


@pytest.mark.parametrize("npz,txt", [("sp1_dna.npz", "ref_sp1_ent.txt"), ("synth_rna.npz", "ref_rna_ent.txt"),
                                     ("ent_adversarial.npz", "ref_ent_adversarial.txt")])
def test_ent_text_equals_reference_stdout(ctx, orc, npz, txt):
    reads = _fmt.load_npz(os.path.join(G, npz))
    res = ctx.run([rd for _, rd in reads], rna=0, want=sg.WANT_ENT)
    got = _fmt.ENT_HDR + "".join(_fmt.ent_line(rid, res.ent[r]) for r, (rid, _) in enumerate(reads))
    assert got == open(os.path.join(G, txt)).read()
    for r, (_, rd) in enumerate(reads):
        assert np.allclose(res.ent[r], orc.ent(rd[0]), rtol=0, atol=TOL), f"read {r}"


def test_ent_full_range_reads_use_the_overflow_histograms(ctx, orc):
    """keys far outside the shared-memory windows (where the reference itself aborts: |delta| >= 16384, ent.c:154),
    twice in a row so that the overflow tables must have been left clean"""
    rng = np.random.default_rng(9)
    reads = [(rng.integers(-32768, 32768, n).astype(np.int16), 8192.0, 0.0, 1400.0) for n in (100000, 3, 65536, 4097)]
    reads.append((np.arange(-32768, 32768, dtype=np.int32).astype(np.int16), 8192.0, 0.0, 1400.0))
    reads += synth.make_reads(6, mean=30000.0, seed=3)
    for _ in range(2):
        res = ctx.run(reads, rna=0, want=sg.WANT_ENT)
        for r, rd in enumerate(reads):
            assert np.allclose(res.ent[r], orc.ent(rd[0]), rtol=0, atol=TOL), f"read {r}: {res.ent[r]} {orc.ent(rd[0])}"


def test_ent_many_reads_more_than_the_grid(ctx, orc):
    """more reads than CTAs (every CTA walks several reads and must clear its windows in between), ragged lengths,
    empty and one-sample records; together with the other outputs"""
    rng = np.random.default_rng(4)
    reads = synth.make_reads(1500, mean=1500.0, seed=8)
    reads[7] = (np.zeros(0, dtype=np.int16), 8192.0, 0.0, 1400.0)
    reads[8] = (np.array([-3], dtype=np.int16), 8192.0, 0.0, 1400.0)
    reads[9] = (rng.integers(-32768, 32768, 5000).astype(np.int16), 8192.0, 0.0, 1400.0)
    res = ctx.run(reads, rna=0, want=sg.WANT_ENT | sg.WANT_STAT)
    for r, rd in enumerate(reads):
        assert np.allclose(res.ent[r], orc.ent(rd[0]), rtol=0, atol=TOL), f"read {r}"
    assert np.array_equal(res.stat[100].view(np.uint32), orc.stat(*reads[100]).view(np.uint32))


def test_ent_long_read_and_svbzd_input(ctx, orc):
    """a 3 M-sample read (one CTA walks it), and the same records as svb-zd streams decoded in HBM"""
    reads = synth.make_reads(1, mean=3.0e6, sigma=0.01, seed=2) + synth.make_reads(5, mean=20000.0, seed=6)
    res = ctx.run(reads, rna=0, want=sg.WANT_ENT)
    for r, rd in enumerate(reads):
        assert np.allclose(res.ent[r], orc.ent(rd[0]), rtol=0, atol=TOL), f"read {r}"
    sres = ctx.run_svbzd([(orc.svbzd_encode(rd[0]), rd[1], rd[2], rd[3]) for rd in reads], rna=0, want=sg.WANT_ENT)
    assert np.array_equal(sres.ent, res.ent)


def test_ent_device_resident_entry(ctx, orc):
    import torch
    reads = synth.make_reads(40, mean=9000.0, seed=12)
    lens = np.array([r[0].shape[0] for r in reads], dtype=np.uint32)
    al = (lens.astype(np.uint64) + 7) // 8 * 8
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(al)
    flat = np.zeros(int(off[-1]), dtype=np.int16)
    for r, rd in enumerate(reads):
        flat[int(off[r]): int(off[r]) + int(lens[r])] = rd[0]
    dev = torch.device("cuda:0")
    d_s = torch.from_numpy(flat).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_len = torch.from_numpy(lens.view(np.int32)).to(dev)
    d_o = torch.zeros(len(reads), dtype=torch.float32, device=dev)
    d_u = torch.ones(len(reads), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    res = ctx.run_device(d_s.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_o.data_ptr(), d_u.data_ptr(),
                         len(reads), int(off[-1]), 0, sg.WANT_ENT)
    got = ctx.d2h(res.ent, np.float64, len(reads) * 3).reshape(-1, 3)
    for r, rd in enumerate(reads):
        assert np.allclose(got[r], orc.ent(rd[0]), rtol=0, atol=TOL)
