"""GPU parity tests: the CUDA path (through the C-ABI, include/sigtk_b200.h) against the CPU oracle and
against the committed stdout of the compiled reference. Bit-exact for boundaries, pA, means, stdv and stat."""
import contextlib
import gzip
import os

import numpy as np
import pytest

import _fmt
from _oracle import Oracle
import sigtk_b200 as sg
from sigtk_b200 import _lib as _sl
from sigtk_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = sg.WANT_EVENTS | sg.WANT_PA | sg.WANT_STAT


@contextlib.contextmanager
def tuned(ctx, chunk_len=0, warmup=0, thr_long=9.0):
    """development parameters of the context (sgpu_set_param), restored afterwards"""
    ctx.set_param(_sl.PARAM_CHUNK_LEN, chunk_len)
    ctx.set_param(_sl.PARAM_WARMUP, warmup)
    ctx.set_param(_sl.PARAM_THR_LONG, thr_long)
    try:
        yield ctx
    finally:
        ctx.set_param(_sl.PARAM_CHUNK_LEN, 0)
        ctx.set_param(_sl.PARAM_WARMUP, 0)
        ctx.set_param(_sl.PARAM_THR_LONG, 9.0)


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def ctx():
    c = sg.Context(device=0, max_samples=1 << 23, max_reads=4096)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_generic():
    c = sg.Context(device=0, max_samples=1 << 23, max_reads=4096, flags=sg.F_FORCE_GENERIC)
    yield c
    c.close()


@pytest.fixture(scope="module")
def sp1():
    return _fmt.load_npz(os.path.join(G, "sp1_dna.npz"))


@pytest.fixture(scope="module")
def rna():
    return _fmt.load_npz(os.path.join(G, "synth_rna.npz"))


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def check_against_oracle(orc, res, reads, rna_flag, want=ALL):
    for r, rd in enumerate(reads):
        if want & sg.WANT_EVENTS:
            st, ln, mn, sd = orc.events(*rd, rna=rna_flag)
            ev = res.events(r)
            assert np.array_equal(ev.start, st), f"read {r}: boundaries differ"
            assert np.array_equal(bits(ev.length), bits(ln)), f"read {r}"
            assert np.array_equal(bits(ev.mean), bits(mn)), f"read {r}: event_mean differs"
            assert np.array_equal(bits(ev.stdv), bits(sd)), f"read {r}: event_std differs"
        if want & sg.WANT_PA:
            assert np.array_equal(bits(res.pa[r]), bits(orc.pa(*rd))), f"read {r}: pA differs"
        if want & sg.WANT_STAT:
            assert np.array_equal(bits(res.stat[r]), bits(orc.stat(*rd))), f"read {r}: stat differs"


# ---- BASELINE config C1: sp1_dna.blow5 ------------------------------------------------------------
def test_sp1_compact_text_equals_reference_stdout(ctx, sp1):
    res = ctx.run([rd for _, rd in sp1], rna=0, want=sg.WANT_EVENTS)
    got = _fmt.EVENT_HDR_COMPACT
    for r, (rid, rd) in enumerate(sp1):
        ev = res.events(r)
        got += _fmt.event_compact(rid, len(rd[0]), ev.start, ev.length)
    assert got == open(os.path.join(G, "ref_sp1_event_c.txt")).read()
    assert int(res.ev_off[-1]) == 92935
    assert ctx.counters()["n_events"] == 92935


def test_sp1_event_dna_exp(ctx, sp1):
    """the reference's own golden file (scripts/test.sh:70-72); its header line is stale"""
    exp = open(os.path.join(G, "event_dna.exp")).read().split("\n", 1)[1]
    rid, rd = next(x for x in sp1 if x[0] == "05d90f17-f4a6-4349-924c-3ffd3457a99d")
    ev = sg.getevents(ctx, *rd)
    assert _fmt.event_long(rid, ev.start, ev.length, ev.mean, ev.stdv) == exp


def test_sp1_long_pa_stat_text(ctx, sp1):
    res = ctx.run([rd for _, rd in sp1], rna=0, want=ALL)
    got = _fmt.EVENT_HDR_LONG
    for r, (rid, _) in enumerate(sp1[:10]):
        ev = res.events(r)
        got += _fmt.event_long(rid, ev.start, ev.length, ev.mean, ev.stdv)
    assert got == gzip.open(os.path.join(G, "ref_sp1_event_first10.txt.gz"), "rt").read()
    got = _fmt.PA_HDR + "".join(_fmt.pa_line(rid, res.pa[r]) for r, (rid, _) in enumerate(sp1[:10]))
    assert got == gzip.open(os.path.join(G, "ref_sp1_pa_first10.txt.gz"), "rt").read()
    got = _fmt.STAT_HDR + "".join(_fmt.stat_line(rid, len(rd[0]), res.stat[r]) for r, (rid, rd) in enumerate(sp1))
    assert got == open(os.path.join(G, "ref_sp1_stat.txt")).read()


def test_sp1_all_reads_vs_oracle(ctx, orc, sp1):
    reads = [rd for _, rd in sp1]
    check_against_oracle(orc, ctx.run(reads, rna=0, want=ALL), reads, 0)


# ---- BASELINE config C2: RNA parameters, long form --------------------------------------------------
def test_rna_text_equals_reference_stdout(ctx, rna):
    res = ctx.run([rd for _, rd in rna], rna=1, want=sg.WANT_EVENTS | sg.WANT_STAT)
    got_l, got_c = _fmt.EVENT_HDR_LONG, _fmt.EVENT_HDR_COMPACT
    for r, (rid, rd) in enumerate(rna):
        ev = res.events(r)
        got_l += _fmt.event_long(rid, ev.start, ev.length, ev.mean, ev.stdv)
        got_c += _fmt.event_compact(rid, len(rd[0]), ev.start, ev.length)
    assert got_l == gzip.open(os.path.join(G, "ref_rna_event.txt.gz"), "rt").read()
    assert got_c == open(os.path.join(G, "ref_rna_event_c.txt")).read()
    got = _fmt.STAT_HDR + "".join(_fmt.stat_line(rid, len(rd[0]), res.stat[r]) for r, (rid, rd) in enumerate(rna))
    assert got == open(os.path.join(G, "ref_rna_stat.txt")).read()


# ---- seeded synthetic batches -----------------------------------------------------------------------
@pytest.mark.parametrize("rna_flag,seed,n,mean", [(0, 21, 64, 12000.0), (1, 22, 32, 20000.0), (0, 23, 300, 3000.0)])
def test_synthetic_batches_vs_oracle(ctx, orc, rna_flag, seed, n, mean):
    reads = synth.make_reads(n, mean=mean, seed=seed, rna=bool(rna_flag))
    res = ctx.run(reads, rna=rna_flag, want=ALL)
    check_against_oracle(orc, res, reads, rna_flag)
    assert res.fixups is not None and res.seq_order is not None


@pytest.mark.parametrize("rna_flag", [0, 1])
def test_fast_equals_sequential_order_path(ctx, ctx_generic, rna_flag):
    """the fast kernels and the reference-order kernels must agree bit for bit (on-device cross-check)"""
    reads = synth.make_reads(48, mean=30000.0, seed=77 + rna_flag, rna=bool(rna_flag))
    a = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS)
    b = ctx_generic.run(reads, rna=rna_flag, want=sg.WANT_EVENTS)
    assert np.array_equal(a.ev_off, b.ev_off)
    assert np.array_equal(a.ev_start, b.ev_start)
    assert np.array_equal(bits(a.ev_mean), bits(b.ev_mean))
    assert np.array_equal(bits(a.ev_stdv), bits(b.ev_stdv))
    assert np.all(b.seq_order == 1)


@pytest.mark.parametrize("rna_flag", [0, 1])
def test_long_reads_many_chunks(ctx, orc, rna_flag):
    """reads of hundreds of chunks: detector hand-over between chunks, events that span warp tiles of the emitter"""
    reads = [synth.make_read(900 + k, n, seed=4, p_change=0.025 if rna_flag else 0.1)
             for k, n in enumerate((1_000_003, 250_000, 2048, 777_777))]
    res = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS | sg.WANT_STAT)
    check_against_oracle(orc, res, reads, rna_flag, want=sg.WANT_EVENTS | sg.WANT_STAT)


# ---- edge cases ---------------------------------------------------------------------------------------
def test_ragged_and_tiny_reads(ctx, orc):
    """lengths around the window sizes and the alignment quantum; the reference aborts below 200 samples, the
    oracle (and the CUDA path) define those as in DESIGN.md: same arithmetic, single event if no peak"""
    rng = np.random.default_rng(5)
    lens = [1, 2, 5, 6, 7, 8, 9, 11, 12, 13, 15, 16, 17, 27, 28, 29, 31, 32, 33, 63, 64, 65, 199, 200, 201, 255, 256,
            257, 1023, 1024, 1025, 4095, 4096, 4097]
    reads = []
    for k, n in enumerate(lens):
        rd = synth.make_read(k, max(n, 8), seed=9)
        reads.append((rd[0][:n].copy(), rd[1], rd[2], rd[3]))
    rng.shuffle(reads)
    for rna_flag in (0, 1):
        res = ctx.run(reads, rna=rna_flag, want=ALL)
        check_against_oracle(orc, res, reads, rna_flag)


def test_many_tiny_reads(ctx, orc):
    """hundreds of reads of a few dozen samples: every read is a single (bounds-checked) chunk, dozens of reads per
    bitmap tile of the emitter"""
    rng = np.random.default_rng(12)
    base = synth.make_read(3, 40000, seed=8)
    reads, p = [], 0
    for k in range(700):
        n = int(rng.integers(1, 60))
        reads.append((base[0][p:p + n].copy(), base[1], float(k % 53), base[3]))
        p += n
    reads.insert(350, synth.make_read(5, 30000, seed=8))  # a normal read in the middle
    for rna_flag in (0, 1):
        res = ctx.run(reads, rna=rna_flag, want=ALL)
        check_against_oracle(orc, res, reads, rna_flag)


def test_constant_and_extreme_signals(ctx, orc):
    n = 5000
    ramp = (np.arange(n) % 4000 - 2000).astype(np.int16)
    reads = [
        (np.full(n, 500, dtype=np.int16), 8192.0, 10.0, 1400.0),          # zero variance everywhere
        (np.full(n, -32768, dtype=np.int16), 8192.0, 0.0, 1400.0),
        (np.where(np.arange(n) % 2, 32767, -32768).astype(np.int16), 8192.0, 3.0, 1400.0),
        (ramp, 8192.0, 7.0, 1400.0),
        (np.concatenate([np.full(2500, 400, np.int16), np.full(2500, 700, np.int16)]), 8192.0, 10.0, 1400.0),
        (synth.make_read(1, n)[0], 2048.0, -250.0, 748.5),                 # other digitisation / negative offset
        (synth.make_read(2, n)[0], 8192.0, 12.0, -1400.0),                 # negative range => negative raw_unit
    ]
    for rna_flag in (0, 1):
        res = ctx.run(reads, rna=rna_flag, want=ALL)
        check_against_oracle(orc, res, reads, rna_flag)


def test_flat_stretch_inside_a_read(ctx, orc):
    """SURVEY 7.3(2) adversarial case: a long constant stretch (t-stat exactly 0) after a sub-threshold bump"""
    rd = synth.make_read(11, 60000, seed=3)
    raw = rd[0].copy()
    raw[20000:45000] = raw[19999]
    reads = [(raw, rd[1], rd[2], rd[3])]
    for rna_flag in (0, 1):
        res = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS)
        check_against_oracle(orc, res, reads, rna_flag, want=sg.WANT_EVENTS)


@pytest.mark.parametrize("chunk_len", [128, 256, 512, 1024])
def test_chunk_lengths(ctx, orc, chunk_len):
    """the chunk length is chosen from the batch size; force every value the DNA path can take"""
    reads = synth.make_reads(24, mean=9000.0, seed=61)
    reads.append(synth.make_read(99, 3 * chunk_len, seed=61))       # exact multiple of the chunk length
    reads.append(synth.make_read(98, 2 * chunk_len + 1, seed=61))   # one-sample last chunk
    with tuned(ctx, chunk_len=chunk_len):
        res = ctx.run(reads, rna=0, want=ALL)
    check_against_oracle(orc, res, reads, 0)
    assert int(res.seq_order.sum()) == 0 and int(res.fixups.sum()) == 0


def test_short_warmup_falls_back(ctx, orc):
    """an 8-sample detector warm-up leaves many chunks in a wrong state: the boundary-state check must catch every
    one of them (fixups > 0), route those reads to the sequential-order kernels, and the results stay bit-exact"""
    reads = synth.make_reads(40, mean=12000.0, seed=62)
    with tuned(ctx, chunk_len=128, warmup=8):
        res = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
    check_against_oracle(orc, res, reads, 0, want=sg.WANT_EVENTS)
    assert int(res.fixups.sum()) > 0
    assert np.all(res.seq_order[res.fixups > 0] == 1)


@pytest.mark.parametrize("rna_flag", [0, 1])
def test_non_positive_pa_stays_on_the_fast_path(ctx, orc, rna_flag):
    """samples with zero or negative pA (raw <= -offset; about 3 reads in 100 of real R9.4 data): the walker takes
    the blocks around them through its exact path and the read stays on the fast path, bit-exact"""
    from test_host_walk import glitch_read
    rd = synth.make_read(7, 9000, seed=63)
    neg = rd[0].copy()
    neg[4000:4003] = -300                       # (raw + offset) < 0
    zero = rd[0].copy()
    zero[100] = -int(rd[2])                     # raw + offset == 0
    real = synth.make_read(9, 30000, seed=63)   # glitches like those of sp1_dna.blow5: raw + offset = -44, -376
    graw = real[0].copy()
    for k, pos in enumerate([0, 1, 7, 8, 1023, 1024, 1025, 5000, 5001, 5002, 12345, 29998, 29999]):
        graw[pos] = -int(real[2]) - (44, 376, 100, 0)[k % 4]
    reads = [rd, (neg, rd[1], rd[2], rd[3]), (zero, rd[1], rd[2], rd[3]), synth.make_read(8, 9000, seed=63),
             (graw, real[1], real[2], real[3]),
             # magnitudes of one unit next to full-scale ones: the sums are not exact any more, these two reads go to
             # the sequential-order kernels (and stay bit-exact)
             glitch_read(9000, 12), glitch_read(30000, 13, where=tuple(range(5, 30000, 997)))]
    res = ctx.run(reads, rna=rna_flag, want=ALL)
    check_against_oracle(orc, res, reads, rna_flag)
    assert list(res.seq_order[:5]) == [0] * 5 and int(res.fixups.sum()) == 0


def test_fractional_offset_with_non_positive_pa(ctx, orc):
    """a fractional offset makes the smallest nonzero |pA| a fraction of a unit; the blocks with low samples report
    their exact magnitudes to the witness, whatever the offset"""
    rd = synth.make_read(7, 9000, seed=64)
    neg = rd[0].copy()
    neg[4000] = -300
    reads = [(neg, rd[1], 7.5, rd[3]), (rd[0], rd[1], 7.5, rd[3])]
    res = ctx.run(reads, rna=0, want=ALL)
    check_against_oracle(orc, res, reads, 0)
    assert list(res.seq_order) == [0, 0]


def test_empty_batch_and_reuse(ctx):
    res = ctx.run([], rna=0, want=sg.WANT_EVENTS)
    assert res.n_reads == 0
    reads = synth.make_reads(3, mean=5000.0, seed=1)
    a = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
    b = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)  # same slot again: no stale state
    assert np.array_equal(a.ev_start, b.ev_start) and np.array_equal(a.ev_off, b.ev_off)


def test_two_slots_in_flight(ctx, orc):
    r0 = synth.make_reads(20, mean=8000.0, seed=31)
    r1 = synth.make_reads(20, mean=8000.0, seed=32, rna=True)
    ctx.fill(0, r0, 0)
    ctx.submit(0, sg.WANT_EVENTS)
    ctx.fill(1, r1, 1)
    ctx.submit(1, sg.WANT_EVENTS)
    check_against_oracle(orc, ctx.wait(0, sg.WANT_EVENTS), r0, 0, want=sg.WANT_EVENTS)
    check_against_oracle(orc, ctx.wait(1, sg.WANT_EVENTS), r1, 1, want=sg.WANT_EVENTS)


def test_errors(ctx):
    with pytest.raises(sg.SgpuError) as e:
        ctx.wait(0, sg.WANT_EVENTS)  # nothing submitted
    assert e.value.code == -6
    small = sg.Context(device=0, max_samples=4096, max_reads=4)
    with pytest.raises(sg.SgpuError) as e:
        small.fill(0, [synth.make_read(0, 5000)], 0)
    assert e.value.code == -5  # TOOBIG
    with pytest.raises(sg.SgpuError) as e:
        small.fill(0, [synth.make_read(i, 1500) for i in range(3)], 0)
    assert e.value.code == -4  # FULL
    small.close()


# ---- size-independent properties at larger sizes -------------------------------------------------------
def test_properties_large_batch(ctx):
    reads = synth.make_reads(150, mean=40000.0, seed=synth.SEED)
    res = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
    off = res.ev_off.astype(np.int64)
    assert off[0] == 0 and np.all(np.diff(off) >= 1)
    for r in range(res.n_reads):
        st = res.ev_start[off[r]:off[r + 1]].astype(np.int64)
        assert st[0] == 0 and np.all(np.diff(st) >= 3) and st[-1] < res.read_len[r]
    # concatenation invariance: every read alone gives the same events as in the batch
    for r in (0, 17, 149):
        ev = sg.getevents(ctx, *reads[r])
        assert np.array_equal(ev.start, res.events(r).start)
        assert np.array_equal(bits(ev.mean), bits(res.events(r).mean))
    # event means re-derived from pA on the host within 1e-4 relative (north_star tolerance)
    pa = sg.signal_in_picoamps(ctx, *reads[5]).astype(np.float64)
    ev = res.events(5)
    ends = np.append(ev.start[1:], len(pa)).astype(np.int64)
    cs = np.concatenate([[0.0], np.cumsum(pa)])
    mean = (cs[ends] - cs[ev.start.astype(np.int64)]) / (ends - ev.start.astype(np.int64))
    assert np.allclose(ev.mean, mean, rtol=1e-4)


def test_device_resident_path(ctx):
    torch = pytest.importorskip("torch")
    reads = synth.make_reads(40, mean=20000.0, seed=55)
    host = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
    lens = np.array([len(r[0]) for r in reads], dtype=np.int64)
    al = (lens + 7) // 8 * 8
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum(al)
    flat = np.zeros(int(off[-1]), dtype=np.int16)
    for r, rd in enumerate(reads):
        flat[off[r]:off[r] + lens[r]] = rd[0]
    dev = torch.device("cuda:0")
    d_s = torch.from_numpy(flat).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_len = torch.from_numpy(lens.astype(np.int32)).to(dev)
    d_o = torch.tensor([np.float32(r[2]) for r in reads], dtype=torch.float32, device=dev)
    d_u = torch.tensor([np.float32(r[3]) / np.float32(r[1]) for r in reads], dtype=torch.float32, device=dev)
    dctx = sg.Context(device=0, max_samples=int(off[-1]), max_reads=len(reads), flags=sg.F_NO_HOST_SLOTS)
    st = torch.cuda.current_stream().cuda_stream
    res = dctx.run_device(d_s.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_o.data_ptr(), d_u.data_ptr(),
                          len(reads), int(off[-1]), 0, sg.WANT_EVENTS, st)
    torch.cuda.synchronize()
    c = dctx.counters()
    assert c["status"] == 0 and c["n_events"] == int(host.ev_off[-1])
    ne = c["n_events"]
    h = dctx.d2h(res.ev_start, np.uint32, ne)  # raw device pointer -> host through the C-ABI
    assert np.array_equal(h, host.ev_start)
    assert np.array_equal(dctx.d2h(res.ev_off, np.uint64, len(reads) + 1), host.ev_off)
    assert np.array_equal(bits(dctx.d2h(res.ev_mean, np.float32, ne)), bits(host.ev_mean))
    dctx.close()


def test_full_size_batch(orc):
    """One batch of the size bench.py runs (16,384 synthetic DNA reads, ~650 M samples, generated on the device):
    size-independent properties over ALL events, and bit-exact parity with the oracle for a spread of reads."""
    torch = pytest.importorskip("torch")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    dev = torch.device("cuda:0")
    lens = synth.read_lengths(16384)
    p = bench.device_batch(torch, dev, lens, 0, 4242)
    dctx = sg.Context(device=0, max_samples=p["span"], max_reads=p["n_reads"], flags=sg.F_NO_HOST_SLOTS)
    st = torch.cuda.current_stream().cuda_stream
    res = dctx.run_device(p["samples"].data_ptr(), p["read_off"].data_ptr(), p["read_len"].data_ptr(),
                          p["offset"].data_ptr(), p["unit"].data_ptr(), p["n_reads"], p["span"], 0, sg.WANT_EVENTS, st)
    torch.cuda.synchronize()
    c = dctx.counters()
    assert c["status"] == 0 and c["n_seq_order_reads"] == 0 and c["n_fixups"] == 0
    ne, nr = c["n_events"], p["n_reads"]
    ev_off = dctx.d2h(res.ev_off, np.uint64, nr + 1).astype(np.int64)
    ev_start = dctx.d2h(res.ev_start, np.uint32, ne).astype(np.int64)
    ev_mean = dctx.d2h(res.ev_mean, np.float32, ne)
    ev_stdv = dctx.d2h(res.ev_stdv, np.float32, ne)
    assert ev_off[0] == 0 and ev_off[-1] == ne and np.all(np.diff(ev_off) >= 1)
    first = np.zeros(ne, dtype=bool)
    first[ev_off[:-1]] = True
    assert np.all(ev_start[first] == 0)                       # event 0 of every read starts at sample 0
    gaps = np.diff(ev_start)[~first[1:]]
    assert gaps.min() >= 3                                    # peaks are at least floor(w1/2)+2 = 3 samples apart
    last = ev_start[ev_off[1:] - 1]
    assert np.all(last < p["host_len"])
    assert 0.15 < ne / p["n_samples"] < 0.25                  # the signal model gives ~0.19 events per sample
    assert np.all(np.isfinite(ev_mean)) and np.all(ev_stdv >= 0)
    assert 50.0 < float(ev_mean.mean()) < 130.0
    # bit-exact against the oracle for reads spread over the batch (first, last, and in between)
    off, hl = p["host_off"], p["host_len"]
    for r in [0, 1, nr // 3, nr // 2, nr - 2, nr - 1, int(np.argmax(hl)), int(np.argmin(hl))]:
        raw = p["samples"][int(off[r]): int(off[r]) + int(hl[r])].cpu().numpy()
        s0, ln, mn, sd = orc.events(raw, synth.DIGITISATION, float(p["host_offset"][r]), synth.RANGE, rna=0)
        a, b = ev_off[r], ev_off[r + 1]
        assert np.array_equal(ev_start[a:b], s0.astype(np.int64)), f"read {r}: boundaries differ"
        assert np.array_equal(bits(ev_mean[a:b]), bits(mn)) and np.array_equal(bits(ev_stdv[a:b]), bits(sd))
    dctx.close()


@pytest.mark.parametrize("rna_flag", [0, 1])
def test_plateaus_emit_old_peaks_on_the_fast_path(ctx, orc, rna_flag):
    """linear ramps keep the t-statistic within peak_height of its maximum: the peak is emitted dozens of samples
    after its position, too old for the walker's per-block register mask; the block then records its peaks one by
    one (redo_block) and the read stays on the fast path"""
    from test_host_walk import ramp_read
    reads = [ramp_read(12000, seed=3), ramp_read(15000, seed=4)] + synth.make_reads(6, mean=9000.0, seed=63)
    with tuned(ctx, chunk_len=4096):
        res = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS)
    check_against_oracle(orc, res, reads, rna_flag, want=sg.WANT_EVENTS)
    assert int(res.seq_order.sum()) == 0 and int(res.fixups.sum()) == 0


def _fuzz_read(rng, k):
    """a synthetic read with random adversarial features: glitches (zero / negative / barely positive pA), flat
    stretches, ramps (plateaus of the t-statistic), steps, saturation"""
    n = int(rng.choice([1, 2, 5, 13, 31, 200, 257, 1000, 1023, 1024, 1025, 3000, 6000, 12000]))
    n = max(1, n + int(rng.integers(-3, 4)))
    rd = synth.make_read(1000 + k, n, seed=int(rng.integers(1, 1 << 30)))
    raw = rd[0].astype(np.int64)
    off = int(rd[2])
    for _ in range(int(rng.integers(0, 6))):
        kind = int(rng.integers(0, 6))
        p = int(rng.integers(0, n))
        ln = int(rng.integers(1, max(2, min(n - p, 400))))
        if kind == 0:      # single glitches like real data: raw + offset = -44, -376, ...
            raw[p] = -off - int(rng.choice([0, 1, 44, 376, 3000]))
        elif kind == 1:    # a run of low values around zero
            raw[p:p + ln] = -off + rng.integers(-5, 70, min(ln, n - p))
        elif kind == 2:    # flat stretch
            raw[p:p + ln] = raw[p]
        elif kind == 3:    # ramp
            step = int(rng.choice([-7, -2, 3, 11, 30]))
            raw[p:p + ln] = raw[p] + step * np.arange(min(ln, n - p))
        elif kind == 4:    # big step
            raw[p:] += int(rng.choice([-150, 200, 400]))
        else:              # saturation
            raw[p:p + ln] = int(rng.choice([32767, 4000]))
    return np.clip(raw, -32768, 32767).astype(np.int16), rd[1], rd[2], rd[3]


@pytest.mark.parametrize("rna_flag,chunk_len,seed", [(0, 128, 1), (0, 1024, 2), (1, 512, 3), (1, 4096, 4), (0, 4096, 5)])
def test_fuzz_adversarial_reads(ctx, orc, rna_flag, chunk_len, seed):
    """random reads with glitches, flat stretches, ramps, steps and saturation: every output bit-exact, whichever
    path (walker, its exact recompute, sequential-order kernels) a read takes"""
    rng = np.random.default_rng(seed)
    reads = [_fuzz_read(rng, k) for k in range(300)]
    with tuned(ctx, chunk_len=chunk_len):
        res = ctx.run(reads, rna=rna_flag, want=ALL)
        check_against_oracle(orc, res, reads, rna_flag)
        # the same reads as svb-zd streams decoded on the GPU
        sres = ctx.run_svbzd([(orc.svbzd_encode(r[0]), r[1], r[2], r[3]) for r in reads], rna=rna_flag, want=ALL)
    assert np.array_equal(sres.ev_off, res.ev_off) and np.array_equal(sres.ev_start, res.ev_start)
    assert np.array_equal(bits(sres.ev_mean), bits(res.ev_mean)) and np.array_equal(bits(sres.ev_stdv), bits(res.ev_stdv))
    assert np.array_equal(bits(sres.stat), bits(res.stat))


# ---- round 2: the long detector's lives (not stepped by the walker, replayed as jobs) ----------------------------
@pytest.mark.parametrize("rna_flag,chunk_len", [(0, 128), (0, 1024), (0, 480), (1, 512), (1, 4096), (1, 736)])
@pytest.mark.parametrize("thr", [0.3, 1.0, 2.5])
def test_long_detector_lives_are_replayed(ctx, orc, rna_flag, chunk_len, thr):
    """with the reference's threshold (9.0) the long detector never emits on these reads; lowered through the
    context's test parameter it emits all the time: every life that may emit is replayed by long_jobs_kernel
    (lives inside a chunk, lives inherited from an earlier chunk, lives that run past their chunk)"""
    from test_host_walk import ramp_read
    reads = synth.make_reads(12, mean=9000.0, seed=171 + rna_flag, rna=bool(rna_flag))
    rd = synth.make_read(11, 30000, seed=3)
    raw = rd[0].copy()
    raw[6000:14000] = raw[5999]                              # no short peak for thousands of samples: one long life
    raw[14000:20000] = raw[5999] + (np.arange(6000) // 40)
    reads += [(raw, rd[1], rd[2], rd[3]), ramp_read(12000, seed=3)]
    with tuned(ctx, chunk_len=chunk_len, thr_long=thr):
        res = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS)
        jobs = ctx.counters()["n_long_jobs"]
    extra = 0
    for r, rd in enumerate(reads):
        if res.seq_order[r]:
            continue  # (the sequential-order kernels keep the reference's constant)
        st = orc.event_starts(*rd, rna=rna_flag, thr_long=thr)
        assert np.array_equal(res.events(r).start.astype(np.int64), st), f"read {r}"
        extra += len(np.setdiff1d(st, orc.events(*rd, rna=rna_flag)[0].astype(np.int64)))
    assert int(res.seq_order.sum()) <= 2 and jobs > 0
    if thr <= 1.0:
        assert extra > 0


def test_every_hot_life_is_replayed(ctx, orc):
    """at the reference's threshold the long detector never emits, so parity cannot tell whether its lives are looked
    at: the number of replayed lives must reach the number of lives in which a stepped position has t2 > 9 (oracle,
    stepped position by position). (A build that silently dropped the job bookkeeping passed every parity test.)"""
    reads = [synth.make_read(k, 12000, seed=77) for k in range(3)]
    for thr in (9.0, 6.0):
        need = sum(orc.hot_lives(*rd, rna=0, thr_long=thr) for rd in reads)
        with tuned(ctx, chunk_len=1024, thr_long=thr):
            res = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
            jobs = ctx.counters()["n_long_jobs"]
        assert need > 0 and jobs >= need and int(res.seq_order.sum()) == 0, (thr, need, jobs)
        assert jobs <= 60 * need + 60


def test_far_from_the_pivot(ctx, orc):
    """samples more than 1023 raw units from a chunk's pivot (level jumps, spikes): every window that holds one is a
    candidate of the long detector's test; results stay bit-exact"""
    rd = synth.make_read(21, 40000, seed=8)
    raw = rd[0].astype(np.int32)
    raw[3000:5000] += 3000
    raw[7000] = 30000
    raw[7001] = -20000
    raw[9000:9100] += np.arange(100) * 40
    raw[20000:] += 1500
    raw = np.clip(raw, -32768, 32767).astype(np.int16)
    reads = [(raw, rd[1], rd[2], rd[3])] + synth.make_reads(4, mean=9000.0, seed=8)
    for rna_flag in (0, 1):
        res = ctx.run(reads, rna=rna_flag, want=ALL)
        check_against_oracle(orc, res, reads, rna_flag)
        assert ctx.counters()["n_long_jobs"] > 0


# ---- BASELINE configs[4]: 2,000,000-sample reads (the reference's width limits: int32_t nsample misc.c:20,
# int peak_pos events.c:277) --------------------------------------------------------------------------------------
@pytest.mark.parametrize("rna_flag", [0, 1])
def test_two_million_sample_reads(ctx, orc, rna_flag):
    reads = [synth.make_read(7000 + k, 2_000_000 + 3 * k, seed=9, p_change=0.025 if rna_flag else 0.1) for k in range(2)]
    reads.append(synth.make_read(7002, 50_001, seed=9))
    res = ctx.run(reads, rna=rna_flag, want=sg.WANT_EVENTS | sg.WANT_STAT)
    check_against_oracle(orc, res, reads, rna_flag, want=sg.WANT_EVENTS | sg.WANT_STAT)
    assert int(res.seq_order.sum()) == 0 and int(res.fixups.sum()) == 0


def test_empty_records_get_defined_values(ctx):
    """a batch of nothing but empty records after a batch with data, and an empty record inside a batch: no stale
    values from the previous batch in any per-read output"""
    reads = synth.make_reads(3, mean=5000.0, seed=90)
    ctx.run(reads, rna=0, want=ALL | sg.WANT_ENT | sg.WANT_JNN | sg.WANT_PREFIX)
    empty = (np.zeros(0, np.int16), 8192.0, 3.0, 1402.882324)
    res = ctx.run([empty, empty], rna=0, want=ALL | sg.WANT_ENT | sg.WANT_JNN | sg.WANT_PREFIX)
    assert np.array_equal(res.ev_off, np.zeros(3, np.uint64)) and res.events(0).n == 0
    assert not res.stat[:, 4:].any() and not res.ent.any() and all(len(j) == 0 for j in res.jnn)
    assert np.array_equal(res.prefix_pos, -np.ones((2, 4), np.int32))
    res = ctx.run([reads[0], empty, reads[1]], rna=0, want=sg.WANT_EVENTS | sg.WANT_STAT)
    assert res.events(1).n == 0 and res.stat[1, 4] == 0.0 and res.stat[1, 5] == 0.0
    assert res.events(0).n > 1 and res.events(2).n > 1


def test_rna_reads_up_to_270k_samples(ctx, orc):
    """BASELINE configs[1] is the reference's RNA file (absent from the checkout: a missing large blob); its golden
    output test/event_rna.exp lists 100 reads of 14,509..269,194 samples. Twenty seeded RNA-like reads over that range
    of lengths (long dwells, the RNA detector parameters), every boundary / mean / stdv / stat bit-exact."""
    rng = np.random.default_rng(2024)
    lens = np.concatenate([[14509, 269194, 200001, 65536, 131072 + 5], rng.integers(14509, 269194, 15)])
    reads = [synth.make_read(5000 + k, int(n), seed=31, p_change=0.025, noise=float(rng.choice([1.0, 2.0, 3.0])))
             for k, n in enumerate(lens)]
    for lo in range(0, len(reads), 5):   # (the test context holds 8 M samples per batch)
        part = reads[lo:lo + 5]
        res = ctx.run(part, rna=1, want=sg.WANT_EVENTS | sg.WANT_STAT)
        check_against_oracle(orc, res, part, 1, want=sg.WANT_EVENTS | sg.WANT_STAT)
        assert int(res.seq_order.sum()) == 0 and int(res.fixups.sum()) == 0


@pytest.mark.parametrize("cta_min", [0, 5000, 1 << 31])
def test_stat_moments_by_warp_and_by_cta(ctx, orc, cta_min):
    """stat.cu adds a read's floats up in the reference's order (stat.h:17-44) with one warp per read or, for long
    reads, one CTA per read (several warps summarise consecutive superblocks in the accumulator's binade). Both
    kernels on the same reads -- tiny ones, lengths around the superblock (1,024) and round (8,192) sizes, negative
    samples (the value-by-value path), a constant read, zeros -- give the oracle's bits."""
    rng = np.random.default_rng(77)
    lens = [1, 2, 31, 32, 33, 1023, 1024, 1025, 8191, 8192, 8193, 16385, 40000, 100001, 3000, 262144 + 7]
    reads = [synth.make_read(300 + k, n, seed=5) for k, n in enumerate(lens)]
    neg = synth.make_read(900, 30000, seed=5)
    reads.append(((neg[0].astype(np.int32) - 520).astype(np.int16), neg[1], 13.0, neg[3]))       # samples of both signs
    reads.append((np.full(20000, 487, np.int16), 8192.0, 4.0, 1402.882324))                      # stdv exactly 0
    reads.append((np.zeros(9000, np.int16), 8192.0, 0.0, 1402.882324))                           # sums stay 0
    reads.append(((rng.integers(-32768, 32767, 50000)).astype(np.int16), 2048.0, -200.0, 748.58))  # full range
    ctx.set_param(_sl.PARAM_STAT_CTA_MIN, cta_min)
    try:
        res = ctx.run(reads, rna=0, want=sg.WANT_STAT | sg.WANT_JNN)
    finally:
        ctx.set_param(_sl.PARAM_STAT_CTA_MIN, 131072)
    for r, rd in enumerate(reads):
        assert np.array_equal(bits(res.stat[r]), bits(orc.stat(*rd))), f"read {r} (n={len(rd[0])}): stat differs"
        assert np.array_equal(res.jnn[r], orc.jnn(rd[0], 0)), f"read {r} (n={len(rd[0])}): jnn differs"
