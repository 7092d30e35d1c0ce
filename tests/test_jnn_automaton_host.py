"""CPU check of the reformulation jnn_walk_kernel (sigtk_b200/csrc/jnn.cu) rests on: the reference's segmenter
(jnn.c:176-266) as a 7-state automaton over the in-band bit, evaluated per 32-sample word as a map of states, composed
by a prefix scan over the 32 words of a 1024-sample block, with at most one candidate per word (the close of a stretch
that entered the word open). This is a plain-Python restatement of the kernel's data flow (same carries: state,
opening position, previous word), compared with the oracle, which keeps the reference's counters (c, w, err, prev_err)
and steps sample by sample. Test infrastructure only."""
import math
import os

import numpy as np
import pytest

import _fmt
from _oracle import Oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLOSED = 6


def _ffs(x):
    return (x & -x).bit_length()


def _clz(x):
    return 32 - x.bit_length()


def _popc(x):
    return bin(x).count("1")


def word_end(bits, zeros, m, s):
    """jnn_word_end: state after the m samples of a word entered in state s"""
    j = 0
    while j < m:
        if s == CLOSED:
            rest = bits >> j
            if rest == 0:
                return CLOSED
            j += _ffs(rest) - 1
            s = 0
        z = zeros >> j
        need = CLOSED - s
        nz = _popc(z)
        if nz < need:
            return s + nz
        for _ in range(1, need):
            z &= z - 1
        j += _ffs(z)
        s = CLOSED
    return s


def band_of(raw):
    """mean -/+ 0.75 stdv of the clamped signal with float32 sums in sample order (jnn.c:181-185), as integers"""
    sig = np.clip(raw, 0, 1200).astype(np.float32)
    acc = np.float32(0)
    for v in sig:
        acc = np.float32(acc + v)
    mn = np.float32(acc / np.float32(len(sig)))
    dev = np.float32(0)
    for v in sig:
        d = np.float32(v - mn)
        dev = np.float32(dev + np.float32(d * d))
    sd = np.float32(np.sqrt(np.float32(dev / np.float32(len(sig)))))
    band = np.float32(sd * np.float32(0.75))
    top, bot = np.float32(mn + band), np.float32(mn - band)
    hi_i = int(math.ceil(min(max(float(top), -1.0), 2000.0))) - 1
    lo_i = int(math.floor(min(max(float(bot), -2.0), 2000.0))) + 1
    return lo_i, hi_i


def walk(raw, rna):
    n = len(raw)
    lo_i, hi_i = band_of(raw)
    window = 1000 if rna else 150
    first_min = np.float32(window) * np.float32(1.0 if rna else 0.25)
    cand_min = min(np.float32(window), first_min)
    sv = np.clip(raw.astype(np.int64), 0, 1200)
    inb = (sv >= lo_i) & (sv <= hi_i)
    n_seg = last_y = 0
    out = []
    state, open_pos, last_word = CLOSED, 0, 0xFFFFFFFF
    for t0 in range(0, n, 1024):
        ins, zs, ms, maps = [], [], [], []
        for lane in range(32):
            i0 = t0 + lane * 32
            m = max(0, min(32, n - i0))
            bits = 0
            for j in range(m):
                if inb[i0 + j]:
                    bits |= 1 << j
            valid = 0xFFFFFFFF if m == 32 else (1 << m) - 1
            ins.append(bits)
            zs.append(~bits & valid)
            ms.append(m)
            maps.append([word_end(bits, zs[-1], m, s) for s in range(7)])
        s_in, cur = [], state
        for lane in range(32):  # the prefix scan of the maps
            s_in.append(cur)
            cur = maps[lane][cur]
        close, perr, last_open = [-1] * 32, [0] * 32, [-1] * 32
        for lane in range(32):
            i0, bits, zeros, m = t0 + lane * 32, ins[lane], zs[lane], ms[lane]
            prev_in = ins[lane - 1] if lane else last_word
            j, s = 0, s_in[lane]
            while j < m:
                if s == CLOSED:
                    rest = bits >> j
                    if rest == 0:
                        break
                    j += _ffs(rest) - 1
                    s = 0
                    last_open[lane] = i0 + j
                z = zeros >> j
                need = CLOSED - s
                if _popc(z) < need:
                    break
                for _ in range(1, need):
                    z &= z - 1
                pos = j + _ffs(z) - 1
                if j == 0 and s_in[lane] != CLOSED and close[lane] < 0:
                    close[lane] = pos
                    below = (bits & ((1 << pos) - 1)) if pos else 0
                    perr[lane] = (pos - 1) - (31 - _clz(below)) if below else pos + _clz(prev_in)
                j, s = pos + 1, CLOSED
        run = open_pos
        for lane in range(32):
            if close[lane] >= 0:
                i0 = t0 + lane * 32
                c, en = i0 + close[lane] - run, i0 + close[lane] - perr[lane]
                if np.float32(c) >= cand_min and (c >= window or (n_seg == 0 and np.float32(c) >= first_min)):
                    if n_seg and run - last_y < 50:
                        out[-1][1] = en
                    else:
                        out.append([run, en])
                        n_seg += 1
                    last_y = en
            run = max(run, last_open[lane])
        open_pos, state, last_word = run, cur, ins[31]
    return np.array(out, dtype=np.int64).reshape(-1, 2)


@pytest.mark.parametrize("npz,rna_flag,take", [("jnn_stalls_dna.npz", 0, 16), ("jnn_stalls_rna.npz", 1, 16),
                                               ("sp1_dna.npz", 0, 25)])
def test_automaton_scan_equals_the_counter_machine(npz, rna_flag, take):
    orc = Oracle()
    n_seg = 0
    for _, rd in _fmt.load_npz(os.path.join(G, npz))[:take]:
        if len(rd[0]) == 0:
            continue
        exp = orc.jnn(rd[0], rna_flag)
        assert np.array_equal(walk(rd[0], rna_flag), exp)
        n_seg += len(exp)
    assert n_seg > 0 or npz == "sp1_dna.npz"
