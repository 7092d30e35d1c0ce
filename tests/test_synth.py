"""The counter-based synthetic set: every read is a pure function of (seed, read index), and the
device generator of bench.py (torch) produces the very samples of sigtk_b200.synth.make_read_cb (numpy), whatever the
batch boundaries -- so every rank and both bench arms draw from the same named 1,000,000-read set."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sigtk_b200 import synth  # noqa: E402


def test_device_generator_equals_numpy_generator():
    lens = np.array([2000, 2001, 7, 4096, 12345, 1, 8, 30000], dtype=np.int64)
    for first in (0, 5, 999_990):
        for p in (0.1, 0.025):
            b = bench.device_batch(torch, torch.device("cpu"), lens, first, synth.SEED, p)
            flat = b["samples"].numpy()
            for k, n in enumerate(lens):
                want = synth.make_read_cb(first + k, int(n), p_change=p)
                a = int(b["host_off"][k])
                assert np.array_equal(flat[a:a + n], want[0]), (first, k)
                assert not flat[a + n:int(b["host_off"][k + 1])].any()          # padding
                assert float(b["host_offset"][k]) == want[2]


def test_reads_do_not_depend_on_the_batch_they_are_in():
    lens = synth.read_lengths(64)[:12]
    whole = bench.device_batch(torch, torch.device("cpu"), lens, 100, synth.SEED)
    part = bench.device_batch(torch, torch.device("cpu"), lens[5:9], 105, synth.SEED)
    a, b = int(whole["host_off"][5]), int(whole["host_off"][9])
    assert np.array_equal(whole["samples"].numpy()[a:b], part["samples"].numpy())


def test_signal_model_statistics():
    """levels U(60,120) pA changing with probability 0.1 per sample, noise of 2 pA standard deviation"""
    raw, dig, off, rng = synth.make_read_cb(3, 200_000)
    pa = (raw.astype(np.float64) + off) * rng / dig
    assert 85.0 < pa.mean() < 95.0
    d = np.diff(pa)
    quiet = np.abs(d) < 6.0                      # inside a level: difference of two noise samples, variance 2 * 4
    assert 2.6 < d[quiet].std() < 3.0
    assert 0.07 < 1.0 - quiet.mean() < 0.11      # level changes (those that move the level by more than the noise)
