#!/usr/bin/env python
"""Cost of the sequential-order path on the GPU box: the same 2,000 synthetic reads (a) as they are, (b) with one
non-positive pA sample in every 33rd read (about the rate seen in sp1_dna.blow5: 3 of 100 reads), which routes those
reads to the sequential-order kernels. Prints the stage times of both runs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sigtk_b200 as sg  # noqa: E402
from sigtk_b200 import synth  # noqa: E402

reads = synth.make_reads(2000, mean=40000.0, seed=7)
bad = [(r[0].copy(), r[1], r[2], r[3]) for r in reads]
for k in range(0, len(bad), 33):
    bad[k][0][len(bad[k][0]) // 2] = -100
n = sum(len(r[0]) for r in reads)
with sg.Context(device=0, max_samples=n + 8 * len(reads) + 64, max_reads=len(reads), flags=sg.F_STAGE_TIMERS) as ctx:
    for name, rs in (("clean", reads), ("3pct_nonpositive", bad)):
        for _ in range(2):
            res = ctx.run(rs, rna=0, want=sg.WANT_EVENTS)
        st = {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}
        print(json.dumps({"case": name, "samples": n, "seq_order_reads": int(res.seq_order.sum()), "stage_ms": st,
                          "kernel_ms_total": round(sum(st.values()), 3)}))
