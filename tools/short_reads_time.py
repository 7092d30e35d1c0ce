#!/usr/bin/env python
"""Kernel time on short reads (sp1_dna.blow5-like lengths: mean 5,000 samples), where the first / last chunks of a
read (the bounds-checked walker instantiation) are a large share of the work."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sigtk_b200 as sg
from sigtk_b200 import synth
for mean, nreads in ((5000.0, 12000), (40000.0, 1500)):
    reads = synth.make_reads(nreads, mean=mean, seed=7, lo=1000)
    n = sum(len(r[0]) for r in reads)
    with sg.Context(device=0, max_samples=n + 8 * len(reads) + 64, max_reads=len(reads), flags=sg.F_STAGE_TIMERS) as ctx:
        for _ in range(3):
            ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
        st = {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}
        tot = sum(st.values())
        print(json.dumps({"mean_len": mean, "reads": nreads, "samples": n, "stage_ms": st, "gsamples_per_s": round(n / tot / 1e6, 1)}))
