#!/usr/bin/env python
"""Device time of the `stat` and `pa` kernel groups on 2,000 synthetic reads (stage timers of the C-ABI)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sigtk_b200 as sg
from sigtk_b200 import synth
reads = synth.make_reads(2000, mean=40000.0, seed=7)
n = sum(len(r[0]) for r in reads)
with sg.Context(device=0, max_samples=n + 8 * len(reads) + 64, max_reads=len(reads), flags=sg.F_STAGE_TIMERS) as ctx:
    for want, name in ((sg.WANT_STAT, "stat"), (sg.WANT_PA, "pa"), (sg.WANT_EVENTS | sg.WANT_PA | sg.WANT_STAT, "all")):
        for _ in range(2):
            ctx.run(reads, rna=0, want=want)
        st = {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}
        print(json.dumps({"want": name, "samples": n, "stage_ms": st}))
