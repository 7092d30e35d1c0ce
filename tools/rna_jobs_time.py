"""Stage times of the event path on synthetic RNA reads for three batch sizes: the long detector's replay (long_jobs) is
as long as its longest life, whatever the batch."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import sigtk_b200 as sg
from sigtk_b200 import synth
lens = synth.read_lengths(1500, 40000.0, 0.6, synth.SEED)
for nr in (100, 400, 1500):
    reads = [synth.make_read_cb(i, int(lens[i]), seed=synth.SEED + 13, p_change=0.025) for i in range(nr)]
    n = sum(len(r[0]) for r in reads)
    with sg.Context(device=0, max_samples=n + 8 * nr + 64, max_reads=nr, flags=sg.F_STAGE_TIMERS) as ctx:
        for _ in range(3):
            res = ctx.run(reads, rna=1, want=sg.WANT_EVENTS)
        st = {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}
        print(json.dumps({"reads": nr, "samples": n, "jobs": ctx.counters()["n_long_jobs"], "stage_ms": st}), flush=True)
