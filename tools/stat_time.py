#!/usr/bin/env python
"""Device time of the `stat` and `pa` kernel groups (stage timers of the C-ABI) on 2,000 synthetic reads of the bench
distribution and on 2,000,000-sample reads, for several values of SGPU_PARAM_STAT_CTA_MIN (the read length from which
the moments kernels give a read a CTA instead of a warp)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sigtk_b200 as sg
from sigtk_b200 import synth, _lib

sets = {
    "2000 reads, lognormal mean 40k": synth.make_reads(2000, mean=40000.0, seed=7),
    "32 reads of 2,000,000 samples": [synth.make_read_cb(k, 2_000_000) for k in range(32)],
}
sets["160 reads of 2,000,000 samples (the 32, five times)"] = sets["32 reads of 2,000,000 samples"] * 5
for name, reads in sets.items():
    n = sum(len(r[0]) for r in reads)
    with sg.Context(device=0, max_samples=n + 8 * len(reads) + 64, max_reads=len(reads), flags=sg.F_STAGE_TIMERS) as ctx:
        for cta_min in (0, 8192, 16384, 32768, 65536, 131072, 1 << 31):
            ctx.set_param(_lib.PARAM_STAT_CTA_MIN, cta_min)
            for want, wname in ((sg.WANT_STAT, "stat"), (sg.WANT_JNN, "jnn")):
                for _ in range(3):
                    ctx.run(reads, rna=0, want=want)
                st = {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}
                tot = sum(st.values())
                print(json.dumps({"set": name, "cta_min": cta_min, "want": wname, "samples": n, "stage_ms": st,
                                  "gsamples_s": round(n / tot / 1e6, 1)}), flush=True)
        ctx.set_param(_lib.PARAM_STAT_CTA_MIN, 131072)
        for _ in range(2):
            ctx.run(reads, rna=0, want=sg.WANT_PA)
        print(json.dumps({"set": name, "want": "pa", "samples": n, "stage_ms": {k: round(ms, 3) for k, ms, _ in ctx.stage_times()}}))
