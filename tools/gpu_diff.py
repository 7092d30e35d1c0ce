#!/usr/bin/env python
"""Debug helper (GPU box): run a batch through the CUDA path and print, per read, where it first differs from the
CPU oracle (test infrastructure). usage: python tools/gpu_diff.py sp1|rna|synth[:n:mean:seed:rna] [generic]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _fmt  # noqa: E402
from _oracle import Oracle  # noqa: E402
import sigtk_b200 as sg  # noqa: E402
from sigtk_b200 import synth  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "sp1"
    flags = sg.F_FORCE_GENERIC if "generic" in sys.argv[2:] else 0
    rna = 0
    if what == "sp1":
        reads = [rd for _, rd in _fmt.load_npz(os.path.join(ROOT, "tests/golden/sp1_dna.npz"))]
    elif what == "rna":
        reads = [rd for _, rd in _fmt.load_npz(os.path.join(ROOT, "tests/golden/synth_rna.npz"))]
        rna = 1
    else:
        parts = what.split(":")
        n, mean, seed, rna = int(parts[1]), float(parts[2]), int(parts[3]), int(parts[4])
        reads = synth.make_reads(n, mean=mean, seed=seed, rna=bool(rna))
    orc = Oracle()
    tot = sum(len(r[0]) for r in reads)
    with sg.Context(0, max_samples=max(1 << 22, 2 * tot), max_reads=max(4096, len(reads)), flags=flags) as ctx:
        res = ctx.run(reads, rna=rna, want=sg.WANT_EVENTS)
        print("counters", ctx.counters(), "seq_order reads:", int(res.seq_order.sum()), "fixups:", int(res.fixups.sum()))
        bad = 0
        off = 0
        for r, rd in enumerate(reads):
            st, ln, mn, sd = orc.events(*rd, rna=rna)
            ev = res.events(r)
            same_b = len(st) == ev.n and np.array_equal(st, ev.start)
            same_m = same_b and np.array_equal(mn.view(np.uint32), ev.mean.view(np.uint32)) and np.array_equal(
                sd.view(np.uint32), ev.stdv.view(np.uint32))
            if not same_m:
                bad += 1
                if bad <= 12:
                    k = 0
                    m = min(len(st), ev.n)
                    while k < m and st[k] == ev.start[k]:
                        k += 1
                    km = 0
                    while km < m and (mn[km:km + 1].view(np.uint32) == ev.mean[km:km + 1].view(np.uint32)).all() and (
                            sd[km:km + 1].view(np.uint32) == ev.stdv[km:km + 1].view(np.uint32)).all():
                        km += 1
                    print(f"read {r} flat_off~{off} n={len(rd[0])} ev oracle={len(st)} gpu={ev.n} first boundary diff @ev {k}"
                          f" (oracle {st[k:k+4]} gpu {ev.start[k:k+4]}), first stat diff @ev {km} seq={res.seq_order[r]}")
                    if km < m:
                        print("   oracle", st[km], mn[km], sd[km], " gpu", ev.start[km], ev.mean[km], ev.stdv[km])
            off += (len(rd[0]) + 7) // 8 * 8
        print(f"{bad} of {len(reads)} reads differ")


if __name__ == "__main__":
    main()
