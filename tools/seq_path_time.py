import sys
sys.path.insert(0, ".")
import sigtk_b200 as sg
from sigtk_b200 import synth
reads = synth.make_reads(200, mean=40000.0, seed=7)
ctx = sg.Context(device=0, max_samples=sum(len(r[0]) for r in reads) + 8 * 200 + 64, max_reads=200, flags=sg.F_FORCE_GENERIC)
for _ in range(2): res = ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
print(max(len(r[0]) for r in reads), sum(len(r[0]) for r in reads))
