import sys, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import sigtk_b200 as sg
from sigtk_b200 import synth, _lib
from _oracle import Oracle
orc=Oracle()
reads=[synth.make_read(k, 40000) for k in range(20)]
with sg.Context(device=0, max_samples=1<<21, max_reads=64) as ctx:
    for thr in (9.0, 2.0, 1.0):
        ctx.set_param(_lib.PARAM_THR_LONG, thr)
        res=ctx.run(reads, rna=0, want=sg.WANT_EVENTS)
        c=ctx.counters()
        bad=0; miss=0; extra=0
        for r,rd in enumerate(reads):
            st=orc.event_starts(*rd, rna=0, thr_long=thr)
            got=res.events(r).start.astype(np.int64)
            if not np.array_equal(got,st): bad+=1; miss+=len(np.setdiff1d(st,got)); extra+=len(np.setdiff1d(got,st))
        print("thr",thr,"gpu jobs",c['n_long_jobs'],"seq",c['n_seq_order_reads'],"reads differing",bad,"missing",miss,"extra",extra)
