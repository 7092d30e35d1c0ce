#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) as markdown: duration, DRAM bytes, pipe utilisation, stall reasons and
the hottest source lines. usage: python tools/ncu_summary.py report.ncu-rep samples_in_launch > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep, n_samples = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def g(name, default="n/a"):
    return m.get(name, ("", default))


def f(name):
    try:
        return float(g(name)[1])
    except ValueError:
        return float("nan")


print(f"# ncu summary: {g('Kernel Name')[1]}\n")
print(f"report `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`), one launch over {n_samples:.0f} samples\n")
dur = f("gpu__time_duration.sum")
unit = g("gpu__time_duration.sum")[0]
dur_ms = dur * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit, 1.0)
rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
sc = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd *= sc.get(g("dram__bytes_read.sum")[0], 1.0)
wr *= sc.get(g("dram__bytes_write.sum")[0], 1.0)
inst = f("smsp__inst_executed.sum")
print("| metric | value |\n|---|---|")
print(f"| duration (under ncu, cold cache) | {dur_ms:.3f} ms  ({n_samples / dur_ms / 1e6:.1f} Gsamples/s) |")
print(f"| dram bytes read / written | {rd / 1e6:.1f} MB / {wr / 1e6:.1f} MB  = {(rd + wr) / n_samples:.2f} B/sample |")
print(f"| dram throughput | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[1]} % of peak |")
print(f"| warp instructions executed | {inst:.3e}  = {inst / n_samples:.2f} per sample |")
print(f"| issue slots busy | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')[1]} % |")
print(f"| warps active | {g('sm__warps_active.avg.pct_of_peak_sustained_active')[1]} % of peak |")
print(f"| registers / thread, grid, block | {g('launch__registers_per_thread')[1]}, {g('launch__grid_size')[1]}, {g('launch__block_size')[1]} |")
for p in ("alu", "fma", "fp64", "xu", "lsu"):
    print(f"| pipe {p} | {g(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active')[1]} % |")
print("\nstall reasons (warp-cycles per issued instruction):\n")
st = sorted(((float(v[1]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_")
             and k.endswith("_per_issue_active.ratio") and v[1] not in ("", "n/a")), reverse=True)
for val, k in st[:8]:
    print(f"* {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {val:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, text=True).stdout
cur = None
h = None
agg = []
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = r
    elif h is not None and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")]), r[1].strip()))
        except ValueError:
            pass
ti, ts = sum(a[2] for a in agg), sum(a[3] for a in agg)
print("\nhottest source lines by stall samples:\n\n| line | samples | instructions | source |\n|---|---|---|---|")
for a in sorted(agg, key=lambda x: -x[3])[:12]:
    print(f"| {a[0]}:{a[1]} | {100 * a[3] / max(ts, 1):.1f} % | {100 * a[2] / max(ti, 1):.1f} % | `{a[4][:70].replace('|', '/')}` |")
