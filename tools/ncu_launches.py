#!/usr/bin/env python
"""Launch list (`ncu --metrics gpu__time_duration.sum --csv`) -> markdown table of OUR kernels:
   python tools/ncu_launches.py launches.csv 'command line' > profiles/rNN_launches.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
ours = ("ent_kernel", "jnn_walk", "walk_chunks", "long_jobs", "prefix_", "verify_chunks", "chunk_count", "emit_events", "count_tile_bits", "scan_counts",
        "read_event_offsets", "init_reads", "build_seq_list", "gen_", "sum_fixups", "pa_kernel", "stat_", "svb_")
agg = defaultdict(list)
for r in rows:
    name, val, unit = r[4], float(r[-1].replace(",", "")), r[-2]
    if not any(o in name for o in ours):
        continue
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * 1000.0 if unit in ("ms", "msecond") else val
    short = name.split("(")[0].split("::")[-1].replace("void ", "")
    agg[short].append(us)
tot = sum(sum(v) for v in agg.values())
print("# ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
print(f"command: `{sys.argv[2] if len(sys.argv) > 2 else ''}`;")
print("per-launch times are cold-cache and serialised: compare SHARES with `roofline.stage_ms_per_step` of bench.py, "
      "not absolutes.\n")
print("| kernel | launches | mean us | share of captured time |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f} % |")
