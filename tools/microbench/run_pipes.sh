#!/bin/bash
# on the GPU box: build and run the pipe-cost microbenchmark, keep its output as text + JSON (profiles/rNN_pipes.json)
# usage: tools/microbench/run_pipes.sh OUT_PREFIX
set -e
cd "$(dirname "$0")"
out=${1:-../../gpurun_out/pipes}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o /tmp/pipes_bench
/tmp/pipes_bench | tee "$out.txt"
python - "$out.txt" "$out.json" <<'PY'
import json, re, sys
rows = []
for line in open(sys.argv[1]):
    m = re.match(r"(.+?)\s+warps/SMSP=(\d+)\s+block0:\s+([\d.]+) cyc/rep/warp \| whole GPU:\s+([\d.]+) cyc/rep/SMSP\s+\((\d+) instr/rep -> ([\d.]+) cyc/instr\)", line)
    if m:
        rows.append({"mix": m.group(1).strip(), "warps_per_smsp": int(m.group(2)), "cyc_per_rep_warp_block0": float(m.group(3)),
                     "cyc_per_rep_smsp": float(m.group(4)), "instr_per_rep": int(m.group(5)), "cyc_per_instr": float(m.group(6))})
json.dump({"what": "issue cost per warp-instruction and SM sub-partition (tools/microbench/pipes.cu), clock 1.965 GHz assumed", "rows": rows},
          open(sys.argv[2], "w"), indent=1)
PY
