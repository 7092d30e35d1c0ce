#!/usr/bin/env python
"""Instruction mix of a kernel's hottest loop from its SASS (cuobjdump -sass), by issue pipe.

  python tools/sass_mix.py sigtk_b200/csrc/walk.o walk_chunks_kernelILi0 [--samples-per-iter 8] [--excerpt N]

Finds the longest backward branch span of the kernel (the straight-line block loop of the chunk walker), counts its
instructions per pipe (ALU = integer / logic / compare / select / min-max, FMA = FADD/FMUL/FFMA/IMAD and their packed
forms, FP64, XU = MUFU and conversions, LSU = loads / stores / atomics, CTL = branches and barriers) and prints the
per-sample numbers next to the pipe costs measured by tools/microbench/pipes.cu (profiles/r02_pipes.json).
"""
import argparse
import collections
import json
import re
import subprocess

PIPE_COST = {"ALU": 2.05, "FMA": 1.12, "FP64": 2.19, "XU": 8.08, "LSU": 1.0, "CTL": 1.0, "OTHER": 1.0}


def pipe_of(op: str) -> str:
    base = op.split(".")[0]
    if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"):
        return "FP64"
    if base in ("MUFU", "F2F", "F2I", "I2F", "I2FP", "F2FP", "FRND", "POPC", "FLO", "BREV"):
        return "XU"
    if base in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "IMAD", "HFMA2", "HADD2", "HMUL2", "FADD32I", "FMUL32I", "FFMA32I"):
        return "FMA"
    if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "RED", "ATOM", "ATOMG", "LDC", "LD", "ST", "LDGSTS", "SHFL"):
        return "LSU"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "BAR", "JMP", "BRX", "YIELD", "BMOV", "DEPBAR"):
        return "CTL"
    if base in ("MOV", "IADD3", "IADD", "LOP3", "SHF", "SEL", "FSEL", "ISETP", "FSETP", "FMNMX", "PLOP3", "PRMT", "LEA", "IABS",
                "VIMNMX", "VIMNMX3", "IMNMX", "SGXT", "BFE", "BFI", "LOP", "SHL", "SHR", "P2R", "R2P", "CS2R", "S2R", "VABSDIFF",
                "VABSDIFF4", "FSET", "ISET", "FCHK", "FMNMX3", "VIADD", "VIADDMNMX", "IADD32I", "ISCADD", "FSWZADD", "R2UR", "UMOV"):
        return "ALU"
    if base.startswith("U"):
        return "OTHER"  # uniform datapath
    return "OTHER"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--samples-per-iter", type=float, default=8.0)
    ap.add_argument("--excerpt", type=int, default=0, help="print the first N instructions of the loop")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", a.obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = next(f for f in funcs if a.kernel in f.split("\n", 1)[0])
    name = body.split("\n", 1)[0].strip()
    ins = []  # (addr, opcode, text)
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), (m.group(2) or "") + m.group(3) + m.group(4)))
    addr_index = {ad: k for k, (ad, _, _) in enumerate(ins)}
    best = None
    for k, (ad, op, t) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < ad and tgt in addr_index:
                    span = k - addr_index[tgt]
                    if best is None or span > best[0]:
                        best = (span, addr_index[tgt], k)
    span, lo, hi = best
    loop = ins[lo:hi + 1]
    by_pipe = collections.Counter(pipe_of(op) for _, op, _ in loop)
    by_op = collections.Counter(op.split(".")[0] for _, op, _ in loop)
    n = len(loop)
    spi = a.samples_per_iter
    out = {"kernel": name, "loop_instructions": n, "samples_per_iteration": spi, "instr_per_sample": round(n / spi, 1),
           "per_pipe_per_sample": {p: round(c / spi, 1) for p, c in sorted(by_pipe.items())},
           "pipe_cycles_per_sample": {p: round(c / spi * PIPE_COST[p], 1) for p, c in sorted(by_pipe.items())},
           "top_opcodes": dict(by_op.most_common(24)),
           "note": "static count of the straight-line loop (rare-path call sites and their never-taken branches included); "
                   "pipe cycles = count x the per-warp-instruction issue cost of profiles/r02_pipes.json"}
    out["sum_pipe_cycles_per_sample"] = round(sum(out["pipe_cycles_per_sample"].values()), 1)
    print(json.dumps(out, indent=1))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    if a.excerpt:
        for ad, op, t in loop[:a.excerpt]:
            print(f"    /*{ad:05x}*/ {t.strip()}")


if __name__ == "__main__":
    main()
