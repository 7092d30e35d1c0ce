#!/usr/bin/env python
"""Whole-CLI comparison on the GPU box (CPU baseline (ii)): the unmodified reference `sigtk`
(oracle/_ref/sigtk, 1 thread) against the drop-in `cli/sigtk` on the same synthetic BLOW5 file. Checks that stdout is
byte-identical (sha256) and prints one JSON line per sub-command with both wall times.

  python tools/cli_bench.py [--reads 2000] [--mean 40000] [--modes event-c,event,pa,stat,jnn,ent] [--gpus 1]
"""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigtk_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "sigtk")
WRITE = os.path.join(ROOT, "oracle", "_ref", "blow5_write")
CLI = os.path.join(ROOT, "cli", "sigtk")


def make_blow5(path, n_reads, mean, rna=False):
    lens = synth.read_lengths(n_reads, mean=mean)
    p = subprocess.Popen([WRITE, path, "rna" if rna else "genomic_dna"], stdin=subprocess.PIPE)
    total = 0
    for i in range(n_reads):
        raw, dig, off, rng = synth.make_read(i, int(lens[i]), p_change=0.025 if rna else 0.1)
        rid = f"synth-{i:08d}".encode()
        p.stdin.write(struct.pack("<I", len(rid)) + rid + struct.pack("<Qddd", len(raw), dig, off, rng) + raw.tobytes())
        total += len(raw)
    p.stdin.close()
    assert p.wait() == 0
    return total


def timed(cmd, out_path):
    """wall time with stdout redirected to a file (a pipe into Python would dominate for GB-sized outputs)"""
    with open(out_path, "wb") as fo:
        t0 = time.time()
        p = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, check=True, env=dict(os.environ, SIGTK_PROFILE="1"))
        dt = time.time() - t0
    timed.open_s = timed.wait_s = None
    for line in p.stderr.decode(errors="replace").splitlines():
        if "host wall clock: open" in line:  # CUDA context creation + pinned slots (seconds on a box without persistence mode)
            timed.open_s = float(line.split("pinned slots) ")[1].split(" s")[0])
            if "main thread waited " in line:   # it runs on a helper thread next to the first reads and inflates
                timed.wait_s = float(line.split("main thread waited ")[1].split(" s")[0])
    h = hashlib.sha256()
    with open(out_path, "rb") as fi:
        for blk in iter(lambda: fi.read(1 << 24), b""):
            h.update(blk)
    return dt, h.hexdigest(), os.path.getsize(out_path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000)
    ap.add_argument("--mean", type=float, default=40000.0)
    ap.add_argument("--modes", default="event-c,event,stat,pa,jnn,ent")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "synth.blow5")
        t0 = time.time()
        n = make_blow5(f, a.reads, a.mean)
        print(f"# {a.reads} reads, {n} samples, {os.path.getsize(f) / 1e6:.1f} MB BLOW5, made in {time.time() - t0:.1f} s",
              file=sys.stderr)
        for mode in a.modes.split(","):
            args = {"event-c": ["event", "-c"], "event": ["event"], "stat": ["stat"], "pa": ["pa"], "jnn": ["jnn"], "ent": ["ent"],
                    "prefix": ["prefix", "--print-stat"]}[mode] + [f]
            extra = ["--gpus", str(a.gpus)] + (["--threads", str(a.threads)] if a.threads else [])
            o = os.path.join(d, "out.txt")
            timed([CLI] + args + extra, o)  # warm-up: driver / file cache
            t_ours, h_ours, nb = timed([CLI] + args + extra, o)
            open_s, wait_s = timed.open_s, timed.wait_s
            t_ref, h_ref, _ = timed([REF] + args, o)
            print(json.dumps({"mode": mode, "reads": a.reads, "samples": n, "stdout_bytes": nb,
                              "identical_stdout": h_ours == h_ref, "reference_s": round(t_ref, 3),
                              "ours_s": round(t_ours, 3), "speedup": round(t_ref / t_ours, 2),
                              "ours_cuda_init_s": open_s, "ours_waited_for_cuda_init_s": wait_s,
                              "speedup_excluding_cuda_init": round(t_ref / max(t_ours - (wait_s if wait_s is not None else (open_s or 0.0)), 1e-3), 2),
                              "ours_msamples_per_s": round(n / t_ours / 1e6, 1),
                              "reference_msamples_per_s": round(n / t_ref / 1e6, 1), "gpus": a.gpus}), flush=True)


if __name__ == "__main__":
    main()
