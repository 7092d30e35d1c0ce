#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (run in the build container only).

Needs /root/reference (read-only) and `make -C oracle ref` (oracle/_ref/{sigtk,blow5_dump,blow5_write}).
Nothing here runs on the GPU box; the fixtures it writes are committed.

  sp1_dna.blow5, event_dna.exp      the reference's own test input + golden file (scripts/test.sh:70-72)
  sp1_dna.npz                       flat dump of sp1_dna.blow5 (what the hot path consumes)
  ref_sp1_{event_c,stat}.txt        full stdout of the compiled reference (BASELINE config C1)
  ref_sp1_{event,pa}_first10.txt.gz stdout for the first 10 read ids; sha256 of the full stdout in sha256.json
  synth_rna.blow5 / .npz            12 seeded synthetic RNA reads, experiment_type=rna (BASELINE config C2;
                                    the real test/sequin_rna.blow5 is a missing large blob)
  ref_rna_{event,event_c,stat}.txt* stdout of the compiled reference on synth_rna.blow5
  ref_{sp1,rna}_ent.txt             stdout of `sigtk ent` (src/ent.c) on the two files
  ent_adversarial.npz / ref_ent_adversarial.txt
                                    seeded reads that leave the narrow value range of real signals (full-range noise,
                                    negative values, +-32767 steps, constant, 1- and 2-sample records) and what the
                                    compiled reference's `sigtk ent` prints for them   (`make_golden.py ent` makes only these)
  ref_{sp1,rna}_jnn{,_c}.txt        stdout of `sigtk jnn` / `jnn -c` (src/jnn.c) on the two files
  jnn_stalls_{dna,rna}.npz / ref_jnn_stalls_{dna,rna}{,_c}.txt
                                    seeded reads with stalls (long stretches near the read's mean, a few outliers inside,
                                    pairs closer than the merge distance, open stretches at the end, clipped spikes) and
                                    the compiled reference's `sigtk jnn` stdout             (`make_golden.py jnn`)
  prefix_dna.exp                    the reference's own golden for `sigtk prefix test/sp1_dna.blow5` (scripts/test.sh:54-56)
  ref_{sp1,rna}_prefix_stat.txt, prefix_adaptor_{dna,rna}.npz / ref_prefix_adaptor_{dna,rna}{,_stat}.txt
                                    `sigtk prefix [--print-stat]` of the compiled reference on the two files and on seeded
                                    reads with an adaptor and a poly-A-like stretch      (`make_golden.py prefix`)
"""
import gzip
import hashlib
import json
import os
import shutil
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sigtk_b200 import synth  # noqa: E402

REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")
SIGTK = os.path.join(BIN, "sigtk")


def run(args):
    return subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout


def parse_dump(blob):
    ids, lens, dig, off, rng, chunks = [], [], [], [], [], []
    p = 0
    while p < len(blob):
        (idl,) = struct.unpack_from("<I", blob, p); p += 4
        ids.append(blob[p:p + idl].decode()); p += idl
        n, d, o, r = struct.unpack_from("<Qddd", blob, p); p += 32
        chunks.append(np.frombuffer(blob, dtype="<i2", count=n, offset=p)); p += 2 * n
        lens.append(n); dig.append(d); off.append(o); rng.append(r)
    read_off = np.zeros(len(ids) + 1, dtype=np.uint64)
    read_off[1:] = np.cumsum(lens)
    return dict(read_ids=np.array(ids), samples=np.concatenate(chunks).astype(np.int16), read_off=read_off,
                digitisation=np.array(dig), offset=np.array(off), range=np.array(rng))


def dump_to_npz(blow5, npz):
    d = parse_dump(run([os.path.join(BIN, "blow5_dump"), blow5]))
    np.savez_compressed(npz, **d)
    return d


def write_blow5(reads, ids, path, exp_type):
    blob = bytearray()
    for (raw, dig, off, rng), rid in zip(reads, ids):
        b = rid.encode()
        blob += struct.pack("<I", len(b)) + b + struct.pack("<Qddd", raw.shape[0], dig, off, rng) + raw.tobytes()
    if os.path.exists(path):
        os.remove(path)
    subprocess.run([os.path.join(BIN, "blow5_write"), path, exp_type], input=bytes(blob), check=True)


def ent_adversarial_reads():
    """every |raw[i]-raw[i-1]| stays below 16,384: beyond that the reference aborts (assert c==out[i], ent.c:154)"""
    rng = np.random.default_rng(20260017)
    reads = []
    reads.append(rng.integers(-8191, 8192, 70000).astype(np.int16))               # 16,383 keys; deltas up to 16,382
    reads.append(rng.integers(-300, 300, 9000).astype(np.int16))                  # keys wrap around 0 / 65535
    reads.append(np.where(rng.random(5000) < 0.5, 8000, -8000).astype(np.int16))  # deltas of +-16,000
    reads.append(np.full(3000, 511, dtype=np.int16))                              # one raw bin: entropy 0
    reads.append(np.array([1234], dtype=np.int16))                                # n = 1: no deltas
    reads.append(np.array([-5, 17], dtype=np.int16))                              # n = 2: one delta
    base = (600 + 80 * rng.standard_normal(40000)).astype(np.int16)
    base[::997] = 8000                                                            # spikes leave the raw window
    base[5::1999] = -7000
    reads.append(base)
    reads.append((rng.integers(0, 2, 20000) * 4096 + 100 + rng.integers(0, 40, 20000)).astype(np.int16))  # window edge
    reads.append(np.arange(-8000, 8000, dtype=np.int32).astype(np.int16))         # 16,000 equal bins
    reads.append((-9000 + rng.integers(-200, 200, 30000)).astype(np.int16))       # all negative (keys >= 32768)
    return [(r, 8192.0, 3.0, 1402.882324) for r in reads]


def ent_goldens():
    sp1 = os.path.join(HERE, "sp1_dna.blow5")
    rna = os.path.join(HERE, "synth_rna.blow5")
    open(os.path.join(HERE, "ref_sp1_ent.txt"), "wb").write(run([SIGTK, "ent", sp1]))
    open(os.path.join(HERE, "ref_rna_ent.txt"), "wb").write(run([SIGTK, "ent", rna]))
    reads = ent_adversarial_reads()
    ids = [f"ent-adv-{i:02d}" for i in range(len(reads))]
    tmp = os.path.join(HERE, "_ent_adv.blow5")
    write_blow5(reads, ids, tmp, "genomic_dna")
    open(os.path.join(HERE, "ref_ent_adversarial.txt"), "wb").write(run([SIGTK, "ent", tmp]))
    os.remove(tmp)
    lens = [r[0].shape[0] for r in reads]
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    np.savez_compressed(os.path.join(HERE, "ent_adversarial.npz"), read_ids=np.array(ids),
                        samples=np.concatenate([r[0] for r in reads]), read_off=off,
                        digitisation=np.array([r[1] for r in reads]), offset=np.array([r[2] for r in reads]),
                        range=np.array([r[3] for r in reads]))


def jnn_stall_reads(rna):
    rng = np.random.default_rng(20260018 + int(rna))
    win = 1000 if rna else 150
    reads = []
    for k in range(14):
        n = int(rng.integers(6 * win, 60 * win))
        raw = (500 + 90 * rng.standard_normal(n)).astype(np.int64)
        level = np.repeat(rng.integers(-150, 150, n // 40 + 1), 40)[:n]
        raw += level
        pos = int(rng.integers(0, win))
        while pos < n:
            length = int(rng.integers(win // 5, 4 * win))
            stop = min(n, pos + length)
            raw[pos:stop] = 500 + rng.integers(-12, 13, stop - pos)          # a stall: inside the band
            if k % 3 == 0 and stop - pos > 50:
                bad = rng.integers(pos + 5, stop - 5, int(rng.integers(1, 8)))  # a few outliers inside (tolerated up to 5)
                raw[bad] = 900
            pos = stop + int(rng.integers(5, 60) if k % 2 else rng.integers(40, 6 * win))  # gaps around the merge distance
        if k % 4 == 0:
            raw[::1013] = 5000                                               # clipped to 1200 by rm_outlier
            raw[7::2029] = -300                                              # clipped to 0
        if k == 5:
            raw[-2 * win:] = 500                                             # a stretch still open at the end
        reads.append((np.clip(raw, -32768, 32767).astype(np.int16), 8192.0, 3.0, 1402.882324))
    reads.append((np.full(5 * win, 480, dtype=np.int16), 8192.0, 3.0, 1402.882324))  # zero deviation: empty band
    reads.append((np.array([5, 900, 20], dtype=np.int16), 8192.0, 3.0, 1402.882324))
    return reads


def jnn_goldens():
    sp1 = os.path.join(HERE, "sp1_dna.blow5")
    rna = os.path.join(HERE, "synth_rna.blow5")
    for tag, path in (("sp1", sp1), ("rna", rna)):
        open(os.path.join(HERE, f"ref_{tag}_jnn.txt"), "wb").write(run([SIGTK, "jnn", path]))
        open(os.path.join(HERE, f"ref_{tag}_jnn_c.txt"), "wb").write(run([SIGTK, "jnn", "-c", path]))
    for flag, kind in ((0, "dna"), (1, "rna")):
        reads = jnn_stall_reads(flag)
        ids = [f"jnn-{kind}-{i:02d}" for i in range(len(reads))]
        tmp = os.path.join(HERE, "_jnn.blow5")
        write_blow5(reads, ids, tmp, "rna" if flag else "genomic_dna")
        open(os.path.join(HERE, f"ref_jnn_stalls_{kind}.txt"), "wb").write(run([SIGTK, "jnn", tmp]))
        open(os.path.join(HERE, f"ref_jnn_stalls_{kind}_c.txt"), "wb").write(run([SIGTK, "jnn", "-c", tmp]))
        os.remove(tmp)
        lens = [r[0].shape[0] for r in reads]
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        np.savez_compressed(os.path.join(HERE, f"jnn_stalls_{kind}.npz"), read_ids=np.array(ids),
                            samples=np.concatenate([r[0] for r in reads]), read_off=off,
                            digitisation=np.array([r[1] for r in reads]), offset=np.array([r[2] for r in reads]),
                            range=np.array([r[3] for r in reads]))


def prefix_adaptor_reads(rna):
    """reads with a low-current adaptor stretch near the start, a flat poly-A-like stretch behind it, then signal"""
    rng = np.random.default_rng(20260019 + int(rna))
    reads = []
    for k in range(12):
        n = int(rng.integers(20000, 90000))
        raw = (525 + 70 * rng.standard_normal(n)).astype(np.int64)
        lead = int(rng.integers(200, 3000))
        alen = int(rng.integers(3500, 12000)) if k != 3 else 0        # read 3: no adaptor at all
        raw[lead:lead + alen] = 380 + 25 * rng.standard_normal(alen)
        plen = int(rng.integers(300, 4000)) if k % 4 != 1 else 0       # every fourth read: no poly-A stretch
        a = lead + alen
        raw[a:a + plen] = 575 + 12 * rng.standard_normal(plen)
        if k % 5 == 0:
            raw[::1501] = 4000                                          # clipped spikes
        if k == 7:
            raw[n // 2: n // 2 + 5000] = 370 + 20 * rng.standard_normal(5000)   # a second low stretch
        reads.append((np.clip(raw, -32768, 32767).astype(np.int16), 8192.0, 3.0, 1402.882324))
    reads.append(((525 + 70 * rng.standard_normal(1500)).astype(np.int16), 8192.0, 3.0, 1402.882324))  # n <= window
    return reads


def prefix_goldens():
    sp1 = os.path.join(HERE, "sp1_dna.blow5")
    rna = os.path.join(HERE, "synth_rna.blow5")
    shutil.copyfile(os.path.join(REF, "test", "prefix_dna.exp"), os.path.join(HERE, "prefix_dna.exp"))
    os.chmod(os.path.join(HERE, "prefix_dna.exp"), 0o644)
    for tag, path in (("sp1", sp1), ("rna", rna)):
        open(os.path.join(HERE, f"ref_{tag}_prefix_stat.txt"), "wb").write(run([SIGTK, "prefix", "--print-stat", path]))
    for flag, kind in ((0, "dna"), (1, "rna")):
        reads = prefix_adaptor_reads(flag)
        ids = [f"adaptor-{kind}-{i:02d}" for i in range(len(reads))]
        tmp = os.path.join(HERE, "_prefix.blow5")
        write_blow5(reads, ids, tmp, "rna" if flag else "genomic_dna")
        open(os.path.join(HERE, f"ref_prefix_adaptor_{kind}.txt"), "wb").write(run([SIGTK, "prefix", tmp]))
        open(os.path.join(HERE, f"ref_prefix_adaptor_{kind}_stat.txt"), "wb").write(run([SIGTK, "prefix", "--print-stat", tmp]))
        os.remove(tmp)
        lens = [r[0].shape[0] for r in reads]
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        np.savez_compressed(os.path.join(HERE, f"prefix_adaptor_{kind}.npz"), read_ids=np.array(ids),
                            samples=np.concatenate([r[0] for r in reads]), read_off=off,
                            digitisation=np.array([r[1] for r in reads]), offset=np.array([r[2] for r in reads]),
                            range=np.array([r[3] for r in reads]))


def main():
    if sys.argv[1:] == ["prefix"]:
        prefix_goldens()
        print("prefix fixtures written to", HERE)
        return
    if sys.argv[1:] == ["jnn"]:
        jnn_goldens()
        print("jnn fixtures written to", HERE)
        return
    if sys.argv[1:] == ["ent"]:
        ent_goldens()
        print("ent fixtures written to", HERE)
        return
    sha = {}
    shutil.copyfile(os.path.join(REF, "test", "sp1_dna.blow5"), os.path.join(HERE, "sp1_dna.blow5"))
    shutil.copyfile(os.path.join(REF, "test", "event_dna.exp"), os.path.join(HERE, "event_dna.exp"))
    os.chmod(os.path.join(HERE, "sp1_dna.blow5"), 0o644)
    os.chmod(os.path.join(HERE, "event_dna.exp"), 0o644)
    sp1 = os.path.join(HERE, "sp1_dna.blow5")
    d = dump_to_npz(sp1, os.path.join(HERE, "sp1_dna.npz"))
    first10 = list(d["read_ids"][:10])
    for mode, args in (("event_c", ["event", "-c"]), ("stat", ["stat"])):
        out = run([SIGTK] + args + [sp1])
        open(os.path.join(HERE, f"ref_sp1_{mode}.txt"), "wb").write(out)
        sha[f"sp1_{mode}"] = hashlib.sha256(out).hexdigest()
    for mode in ("event", "pa"):
        full = run([SIGTK, mode, sp1])
        sha[f"sp1_{mode}"] = hashlib.sha256(full).hexdigest()
        part = run([SIGTK, mode, sp1] + first10)
        with gzip.GzipFile(os.path.join(HERE, f"ref_sp1_{mode}_first10.txt.gz"), "wb", mtime=0) as f:
            f.write(part)

    # synthetic RNA (config C2): experiment_type=rna makes the reference pick the RNA parameters
    reads = synth.make_reads(12, mean=25000.0, sigma=0.5, seed=11, rna=True)
    ids = [f"synth-rna-{i:04d}" for i in range(len(reads))]
    rna = os.path.join(HERE, "synth_rna.blow5")
    write_blow5(reads, ids, rna, "rna")
    dump_to_npz(rna, os.path.join(HERE, "synth_rna.npz"))
    for mode, args in (("event_c", ["event", "-c"]), ("stat", ["stat"])):
        out = run([SIGTK] + args + [rna])
        open(os.path.join(HERE, f"ref_rna_{mode}.txt"), "wb").write(out)
        sha[f"rna_{mode}"] = hashlib.sha256(out).hexdigest()
    full = run([SIGTK, "event", rna])
    sha["rna_event"] = hashlib.sha256(full).hexdigest()
    with gzip.GzipFile(os.path.join(HERE, "ref_rna_event.txt.gz"), "wb", mtime=0) as f:
        f.write(full)
    full = run([SIGTK, "pa", rna])
    sha["rna_pa"] = hashlib.sha256(full).hexdigest()
    ent_goldens()
    jnn_goldens()
    prefix_goldens()
    json.dump(sha, open(os.path.join(HERE, "sha256.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
