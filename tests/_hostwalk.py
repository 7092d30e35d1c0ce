"""Builds and loads tests/tools/host_walk.cpp: the chunk walker of sigtk_b200/csrc/walk_core.cuh (the code every
thread of walk_chunks_kernel runs) compiled for the CPU. TEST INFRASTRUCTURE ONLY: it lets the CPU suite check the
chunking / register-ring / warm-up / boundary-state logic against the oracle; nothing in the product loads it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tools", "host_walk.cpp")
CORE = os.path.join(os.path.dirname(HERE), "sigtk_b200", "csrc", "walk_core.cuh")
OUT = os.path.join(HERE, "tools", "libhostwalk.so")


def load(force_redo=False, force_far=False):
    """force_redo: the build takes the rare exact-recompute path on every third block (WALK_TEST_REDO);
    force_far: every peak emitted later than the earliest possible step counts as too old for the block's mask
    (WALK_TEST_FAR), which sends its block through the same path"""
    out = OUT.replace(".so", "_redo.so") if force_redo else OUT.replace(".so", "_far.so") if force_far else OUT
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(__file__))
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC"]
                              + (["-DWALK_TEST_REDO"] if force_redo else []) + (["-DWALK_TEST_FAR"] if force_far else [])
                              + ["-x", "c++", SRC, "-o", out])
    lib = C.CDLL(out)
    lib.host_walk_read.restype = C.c_int
    lib.host_walk_read.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    return lib


def walk(lib, raw, dig, off, rng, rna, L, W, sh=0, thr_long=9.0, jobs_out=None):
    """-> (boundary mismatches, event starts int64[], pA f32[n], (raw min, raw max)); thr_long: the long detector's
    threshold (the reference's is 9.0); jobs_out: a list that receives the number of replayed long-detector lives"""
    n = len(raw)
    pad = np.zeros((n + 7) // 8 * 8 + 16, np.int16)
    pad[:n] = raw
    bm = np.zeros((n + sh + 31) // 32 + 2, np.uint32)
    pa = np.zeros(len(pad), np.float32)
    mm = np.zeros(2, np.int32)
    nj = np.zeros(1, np.int32)
    unit = np.float32(np.float32(rng) / np.float32(dig))  # misc.c:17-19,26
    mism = lib.host_walk_read(pad.ctypes.data, n, C.c_float(np.float32(off)), C.c_float(unit), int(rna), L, W, sh,
                              bm.ctypes.data, pa.ctypes.data, mm.ctypes.data, C.c_float(thr_long), nj.ctypes.data)
    if jobs_out is not None:
        jobs_out.append(int(nj[0]))
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")
    return mism, np.nonzero(bits)[0].astype(np.int64) - sh, pa[:n], (int(mm[0]), int(mm[1]))
