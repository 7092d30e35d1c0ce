#!/usr/bin/env python
"""Generates tests/golden/svbzd_golden.npz with the UNMODIFIED reference slow5lib (oracle/_ref/libsigtk_ref.so,
built from /root/reference by `make -C oracle ref`): raw int16 signals and the svb-zd streams
slow5_ptr_compress_solo(SLOW5_COMPRESS_SVB_ZD) produces for them, plus hand-made streams with 4-byte values
(which the encoder never emits for int16 input) decoded by slow5_ptr_depress_solo. Run here, commit the .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _fmt  # noqa: E402
from _oracle import Reference  # noqa: E402

ref = Reference()
rng = np.random.default_rng(20260017)
sp1 = _fmt.load_npz(os.path.join(HERE, "sp1_dna.npz"))
raws = {
    "empty": np.zeros(0, np.int16),
    "one": np.array([-7], np.int16),
    "three": np.array([5, -5, 5], np.int16),
    "extremes": np.array([32767, -32768, 32767, 0, -32768, -32768, 32767, 1, 2, 3, 300, -300, 129, -129], np.int16),
    "noise_wide": rng.integers(-32768, 32768, 1537, dtype=np.int64).astype(np.int16),
    "noise_narrow": (500 + rng.integers(-40, 41, 4099)).astype(np.int16),
    "walk": np.cumsum(rng.integers(-200, 201, 2050)).astype(np.int16),
    "sp1_read0": sp1[0][1][0][:3000].astype(np.int16),
    "sp1_read7": sp1[7][1][0].astype(np.int16),
}
out = {}
for k, raw in raws.items():
    st = ref.svbzd_encode(raw)
    back = ref.svbzd_decode(st)
    assert back is not None and np.array_equal(back, raw), k
    out["raw_" + k] = raw
    out["svb_" + k] = st
# streams with 3- and 4-byte values: count=6, codes 3,0,2,3,1,3 ; decoded by the reference
count = np.array([6], np.uint32).view(np.uint8)
codes = [3, 0, 2, 3, 1, 3]
keys = np.zeros(2, np.uint8)
data = []
vals = [0x01020304, 0x7f, 0x0a0b0c, 0xfffffffe, 0x1234, 0x80000001]
for i, (c, v) in enumerate(zip(codes, vals)):
    keys[i >> 2] |= c << (2 * (i & 3))
    data += [(v >> (8 * b)) & 0xff for b in range(c + 1)]
st = np.concatenate([count, keys, np.array(data, np.uint8)])
dec = ref.svbzd_decode(st)
assert dec is not None and len(dec) == 6
out["svb_wide_codes"] = st
out["raw_wide_codes"] = dec
# a truncated stream must be rejected (slow5_press.c:1103)
assert ref.svbzd_decode(out["svb_walk"][:-1], cap=len(raws["walk"])) is None
np.savez_compressed(os.path.join(HERE, "svbzd_golden.npz"), **out)
print("wrote", len(out) // 2, "cases,", sum(v.nbytes for v in out.values()), "bytes")
