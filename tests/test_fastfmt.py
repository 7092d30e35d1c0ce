"""cli/fastfmt.h (the exact "%f" / "%ld" formatters of the drop-in command line's output path) against glibc's
snprintf: the CLI's stdout must stay byte-identical to the reference's printf output (src/cfunc.c)."""
import ctypes as C
import os
import struct
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tools", "fastfmt_shim.c")
HDR = os.path.join(os.path.dirname(HERE), "cli", "fastfmt.h")
OUT = os.path.join(HERE, "tools", "libfastfmt.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", SRC, "-o", OUT])
    l = C.CDLL(OUT)
    l.fastfmt_selftest.restype = C.c_uint64
    l.fastfmt_selftest.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]
    l.fastfmt_int_selftest.restype = C.c_uint64
    l.fastfmt_int_selftest.argtypes = [C.c_uint64, C.c_uint64]
    l.fastfmt_one.argtypes = [C.c_float, C.c_char_p]
    return l


@pytest.mark.parametrize("mode,n", [(0, 4_000_000), (1, 4_000_000), (2, 4_000_000)])
def test_f6_equals_printf(lib, mode, n):
    bad_f = C.c_float(0)
    bad = lib.fastfmt_selftest(1234 + mode, n, mode, C.byref(bad_f))
    assert bad == 0, f"first mismatch at {bad_f.value!r}"


def test_f6_special_values(lib):
    buf = C.create_string_buffer(64)
    for bits, want in [(0x00000000, b"0.000000"), (0x80000000, b"-0.000000"), (0x00000001, b"0.000000"),
                       (0x3c000000, b"0.007812"),   # 2^-7 = 0.0078125: an exact tie, rounds to even
                       (0x3c400000, b"0.011719"),   # 0.01171875: the next tie, rounds up to even
                       (0x7f800000, b"inf"), (0xff800000, b"-inf"), (0x7f7fffff, None), (0x5f000000, None)]:
        f = struct.unpack("<f", struct.pack("<I", bits))[0]
        lib.fastfmt_one(f, buf)
        assert buf.value == (want if want is not None else (b"%f" % f))


def test_integers(lib):
    assert lib.fastfmt_int_selftest(7, 1_000_000) == 0
