"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sigtk_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from sigtk_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sigtk_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sigtk_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert lib.sgpu_abi_version() == 5


def test_python_mirror_matches_the_header():
    """the WANT_* / F_* constants and the result struct of the ctypes mirror follow include/sigtk_b200.h"""
    hdr = open(os.path.join(ROOT, "include", "sigtk_b200.h")).read()
    defs = {k: int(v.rstrip("u"), 0) for k, v in re.findall(r"#define\s+(SGPU_[A-Z_]+)\s+(-?\d+u?)\b", hdr)}
    for name in ("EVENTS", "PA", "STAT", "ENT", "JNN"):
        assert getattr(_lib, "WANT_" + name) == defs["SGPU_WANT_" + name]
    for name in ("DEFAULT", "FORCE_GENERIC", "NO_HOST_SLOTS", "STAGE_TIMERS"):
        assert getattr(_lib, "F_" + name) == defs["SGPU_F_" + name]
    assert _lib.ALIGN == defs["SGPU_ALIGN"] and _lib.E_FULL == defs["SGPU_E_FULL"] and _lib.E_TOOBIG == defs["SGPU_E_TOOBIG"]
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    fields = re.findall(r"\*?\s*(\w+);", re.search(r"typedef struct \{([^}]*)\} sgpu_result_t;", body).group(1))
    assert fields == [f[0] for f in _lib.Result._fields_]
    fields = re.findall(r"\*?\s*(\w+);", re.search(r"typedef struct \{([^}]*)\} sgpu_counters_t;", body).group(1))
    assert fields == [f[0] for f in _lib.Counters._fields_]
    for name in ("CHUNK_LEN", "WARMUP", "THR_LONG", "PORE", "STAT_CTA_MIN"):
        assert getattr(_lib, "PARAM_" + name) == defs["SGPU_PARAM_" + name]


def test_strerror_and_argument_checks():
    lib = _lib.load()
    assert lib.sgpu_strerror(0) == b"success"
    assert b"no CPU fallback" in lib.sgpu_strerror(-2)
    h = C.c_void_p()
    assert lib.sgpu_create(C.byref(h), 0, 0, 0, 2, 0) == -1  # INVAL before any CUDA call
    assert lib.sgpu_submit(None, 0, 1) == -1
    assert lib.sgpu_wait(None, 0, None) == -1


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load()
    if lib.sgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.sgpu_create(C.byref(h), 0, 1 << 20, 16, 2, 0) == -2  # SGPU_E_CUDA, loudly
    assert not h.value


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sigtk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                code = "\n".join(ln.split("//")[0] for ln in src.splitlines())  # comments may cite the proofs
                assert "oracle" not in code and "libsigtk_ref" not in code and "_ref/" not in code, f"{f} uses the oracle"
