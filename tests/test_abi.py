"""The C-ABI library loads on a CPU-only box and exports exactly what include/dmp_b200.h declares."""
import ctypes
import os
import re

from dualmessagepassing_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dmp_b200.h")).read()
    return sorted(set(re.findall(r"DMP_API\s+[\w\s\*]+?\b(dmp_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libdmp_b200.so does not export %s" % n


def test_binding_table_matches_header():
    assert sorted(list(_lib.SIGNATURES) + ["dmp_last_error"]) == _declared()


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.dmp_version() >= 100
    # an argument error must come back as a status + message, without touching the GPU
    n = ctypes.c_int64(0)
    rc = lib.dmp_plan_workspace_bytes(-1, 0, ctypes.byref(n))
    assert rc == -1
    assert b"N and E" in lib.dmp_last_error()


def test_header_constants_match_python_mirror():
    text = open(os.path.join(ROOT, "include", "dmp_b200.h")).read()
    defs = dict(re.findall(r"#define\s+(DMP_\w+)\s+\(?(-?\w+)\)?", text))
    assert int(defs["DMP_SEG_SIGN_BY_REV"]) == _lib.SEG_SIGN_BY_REV
    assert int(defs["DMP_SEG_NEGATE_OUT"]) == _lib.SEG_NEGATE_OUT
    assert int(defs["DMP_SEG_ONLY_FWD"]) == _lib.SEG_ONLY_FWD and int(defs["DMP_SEG_ONLY_REV"]) == _lib.SEG_ONLY_REV
    assert int(defs["DMP_SEG_SPLIT_BY_REV"]) == _lib.SEG_SPLIT_BY_REV and int(defs["DMP_SEG_SHORT"]) == _lib.SEG_SHORT
    assert int(defs["DMP_EDGE_MIRRORED_HALVES"]) == _lib.EDGE_MIRRORED_HALVES
    assert int(defs["DMP_ORDER_SCM"]) == _lib.ORDER_SCM and int(defs["DMP_ORDER_UNC"]) == _lib.ORDER_UNC
    assert int(defs["DMP_EID_MASK"], 16) == _lib.EID_MASK
    for i, k in enumerate(["NONE", "RELU", "LEAKY_RELU", "TANH", "SIGMOID"]):
        assert int(defs["DMP_ACT_" + k]) == i
