"""CPU: libisg.so builds for sm_100a, loads, and exports every symbol include/isg.h declares;
the ctypes table in isg_b200.lib covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import isg_b200
from isg_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "isg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isg_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_header_symbols():
    import __graft_entry__ as ge

    ge.build()
    assert os.path.exists(L.LIB_PATH)
    dll = ctypes.CDLL(L.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/isg.h but not exported by libisg.so"


def test_ctypes_table_matches_header():
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_header_argument_counts_match_ctypes_table():
    text = open(os.path.join(ROOT, "include", "isg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_res, args) in L.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_cubin_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_version_and_error_strings():
    lib = L.load()
    assert lib.isg_version() >= 100
    assert b"ISG_EINVAL" in lib.isg_error_string(-1)
    assert lib.isg_error_string(0) == b"ok"


def test_argument_validation_without_gpu():
    """Entry points validate arguments before touching the device: bad shapes return ISG_E* codes."""
    lib = L.load()
    assert lib.isg_gat_edge_fwd(None, None, 0, None, None, None, None, None, None, None, None, None, 0, None, 4, 4, 4,
                                301, 0.2, 0, None) == -2  # C % 4 != 0 -> ISG_EUNSUPPORTED
    assert lib.isg_linear_fwd(None, 0, None, None, None, None, None, 0, None, 0, -1, 4, 4, 0, 0, 0, None) == -1
    assert lib.isg_csr_build(None, -1, 4, None, None, None, None, None, None, None, None, 0, None) == -1
