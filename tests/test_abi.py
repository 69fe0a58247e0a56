"""CPU: libisg.so builds for sm_100a, loads, and exports every symbol include/isg.h declares;
the ctypes table in isg_b200.lib covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

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


def test_executor_slots_used_by_the_host_exist():
    """isubgvqa/executor.py addresses the layer executor's flat arrays by slot NAME (isg_layer_slot); every name the
    host side uses — including the launch-overlap slots (side-stream lin_edge, deferred join) — resolves, and the three
    tables have the sizes the library reports."""
    dll = ctypes.CDLL(L.LIB_PATH)
    dll.isg_layer_slot.restype = ctypes.c_int
    dll.isg_layer_slot.argtypes = [ctypes.c_char_p]
    src = open(os.path.join(ROOT, "intrinsic-subgraph-generation-for-vqa_b200", "isubgvqa", "executor.py")).read()
    names = set(re.findall(r"\bs\.((?:D|F|P)_[A-Z0-9_]+)\b", src)) | set(re.findall(r"\"(P_[A-Z0-9_]*[A-Z0-9])\"", src))
    assert {"D_EPROJ_READY", "P_EV_EPROJ", "D_DEFER_JOIN", "P_EV_WS_FREE", "D_SIDE_WGRAD", "P_SIDE_STREAM"} <= names
    missing = [n for n in sorted(names) if dll.isg_layer_slot(n.encode()) < 0]
    assert not missing, f"slots used by executor.py but unknown to libisg.so: {missing}"
    assert dll.isg_layer_slot(b"P_NO_SUCH_SLOT") < 0
    for which in (0, 1, 2):
        assert dll.isg_layer_slot_count(which) > 0


def test_header_is_c99_and_links_from_c(tmp_path):
    """include/isg.h is a C header (strict C99, no C++ / torch types) and libisg.so is callable from a plain-C
    program: tests/abi_c/abi_check.c is compiled with gcc -std=c99 -pedantic -Werror, linked against the library and
    run (version, error strings, host-side argument validation — no device needed)."""
    import __graft_entry__ as ge

    ge.build()
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(L.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "abi_c", "abi_check.c"), "-o", exe, "-L", libdir, "-lisg",
                         "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi_check ok" in run.stdout


_SWEEP = r"""
import ctypes, sys
sys.path.insert(0, sys.argv[1])
FILL = int(sys.argv[2])
FAKE = ctypes.c_void_p(0x7f0000001000) if len(sys.argv) > 3 else None
HOST_ARRAYS = {"isg_weights_to_bf16", "isg_colsum_multi", "isg_grad_sq_partials", "isg_adam_update", "isg_opt_blocks",
               "isg_mgat_layer_fwd", "isg_mgat_layer_bwd", "isg_mgat_layer_bwd_workspace_bytes", "isg_layer_slot",
               "isg_error_string"}  # these read HOST memory through their pointers: not called with a fake address
from isg_b200 import lib as L
lib = L.load()
for name in sorted(L.SIGNATURES):
    if FAKE is not None and name in HOST_ARRAYS:
        continue
    res, args = L.SIGNATURES[name]
    vals = []
    for a in args:
        if a in (ctypes.c_float, ctypes.c_double):
            vals.append(0.0)
        elif a in (ctypes.c_size_t, ctypes.c_uint64):
            vals.append(max(FILL, 0))
        elif a in (ctypes.c_int, ctypes.c_int64, ctypes.c_int32):
            vals.append(FILL)
        else:
            vals.append(FAKE)  # every pointer null — or, third argument "fake", a non-null address that is never mapped
    print("CALL", name, flush=True)
    r = getattr(lib, name)(*vals)
    print("RET", name, r if not isinstance(r, bytes) else r.decode(), flush=True)
"""


@pytest.mark.parametrize("fill", [0, -1, 8, 300, 2147483647])
def test_every_entry_point_survives_null_pointers(tmp_path, fill):
    """"Nothing throws, nothing exits" (include/isg.h): each of the entry points is called in a child process with null
    pointers and every integer argument set to `fill` (an empty problem, negative sizes, plausible sizes, INT_MAX) — it must
    return (ISG_OK for an empty problem, a negative ISG_E* code for a rejected one, a byte count for the sizing
    helpers), never crash, and never reach the device (a positive return is a cudaError_t)."""
    script = tmp_path / "sweep.py"
    script.write_text(_SWEEP)
    run = subprocess.run([os.sys.executable, str(script), ROOT, str(fill)], capture_output=True, text=True, timeout=300)
    calls = re.findall(r"^CALL (\S+)$", run.stdout, flags=re.M)
    rets = dict(re.findall(r"^RET (\S+) (.*)$", run.stdout, flags=re.M))
    crashed = [c for c in calls if c not in rets]
    assert run.returncode == 0 and not crashed, f"crashed in {crashed} (exit code {run.returncode}): {run.stderr[-400:]}"
    assert sorted(rets) == sorted(L.SIGNATURES)
    sizing = {n for n in rets if n.endswith(("_bytes", "_count", "_npad", "_blocks", "_tensors")) or n in ("isg_version",)}
    for name, r in rets.items():
        if name == "isg_error_string":
            assert r
        elif name in sizing:  # (isg_layer_slot_count answers -1 for a table that does not exist)
            assert int(r) >= 0 or (name == "isg_layer_slot_count" and int(r) == -1), (name, r)
        else:
            assert int(r) in (0, -1, -2, -3), (name, r)


_HOST_ARRAYS = r"""
import ctypes, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from isg_b200 import lib as L
lib = L.load()
nd, nf, npt = (lib.isg_layer_slot_count(i) for i in range(3))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
slot = lambda n: lib.isg_layer_slot(n.encode())
base = dict(D_N=100, D_E=500, D_B=4, D_D=300, D_H=4, D_HID=600, D_NMAX=30)
for label, setup in [("zeros", {}), ("null_ptrs", base), ("masked_null_ptrs", dict(base, D_MASKED=1, D_SAMPLER=1, D_K=2)),
                     ("negative", dict(base, D_N=-5, D_E=-1, D_B=-1)), ("bf16_odd", dict(D_N=10, D_E=20, D_B=1, D_D=7, D_H=3, D_HID=9, D_BF16=1))]:
    d, f, p = np.zeros(nd, dtype=np.int64), np.zeros(nf, dtype=np.float64), np.zeros(npt, dtype=np.uint64)
    for k, v in setup.items():
        d[slot(k)] = v
    print("CALL", label, flush=True)
    print("RET", label, lib.isg_mgat_layer_bwd_workspace_bytes(P(d)), lib.isg_mgat_layer_fwd(P(d), P(f), P(p), None),
          lib.isg_mgat_layer_bwd(P(d), P(f), P(p), None), flush=True)
n = 3
vp, ints, i64 = (ctypes.c_void_p * n)(), (ctypes.c_int * n)(*[8] * n), (ctypes.c_int64 * n)(*[8] * n)
for label, fn in [("weights_to_bf16", lambda: lib.isg_weights_to_bf16(n, vp, ints, ints, vp, ints, vp, ints, None)),
                  ("weights_to_bf16_too_many", lambda: lib.isg_weights_to_bf16(100, vp, ints, ints, vp, ints, vp, ints, None)),
                  ("colsum_multi", lambda: lib.isg_colsum_multi(n, vp, ints, i64, i64, ints, vp, None, 0, None)),
                  ("grad_sq_partials", lambda: lib.isg_grad_sq_partials(vp, i64, n, None, None, 0, None, None)),
                  ("adam_update", lambda: lib.isg_adam_update(vp, vp, vp, vp, i64, n, None, None, None, 1e-3, 0.9, 0.999, 1e-8, None)),
                  ("adam_update_too_many", lambda: lib.isg_adam_update(vp, vp, vp, vp, i64, 1000, None, None, None, 1e-3, 0.9, 0.999, 1e-8, None))]:
    print("CALL", label, flush=True)
    print("RET", label, 0, fn(), 0, flush=True)
"""


def test_host_array_entry_points_reject_null_device_pointers(tmp_path):
    """The entry points that take HOST arrays (the layer executor's dims / scalars / pointer tables, the lists of
    tensors of the optimizer, the column sums and the bf16 weight conversion) get well-formed host arrays whose device
    pointers are all null, with empty, plausible, negative and unsupported dimensions: every call returns a negative
    ISG_E* code without touching the device, and the workspace query never wraps around."""
    script = tmp_path / "host_arrays.py"
    script.write_text(_HOST_ARRAYS)
    run = subprocess.run([os.sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=300)
    calls = re.findall(r"^CALL (\S+)$", run.stdout, flags=re.M)
    rets = {m[0]: tuple(int(v) for v in m[1:]) for m in re.findall(r"^RET (\S+) (-?\d+) (-?\d+) (-?\d+)$", run.stdout, flags=re.M)}
    assert run.returncode == 0 and calls and sorted(calls) == sorted(rets), (run.returncode, run.stderr[-400:])
    for label, (ws, a, b) in rets.items():
        assert a in (-1, -2, -3) and b in (0, -1, -2, -3), (label, a, b)
        assert 0 <= ws < 1 << 40, (label, ws)
    assert rets["zeros"][0] == 0 and rets["negative"][0] == 0 and rets["null_ptrs"][0] > 0


def test_negative_sizes_are_rejected_before_any_pointer_is_used(tmp_path):
    """Non-null (never mapped) device pointers with every size negative: each entry point answers a negative ISG_E*
    code (or 0 bytes) — no launch is attempted, nothing is dereferenced on the host."""
    script = tmp_path / "sweep.py"
    script.write_text(_SWEEP)
    run = subprocess.run([os.sys.executable, str(script), ROOT, "-3", "fake"], capture_output=True, text=True, timeout=300)
    calls = re.findall(r"^CALL (\S+)$", run.stdout, flags=re.M)
    rets = dict(re.findall(r"^RET (\S+) (.*)$", run.stdout, flags=re.M))
    assert run.returncode == 0 and len(calls) >= 50 and sorted(calls) == sorted(rets), (run.returncode, run.stderr[-400:])
    for name, r in rets.items():
        if name.endswith(("_bytes", "_count", "_npad", "_blocks", "_tensors")) or name == "isg_version":
            assert -1 <= int(r) < 1 << 20, (name, r)  # a fixed header of a few hundred bytes at most
        else:
            assert int(r) in (-1, -2, -3), (name, r)
