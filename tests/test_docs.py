"""CPU: the documents a maintainer reads stay in step with the code.  (1) INTEGRATION.md's table of entry points —
"the complete table" — names every symbol include/isg.h declares; (2) DESIGN.md section 9 lists exactly the ISG_*
environment switches the library, the host package and bench.py read; (3) the entry-point count quoted in README.md /
DESIGN.md is the header's."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "intrinsic-subgraph-generation-for-vqa_b200")


def _header_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "isg.h")).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(isg_[a-z0-9_]+)\s*\(", text)))


def test_integration_table_names_every_entry_point():
    mentioned = set()
    for line in open(os.path.join(ROOT, "INTEGRATION.md")).read().splitlines():
        if not line.startswith("|"):
            continue
        full = re.findall(r"`(isg_[a-z0-9_]+)`", line)
        suffixes = re.findall(r"`(_[a-z0-9_]+)`", line)  # the table's shorthand: `isg_x_fwd` / `_bwd`
        mentioned.update(full)
        for name in full:
            parts = name.split("_")
            for cut in range(1, len(parts)):
                mentioned.update("_".join(parts[:cut]) + s for s in suffixes)
    missing = [s for s in _header_symbols() if s not in mentioned]
    assert not missing, f"declared in include/isg.h but absent from INTEGRATION.md's table: {missing}"


def test_design_lists_exactly_the_environment_switches_the_code_reads():
    read = set()
    files = [p for p in glob.glob(os.path.join(PKG, "**", "*"), recursive=True)
             if p.endswith((".py", ".cu", ".cuh", ".h")) and os.sep + "build" + os.sep not in p]
    for p in files + [os.path.join(ROOT, "bench.py")]:
        read.update(re.findall(r'(?:getenv\(|environ\.get\(|environ\[)\s*"(ISG_[A-Z0-9_]+)"', open(p).read()))
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    documented = set(re.findall(r"`(ISG_[A-Z0-9_]+)`", design[design.index("## 9. Switches"):]))
    assert read - documented == set(), f"read by the code, missing from DESIGN.md section 9: {sorted(read - documented)}"
    assert documented - read == set(), f"documented in DESIGN.md section 9 but read nowhere: {sorted(documented - read)}"


def test_quoted_entry_point_count_is_the_headers():
    n = len(_header_symbols())
    for doc in ("README.md", "DESIGN.md"):
        counts = re.findall(r"(\d+) entry points", open(os.path.join(ROOT, doc)).read())
        assert counts and all(int(c) == n for c in counts), (doc, counts, n)


def test_reference_citations_point_at_existing_lines():
    """Every `path.py:line[-line]` citation of the reference — in include/isg.h, the documents, the package, the oracle,
    the tests and bench.py — names a file of the reference tree and a line range inside it (skipped where the tree is
    absent)."""
    import pytest

    ref = os.environ.get("ISG_REFERENCE_SRC", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "ISubGVQA", "models")):
        pytest.skip("the reference tree is not present on this machine")
    lengths = {}
    for p in glob.glob(os.path.join(ref, "**", "*.py"), recursive=True):
        with open(p, errors="ignore") as f:
            lengths[os.path.relpath(p, ref)] = sum(1 for _ in f)
    own = {os.path.relpath(p, ROOT) for p in glob.glob(os.path.join(ROOT, "**", "*.py"), recursive=True)}
    files = [os.path.join(ROOT, n) for n in ("include/isg.h", "DESIGN.md", "INTEGRATION.md", "README.md", "bench.py")]
    for pattern in ("intrinsic-subgraph-generation-for-vqa_b200/**/*.py", "intrinsic-subgraph-generation-for-vqa_b200/csrc/*",
                    "oracle/*.py", "tests/*.py"):
        files += glob.glob(os.path.join(ROOT, pattern), recursive=True)
    cite = re.compile(r"((?:[A-Za-z_]+/)*[A-Za-z_0-9]+\.py):(\d+)(?:-(\d+))?")
    checked, bad = 0, []
    for f in files:
        if os.sep + "build" + os.sep in f or not os.path.isfile(f):
            continue
        for m in cite.finditer(open(f, errors="ignore").read()):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = [r for r in lengths if r == path or r.endswith("/" + path)]
            if not cands:
                if path in own or any(o.endswith("/" + path) for o in own):
                    continue  # a citation of this repository's own file
                bad.append((os.path.relpath(f, ROOT), m.group(0), "no such file in the reference"))
                continue
            checked += 1
            if not any(a <= b <= lengths[r] for r in cands):
                bad.append((os.path.relpath(f, ROOT), m.group(0), [(r, lengths[r]) for r in cands]))
    assert checked >= 300 and not bad, bad[:20]
