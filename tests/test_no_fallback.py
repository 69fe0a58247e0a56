"""CPU: the product path has no CPU / PyTorch / oracle fallback.  (1) Calling the drop-in MGAT with CPU tensors raises;
(2) lib.load() raises when libisg.so is absent; (3) nothing under the package, and nothing on bench.py's GPU arm,
imports or executes oracle/ (the oracle is test infrastructure: tests/, smoke() and bench.py's CPU legs only)."""
import ast
import os

import pytest
import torch

import isg_b200
from isg_b200 import lib as L
from isg_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "intrinsic-subgraph-generation-for-vqa_b200")


def _model(sampler="imle"):
    from isg_b200.isubgvqa import MGAT

    return MGAT(channels=16, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
                interpretable_mode=False, sampler_type=sampler, sample_k=2, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)


@pytest.mark.parametrize("executor", [True, False])
def test_cpu_tensors_are_rejected_not_computed(executor):
    from isg_b200.isubgvqa import mgat as mgat_mod

    b = synth.make_batch(3, channels=16, mean_nodes=5, mean_edges=12, seed=1)
    prev = mgat_mod._USE_EXECUTOR
    mgat_mod.set_executor(executor)
    try:
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            _model()(b["x"], b["edge_index"], b["instr_vectors"], b["global_language_feats"], b["edge_attr"], b["batch"])
    finally:
        mgat_mod.set_executor(prev)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libisg.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.load()


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom):
            yield node.module or ""


def test_package_never_touches_the_oracle():
    banned = ("isg_oracle", "reference_loader", "stage_reference", "make_golden", "oracle")
    for dirpath, _dirs, files in os.walk(PKG):
        if "build" in dirpath.split(os.sep) or "__pycache__" in dirpath:
            continue
        for f in files:
            p = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for mod in _imports(p):
                    assert not any(mod == b or mod.startswith(b + ".") for b in banned), (p, mod)
                src = open(p).read()
                assert "oracle/" not in src and "oracle\"" not in src and "/root/reference" not in src, p
            elif f.endswith((".cu", ".cuh", ".h")):
                code = "\n".join(line.split("//")[0] for line in open(p).read().splitlines())
                assert "oracle" not in code, p


def test_bench_gpu_arm_does_not_reach_the_oracle():
    """bench.py may execute oracle/ only in its CPU legs (run_cpu_port / run_cpu_reference / cpu_baseline)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    cpu_legs = set()
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            body = ast.get_source_segment(src, node)
            if "oracle" in body.replace(node.name, "") and ("sys.path.insert" in body or "import isg_oracle" in body):
                cpu_legs.add(node.name)
    assert cpu_legs, "expected bench.py to have CPU legs that import the oracle"
    assert all(("cpu" in n or "reference" in n) for n in cpu_legs), cpu_legs
