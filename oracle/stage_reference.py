"""TEST / BENCH INFRASTRUCTURE — stages the reference's hot-path Python files, UNMODIFIED, from
/root/reference into the git-ignored oracle/_ref/ so that they travel to the GPU box with the gpurun snapshot
(/root/reference itself does not exist there).  Nothing is copied into the tracked tree and nothing under
isg_b200/ imports it: `bench.py --impl reference` (and the GPU arm's `cpu_baseline` leg) time these files on
the host cores through oracle/reference_loader.py + oracle/shim — that is the reference's own implementation
of the path (kind "reference"), not the port.

    python oracle/stage_reference.py        # idempotent; called by __graft_entry__.build() when the reference is present
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ISG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the files MGAT.forward / backward executes plus the package markers
FILES = [
    "ISubGVQA/__init__.py",
    "ISubGVQA/models/__init__.py",
    "ISubGVQA/models/mgat.py",
    "ISubGVQA/models/mgat_v2_conv.py",
    "ISubGVQA/models/masking.py",
    "ISubGVQA/models/att_pooling.py",
    "ISubGVQA/models/scene_graph_encoder.py",
    "ISubGVQA/sampling/__init__.py",
    "ISubGVQA/sampling/node_edge_masks.py",
    "ISubGVQA/sampling/methods/__init__.py",
    "ISubGVQA/sampling/methods/wrapper.py",
    "ISubGVQA/sampling/methods/aimle.py",
    "ISubGVQA/sampling/methods/noise.py",
    "ISubGVQA/sampling/methods/target.py",
    "ISubGVQA/sampling/methods/target_aimle.py",
    "ISubGVQA/sampling/methods/imle_scheme.py",
    "ISubGVQA/sampling/methods/deterministic_scheme.py",
    "ISubGVQA/sampling/methods/gumbel_scheme.py",
    "ISubGVQA/sampling/methods/simple_scheme.py",
    "ISubGVQA/sampling/methods/simple.py",
    "ISubGVQA/sampling/methods/create_simple_constraint.py",
    "ISubGVQA/sampling/methods/node.py",
    "ISubGVQA/sampling/methods/tensor_utils.py",
    "ISubGVQA/utils/__init__.py",
    "ISubGVQA/utils/scatter_scaled_dot_product.py",
    "ISubGVQA/utils/topk.py",
]


def stage(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "ISubGVQA", "models")):
        return None
    n = 0
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC}\nunmodified copies of {len(FILES)} files; regenerate with oracle/stage_reference.py\n")
    if verbose:
        print(f"staged {len(FILES)} reference files into {DST} ({n} updated)")
    return DST


if __name__ == "__main__":
    if stage(verbose=True) is None:
        sys.exit(f"reference tree not found at {SRC}")
