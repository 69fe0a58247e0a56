"""TEST INFRASTRUCTURE — golden vectors for the scene-graph record conversion.

Runs the UNMODIFIED reference `GQASceneGraphs.query_and_translate` / `convert_one_gqa_scene_graph`
(datasets/scene_graph.py:67-141, 199-389) on seeded synthetic GQA-style records and writes inputs + outputs to
tests/golden/scene_graph_convert.json.  The reference class is built with `object.__new__` (its constructor loads GloVe
and the GQA JSON files, neither available offline) and given a stub vocabulary; the conversion itself is untouched.

    python oracle/make_golden_scene_graphs.py        # needs /root/reference

Records of the committed fixture carry at most ONE distinct attribute per object: the reference orders attribute tokens
by iterating `set(attributes)`, i.e. by string hash, which changes from process to process (PYTHONHASHSEED); records with
several attributes are compared live, in one process, by tests/test_scene_graph_data.py."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "scene_graph_convert.json")

NAMES = ["man", "woman", "helmet", "tree", "car", "dog", "table", "sky", "window", "shirt"]
ATTRS = ["red", "blue", "tall", "wooden", "small", "wet", "open", "striped"]
RELS = ["on", "wearing", "to the left of", "to the right of", "holding", "near", "behind"]


def make_stoi():
    toks = ["<pad>", "<unk>"] + NAMES + ATTRS + RELS + ["<self>", "pokemon"]
    return {t: i for i, t in enumerate(toks)}


def make_records(seed, count, max_attrs):
    """Seeded GQA-style records: numeric-string object ids of mixed length (string sort != numeric sort), names /
    attributes / relations partly outside the vocabulary, self relations, reciprocal and duplicate relations, optional
    boxes; plus the degenerate records the reference special-cases."""
    rng = np.random.default_rng(seed)
    recs = {}
    for i in range(count):
        n = int(rng.integers(1, 13))
        ids = [str(v) for v in rng.choice(np.arange(1, 3000), size=n, replace=False)]
        objs = {}
        for oid in ids:
            k = int(rng.integers(0, max_attrs + 1))
            attrs = [str(rng.choice(ATTRS + ["glowing"]))] * 2 if (max_attrs == 1 and k) else \
                [str(a) for a in rng.choice(ATTRS + ["glowing"], size=k)]
            rels = [{"object": str(rng.choice(ids)), "name": str(rng.choice(RELS + ["orbiting"]))}
                    for _ in range(int(rng.integers(0, 4)))]
            o = {"name": str(rng.choice(NAMES + ["zeppelin"])), "attributes": attrs, "relations": rels}
            if rng.random() < 0.5:
                x1, y1 = int(rng.integers(0, 300)), int(rng.integers(0, 300))
                o.update(x1=x1, y1=y1, x2=x1 + int(rng.integers(1, 200)), y2=y1 + int(rng.integers(1, 200)))
            objs[oid] = o
        recs[f"img{seed}_{i}"] = {"objects": objs}
    recs[f"img{seed}_empty"] = {"objects": {}}                                                    # -> 2-node dummy
    recs[f"img{seed}_lonely"] = {"objects": {"7": {"name": "dog", "attributes": [], "relations": []}}}  # -> 6-node dummy
    recs[f"img{seed}_selfrel"] = {"objects": {"7": {"name": "dog", "attributes": ["wet"],
                                                     "relations": [{"object": "7", "name": "near"}]}}}
    return recs


def reference_store(records, stoi):
    """The reference object, constructor bypassed, on the pure-torch shims."""
    # the file is loaded by PATH under a private module name: oracle/reference_loader.py may have registered a stub
    # under "ISubGVQA.datasets.scene_graph" (its scene-graph-encoder harness) in this process
    import importlib.util

    shim = os.path.join(HERE, "shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    src = os.path.join(os.environ.get("ISG_REFERENCE_SRC", "/root/reference"), "ISubGVQA", "datasets", "scene_graph.py")
    mod = sys.modules.get("isg_ref_datasets_scene_graph")
    if mod is None:
        spec = importlib.util.spec_from_file_location("isg_ref_datasets_scene_graph", src)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["isg_ref_datasets_scene_graph"] = mod
    GQASceneGraphs = mod.GQASceneGraphs

    class _Vocab:
        def get_stoi(self):
            return stoi

    ref = object.__new__(GQASceneGraphs)
    ref.vocab_sg, ref.scene_graphs = _Vocab(), records
    ref.rel_mapping, ref.obj_mapping, ref.attr_mapping = {"near": "on"}, {"car": "tree"}, {"wet": "blue"}
    return ref


def reference_outputs(records, stoi, image_ids):
    ref = reference_store(records, stoi)
    out = {}
    for iid in image_ids:
        d = ref.query_and_translate(iid)
        out[iid] = {k: getattr(d, k) for k in ("x", "edge_index", "edge_attr", "added_sym_edge", "x_bbox")}
    return out


MAPPINGS = dict(rel_mapping={"near": "on"}, obj_mapping={"car": "tree"}, attr_mapping={"wet": "blue"})


def main():
    stoi = make_stoi()
    records = make_records(811, 14, max_attrs=1)
    ids = list(records) + ["img_not_in_the_store"]
    want = reference_outputs(records, stoi, ids)
    blob = {"generator": "oracle/make_golden_scene_graphs.py (unmodified reference datasets/scene_graph.py)",
            "stoi": stoi, "mappings": MAPPINGS, "records": records, "image_ids": ids,
            "expected": {iid: {k: v.tolist() for k, v in d.items()} for iid, d in want.items()}}
    with open(OUT, "w") as f:
        json.dump(blob, f, sort_keys=True, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(ids), "records")


if __name__ == "__main__":
    main()
