"""GQA scene-graph JSON -> the tensors the hot path consumes, with the per-image CSR built once (SURVEY.md section 8
row f3, the wire-format side of the path).

Reference: `GQASceneGraphs.convert_one_gqa_scene_graph` / `query_and_translate` (datasets/scene_graph.py:199-389,
67-141) turn one image's `{"objects": {id: {name, attributes, relations, x1..y2}}}` record into a PyG `Data`
(`x` [n,4] tokens, `edge_index` [2,e], `edge_attr` [e,1] tokens, `added_sym_edge`, `x_bbox`); `GQADataset.__getitem__`
caches it per image (datasets/gqa.py:169-177) and `gqa_collate` batches with `Batch.from_data_list` (:260).

Here the same record becomes a plain dict of tensors plus a `collate.GraphCsr`, cached per image by `SceneGraphStore`, so
that `collate.collate_scene_graphs` only concatenates (no sort on the GPU, no read-back).  The edge ORDER is part of the
contract — `edge_attr` rows, `added_sym_edge` and every mask the sampler returns are indexed by it — and follows the
reference exactly:

  nodes     object ids sorted as STRINGS ("10" < "2"), node index = rank;
  per node  v ascending: the self loop (v, v) with the "<self>" token first, then each of v's relations (v, t) in the
            record's order, each IMMEDIATELY followed by the reverse edge (t, v) carrying the same relation token iff
            (t, v) is not itself a relation of the record; the positions of those added edges are `added_sym_edge`;
  tokens    x[:, 0] = name, x[:, 1..3] = up to three attributes in the iteration order of `set(attributes)` (the
            reference's own, hash-order dependent), "<pad>" elsewhere; unknown strings map to index 1;
  fallbacks a record without objects becomes the two-node "<unk>" graph (:201-228); a record whose graph has a single
            edge (one object, no relations) or an unknown image id becomes the six-node "<unk>" graph (:68-141).

Host code (numpy): it runs in the data-loader workers, beside the reference's tokenizer."""
import numpy as np
import torch

from .collate import GraphCsr

MAX_OBJ_TOKENS = 4  # 1 name + 3 attributes (datasets/scene_graph.py:272-273)
UNK_INDEX = 1       # `.get(token, 1)` throughout the reference


def _unk_objects(targets):
    return {"objects": {str(i): {"name": "<unk>", "relations": [{"object": str(t), "name": "<unk>"}],
                                 "attributes": ["<unk>"]} for i, t in enumerate(targets)}}


EMPTY_RECORD_2 = _unk_objects([1, 0])              # convert_one_gqa_scene_graph's dummy (:205-228)
EMPTY_RECORD_6 = _unk_objects([1, 0, 3, 1, 5, 3])  # query_and_translate's `empty_sg` (:68-137)


def convert_scene_graph(record, stoi, obj_mapping=None, attr_mapping=None, rel_mapping=None, with_csr=True):
    """One GQA scene-graph record -> dict(x [n,4] int64, edge_index [2,e] int64, edge_attr [e,1] int64,
    added_sym_edge [m] int64, x_bbox [n,4], csr GraphCsr | None).  `stoi`: token -> index of the scene-graph
    vocabulary (needs "<pad>" and "<self>")."""
    obj_mapping, attr_mapping, rel_mapping = obj_mapping or {}, attr_mapping or {}, rel_mapping or {}
    objects = record["objects"]
    if len(objects) == 0:
        objects = EMPTY_RECORD_2["objects"]
    ids = sorted(objects.keys())
    rank = {oid: i for i, oid in enumerate(ids)}
    n = len(ids)
    related = {(rank[oid], rank[rel["object"]]) for oid in ids for rel in objects[oid]["relations"]}

    pad, self_tok = stoi.get("<pad>"), stoi["<self>"]
    x = np.full((n, MAX_OBJ_TOKENS), pad, dtype=np.int64)
    bbox = []
    src, dst, tok, added = [], [], [], []
    for v, oid in enumerate(ids):
        obj = objects[oid]
        x[v, 0] = stoi.get(obj_mapping.get(obj["name"], obj["name"]), UNK_INDEX)
        for slot, attr in enumerate(set(obj["attributes"])):
            if slot >= MAX_OBJ_TOKENS - 1:
                break
            x[v, slot + 1] = stoi.get(attr_mapping.get(attr, attr), UNK_INDEX)
        bbox.append([obj.get(k, -1) for k in ("x1", "y1", "x2", "y2")])
        src.append(v), dst.append(v), tok.append(self_tok)
        for rel in obj["relations"]:
            t = rank[rel["object"]]
            r = stoi.get(rel_mapping.get(rel["name"], rel["name"]), UNK_INDEX)
            src.append(v), dst.append(t), tok.append(r)
            if (t, v) not in related:
                src.append(t), dst.append(v), tok.append(r)
                added.append(len(tok) - 1)
    edge_index = torch.from_numpy(np.array([src, dst], dtype=np.int64))
    out = {
        "x": torch.from_numpy(x),
        "edge_index": edge_index,
        "edge_attr": torch.from_numpy(np.array(tok, dtype=np.int64).reshape(-1, 1)),
        "added_sym_edge": torch.from_numpy(np.array(added, dtype=np.int64)),
        "x_bbox": torch.from_numpy(np.array(bbox)),  # dtype follows the record (int64 for ints / the -1 default)
    }
    out["csr"] = GraphCsr(edge_index, n) if with_csr else None
    return out


class SceneGraphStore:
    """image id -> converted scene graph (+ its CSR), converted on first use and kept: `GQASceneGraphs.query_and_translate`
    (datasets/scene_graph.py:67-141) and `GQADataset.sg_cache` (datasets/gqa.py:169-177) in one place.  `get` returns the
    tensors the way `__getitem__` leaves them: `x` [n,4] and `edge_attr` squeezed to [e] (gqa.py:173-174)."""

    def __init__(self, records, stoi, obj_mapping=None, attr_mapping=None, rel_mapping=None):
        self.records, self.stoi = records, stoi
        self.maps = (obj_mapping or {}, attr_mapping or {}, rel_mapping or {})
        self._cache = {}
        self.hits = self.misses = 0

    def translate(self, image_id):
        g = convert_scene_graph(self.records.get(image_id, EMPTY_RECORD_6), self.stoi, *self.maps)
        if g["edge_index"].size(1) == 1:
            g = convert_scene_graph(EMPTY_RECORD_6, self.stoi, *self.maps)
        return g

    def get(self, image_id):
        g = self._cache.get(image_id)
        if g is not None:
            self.hits += 1
            return g
        self.misses += 1
        g = self.translate(image_id)
        g["x"], g["edge_attr"] = g["x"].squeeze(), g["edge_attr"].squeeze()
        g["image_id"] = image_id
        self._cache[image_id] = g
        return g

    def __len__(self):
        return len(self._cache)
