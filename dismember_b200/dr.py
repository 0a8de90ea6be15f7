"""Host-side mirror of the reference's Deep Retrieval API on top of the C ABI.

deep-retrieval/src/main/scala/com/mass/dr/model/DeepRetrieval.scala:26-46 (``recommend``),
MappingOp (MappingOp.scala:14-43): itemIdMapping item -> id, itemPathMapping id -> J paths,
pathItemMapping path -> items.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._capi import Engine
from .formats import tree_file


def build_path_csr(ids, paths, K: int, reference_quirk: bool = False):
    """MappingOp.pathToItems as a CSR over path keys sum_d c_d K^(D-1-d).

    The general form lists every item of a path in ascending item-index order.  The
    reference builds the map with `flatMap` over a Map, which collapses duplicate paths to ONE
    item per path (the last in hash-iteration order, MappingOp.scala:23-28); that order is a JVM
    HashMap artefact that cannot be verified without a JVM, so `reference_quirk=True` keeps one item
    per path (the largest item index) -- documented, not claimed bit-identical to the JVM.
    """
    ids = np.asarray(ids, np.int64)
    paths = np.asarray(paths, np.int64)                     # [n, J, D]
    n, J, D = paths.shape
    keys = np.zeros((n, J), np.int64)
    for d in range(D):
        keys = keys * K + paths[:, :, d]
    flat_keys = keys.ravel()
    flat_items = np.repeat(ids, J)
    order = np.lexsort((flat_items, flat_keys))
    flat_keys, flat_items = flat_keys[order], flat_items[order]
    keep = np.ones(len(flat_keys), bool)
    keep[1:] = (flat_keys[1:] != flat_keys[:-1]) | (flat_items[1:] != flat_items[:-1])   # an item listed once per path
    flat_keys, flat_items = flat_keys[keep], flat_items[keep]
    if reference_quirk:
        last = np.ones(len(flat_keys), bool)
        last[:-1] = flat_keys[1:] != flat_keys[:-1]
        flat_keys, flat_items = flat_keys[last], flat_items[last]
    n_keys = K ** D
    off = np.zeros(n_keys + 1, np.int64)
    np.add.at(off, flat_keys + 1, 1)
    off = np.cumsum(off)
    return off, flat_items.astype(np.int32)


class DeepRetrieval:
    def __init__(self, engine: Optional[Engine] = None, device: int = 0):
        self.engine = engine or Engine(device)
        self.item_id_mapping: Dict[int, int] = {}
        self.id_item_mapping: Dict[int, int] = {}

    def set_model(self, num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b):
        self.shape = (num_item, K, D, T, E)
        self.engine.dr_load(num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
        return self

    def load_mapping(self, path: str, reference_quirk: bool = False):
        items, ids, paths = tree_file.read_dr_mapping(path)
        return self.set_mapping(items, ids, paths, reference_quirk)

    def set_mapping(self, items, ids, paths, reference_quirk: bool = False):
        self.item_id_mapping = {int(a): int(b) for a, b in zip(items, ids)}
        self.id_item_mapping = {v: k for k, v in self.item_id_mapping.items()}
        off, flat = build_path_csr(ids, paths, self.shape[1], reference_quirk)
        self.engine.dr_load_paths(off, flat)
        return self

    def sequence_ids(self, sequences) -> np.ndarray:
        seqs = np.asarray(sequences, np.int64)
        get = self.item_id_mapping.get
        return np.array([[get(int(x), -1) for x in row] for row in seqs.reshape(-1, seqs.shape[-1])], np.int32)

    def recommend(self, sequence: Sequence[int], topk: int, beam_size: int) -> List[Tuple[int, float]]:
        ids = self.sequence_ids(np.asarray(sequence)[None])
        items, scores, counts = self.engine.dr_retrieve(ids, beam_size, topk)
        n = int(counts[0])
        prob = 1.0 / (1.0 + np.exp(-scores[0, :n]))                                       # dr/package.scala:21
        return [(self.id_item_mapping[int(i)], float(p)) for i, p in zip(items[0, :n], prob)]
