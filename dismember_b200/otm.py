"""Host-side mirror of the reference's OTM retrieval API on top of the C ABI.

otm/src/main/scala/com/mass/otm/model/OTM.scala:14-23 (``recommend``) and
CandidateSearcher.batchBeamSearch (CandidateSearcher.scala:15-56).  The scorer is
DeepModel[Double]; the tree is complete with leafLevel = upperLog2(#items).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._capi import Engine
from .formats import tree_file


def upper_log2(n: int) -> int:               # otm/package.scala:17
    return int(math.ceil(math.log(n) / math.log(2)))


class OTM:
    def __init__(self, engine: Optional[Engine] = None, device: int = 0, model_name: str = "din"):
        name = model_name.lower()
        if name not in ("din", "deepfm"):
            raise ValueError("DeepModel should be `DIN` or `DeepFM`")                     # OTM.scala:54-58
        self.engine = engine or Engine(device)
        self.model_name = name
        self.use_mask = name == "din"                                                     # OTM.scala:33 (DeepFM takes no mask)
        self.item_id_mapping: Dict[int, int] = {}
        self.leaf_level = 0

    def load_mapping(self, mapping_path: str) -> "OTM":
        items, leaves = tree_file.read_otm_mapping(mapping_path)
        return self.set_mapping(items, leaves)

    def set_mapping(self, item_ids, leaf_ids) -> "OTM":
        item_ids = np.asarray(item_ids, np.int32)
        leaf_ids = np.asarray(leaf_ids, np.int32)
        self.item_id_mapping = {int(a): int(b) for a, b in zip(item_ids, leaf_ids)}
        self.leaf_level = upper_log2(len(self.item_id_mapping))                           # OTM.scala:12
        self.engine.load_tree_complete(self.leaf_level, item_ids, leaf_ids)
        return self

    def set_parameters(self, params: np.ndarray, embed_size: int, seq_len: int) -> "OTM":
        rows = (1 << (self.leaf_level + 1)) - 1
        if self.model_name == "deepfm":                                                   # otm/.../model/DeepFM.scala:12-48
            self.engine.load_deepfm_weights(np.asarray(params, np.float64), rows, embed_size, seq_len)
        else:
            self.engine.load_din_weights(np.asarray(params, np.float64), rows, embed_size, seq_len)
        return self

    def sequence_ids(self, sequences) -> np.ndarray:
        """sequence.map(itemIdMapping.getOrElse(_, paddingIdx))  (OTM.scala:15)"""
        seqs = np.asarray(sequences, np.int64)
        get = self.item_id_mapping.get
        return np.array([[get(int(x), -1) for x in row] for row in seqs.reshape(-1, seqs.shape[-1])], np.int32)

    def recommend(self, sequence: Sequence[int], topk: int, beam_size: int) -> List[Tuple[int, float]]:
        ids = self.sequence_ids(np.asarray(sequence)[None])
        items, scores, counts = self.engine.otm_retrieve(ids, beam_size, topk, self.use_mask)
        n = int(counts[0])
        prob = 1.0 / (1.0 + np.exp(-scores[0, :n]))                                       # OTM.sigmoid
        return list(zip(items[0, :n].tolist(), prob.tolist()))

    def batch_beam_search(self, sequences_leaf_ids, beam_size: int):
        """-> (ids[B, 2*beam], scores[B, 2*beam], counts[B]); sequences already mapped to leaf ids."""
        return self.engine.otm_beam_search(np.asarray(sequences_leaf_ids, np.int32), beam_size, self.use_mask)

    def recommend_batch(self, sequences, topk: int, beam_size: int):
        return self.engine.otm_retrieve(self.sequence_ids(sequences), beam_size, topk, self.use_mask)
