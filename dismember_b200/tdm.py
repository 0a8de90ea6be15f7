"""Host-side mirror of the reference's TDM/JTM retrieval API on top of the C ABI.

Names and argument meaning follow tdm/src/main/scala/com/mass/tdm/model/TDM.scala:
``TDM.loadTree`` / ``TDM.loadModel`` / ``TDM.recommend(sequence, topk, candidateNum)``
and the batched caller ``Evaluator.evaluate`` (recommendItems with consumed items).
All scoring, beam pruning and sorting happen in the CUDA engine; this file only
moves arrays.  No CPU fallback exists.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._capi import Engine
from .formats import tree_file


def sigmoid(logit) -> np.ndarray:
    """TDM.sigmoid (TDM.scala:56-58): Float logit widened to Double, 1/(1+exp(-x))."""
    x = np.asarray(logit, np.float32).astype(np.float64)
    return 1.0 / (1.0 + np.exp(-x))


class TDM:
    """TDM(dlModel, useMask) bound to one GPU engine."""

    def __init__(self, engine: Optional[Engine] = None, device: int = 0, model_name: str = "din"):
        name = model_name.lower()
        if name not in ("din", "deepfm"):
            raise ValueError("DeepModel name should either be DeepFM or DIN")            # TDM.scala:43
        self.model_name = name
        self.engine = engine or Engine(device)
        self.use_mask = name == "din"                                                     # TDM.scala:27
        self.tree: Optional[tree_file.TreeFile] = None

    # -- TDM.loadTree(treePbPath) / TDMOp.initTree
    def load_tree(self, tree_pb_path: str) -> "TDM":
        return self.set_tree(tree_file.read_tree(tree_pb_path))

    def set_tree(self, tf: tree_file.TreeFile) -> "TDM":
        self.tree = tf
        self.engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        return self

    # -- weights: compact vector of Module.parameters() (Graph.scala:37-48)
    def set_parameters(self, params: np.ndarray, embed_size: int, seq_len: int) -> "TDM":
        rows = (1 << (self.tree.max_level + 1)) - 1                                       # DIN.scala:18, DeepFM.scala:14
        if self.model_name == "deepfm":
            self.engine.load_deepfm_weights(np.asarray(params, np.float32), rows, embed_size, seq_len)
        else:
            self.engine.load_din_weights(np.asarray(params, np.float32), rows, embed_size, seq_len)
        return self

    # -- TDM.recommend(sequence, topk, candidateNum): Array[(Int, Double)]
    def recommend(self, sequence: Sequence[int], topk: int, candidate_num: int) -> List[Tuple[int, float]]:
        items, logits, counts = self.engine.tdm_retrieve(np.asarray(sequence, np.int32)[None], candidate_num, topk,
                                                         self.use_mask)
        n = int(counts[0])
        return list(zip(items[0, :n].tolist(), sigmoid(logits[0, :n]).tolist()))

    # -- Evaluator.evaluate inner loop: recommendItems(seq, ..., Some(consumed)) for a batch of users
    def recommend_items(self, sequences, topk: int, candidate_num: int, consumed: Optional[Sequence[Sequence[int]]] = None):
        seqs = np.asarray(sequences, np.int32)
        if consumed is None:
            items, _, counts = self.engine.tdm_retrieve(seqs, candidate_num, topk, self.use_mask)
        else:
            off = np.zeros(len(consumed) + 1, np.int64)
            off[1:] = np.cumsum([len(c) for c in consumed])
            flat = np.concatenate([np.asarray(c, np.int32) for c in consumed]) if off[-1] else np.zeros(0, np.int32)
            # candidateNum widens per user for every model: max((|consumed| + topk) / 2, candidateNum)  (Recommender.scala:27-33)
            items, _, counts = self.engine.tdm_retrieve(seqs, candidate_num, topk, self.use_mask, off, flat, True)
        return [items[u, :counts[u]].tolist() for u in range(len(seqs))]

    def recommend_batch(self, sequences, topk: int, candidate_num: int):
        """(items[B,topk], logits[B,topk], counts[B]) -- raw arrays for bulk callers."""
        return self.engine.tdm_retrieve(np.asarray(sequences, np.int32), candidate_num, topk, self.use_mask)
