"""Host-side mirror of the OTM training loop on top of the C ABI.

otm/src/main/scala/com/mass/otm/optim/LocalOptimizer.scala:55-140 with
OTMTree.optimalPseudoTargets / normalTargets / beamSearchNodes (otm/.../tree/OTMTree.scala:27-212)
and MiniBatch.batchTransform (otm/.../dataset/MiniBatch.scala:16-39).  Every scorer call
(`model.forward`), the beam search and the fwd/bwd/Adam step run in the CUDA engine; this file
holds only the per-level bookkeeping the Scala driver does with Lists.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from ._capi import Engine


def lower_log2(n: int) -> int:
    return int(n).bit_length() - 1


class OTMTrainer:
    def __init__(self, engine: Engine, leaf_level: int, beam_size: int, seq_len: int, use_mask: bool = True):
        self.e = engine
        self.leaf_level = leaf_level
        self.beam = beam_size
        self.start_level = lower_log2(beam_size)                     # LocalOptimizer.scala:37
        self.T = seq_len
        self.use_mask = use_mask
        self.step_t = 0

    # ---- model.forward on (nodes, per-row sequences) --------------------------------------------
    def _forward(self, nodes: np.ndarray, seqs: np.ndarray) -> np.ndarray:
        mask = np.flatnonzero((seqs == -1).ravel()).astype(np.int32) if self.use_mask else None
        return self.e.score_pairs(nodes, seqs, mask)

    # ---- OTMTree.normalTargets :50-63 : ancestors of the target leaves, score 1 -------------------
    def normal_targets(self, target_leaves: Sequence[Sequence[int]]) -> List[List[Dict[int, float]]]:
        levels = []
        cur = [list(t) for t in target_leaves]
        per_level = []
        for _ in range(self.leaf_level, self.start_level, -1):
            per_level.append([{int(i): 1.0 for i in items} for items in cur])
            cur = [[(i - 1) >> 1 for i in items] for items in cur]
        levels = per_level[::-1]                                     # ascending levels start+1 .. leaf
        return levels

    # ---- OTMTree.optimalPseudoTargets :27-46 + computeTargets :104-129 ----------------------------
    def optimal_pseudo_targets(self, seqs: np.ndarray, target_leaves: Sequence[Sequence[int]]):
        """seqs: B x T leaf ids (-1 pad).  Returns per level (ascending) a list over users of {node id: pseudo target}.
        Computed on the device (dmg_otm_pseudo_targets: expand / model.forward / combine kernels per level)."""
        off = np.zeros(len(target_leaves) + 1, np.int64)
        off[1:] = np.cumsum([len(t) for t in target_leaves])
        flat = np.concatenate([np.asarray(t, np.int32) for t in target_leaves]) if off[-1] else np.zeros(0, np.int32)
        ids, vals, cnt = self.e.otm_pseudo_targets(seqs, off, flat, self.leaf_level, self.start_level, self.use_mask)
        return [[{int(i): float(v) for i, v in zip(ids[li, u, :cnt[li, u]], vals[li, u, :cnt[li, u]])} for u in range(len(seqs))]
                for li in range(ids.shape[0])]

    def optimal_pseudo_targets_host(self, seqs: np.ndarray, target_leaves: Sequence[Sequence[int]]):
        """The same with the Scala driver's List / Map bookkeeping on the host around model.forward (readable mirror, used by the
        tests as a third restatement).  Insertion order of the dicts = order of first appearance."""
        B = len(seqs)
        level_nodes = [{int(i): 1.0 for i in t} for t in target_leaves]     # leaf level: Node(_, 1.0)
        out = [level_nodes]
        for _ in range(self.leaf_level - 1, self.start_level, -1):
            children = out[0]
            pos, neg, rows_user = [], [], []
            for u in range(B):
                for n in children[u]:
                    pos.append(n)
                    neg.append(n - 1 if n % 2 == 0 else n + 1)           # OTMTree.scala:143
                    rows_user.append(u)
            pos = np.array(pos, np.int32); neg = np.array(neg, np.int32)
            rows_seq = seqs[np.array(rows_user, np.int64)] if len(rows_user) else np.zeros((0, self.T), np.int32)
            # quirk kept: without a mask the reference scores the NEGATIVE tensor twice (OTMTree.scala:157-161)
            pos_pred = self._forward(pos if self.use_mask else neg, rows_seq)
            neg_pred = self._forward(neg, rows_seq)
            parents: List[Dict[int, float]] = []
            k = 0
            for u in range(B):
                acc: Dict[int, float] = {}
                for n, z in children[u].items():
                    sib = int(neg[k])
                    label = z if pos_pred[k] >= neg_pred[k] else children[u].get(sib, 0.0)
                    par = (n - 1) >> 1
                    acc[par] = acc.get(par, 0.0) + label                  # groupMapReduce(_ + _)
                    k += 1
                parents.append({p: max(0.0, min(1.0, v)) for p, v in acc.items()})   # clipValue
            out.insert(0, parents)
        return out

    # ---- OTMTree.beamSearchNodes :67-91 ------------------------------------------------------------
    def beam_search_nodes(self, seqs: np.ndarray):
        return self.e.otm_beam_search_levels(seqs, self.beam, self.leaf_level, self.use_mask)

    # ---- LocalOptimizer.optimize body for ONE mini-batch :62-81 -----------------------------------
    def train_minibatch(self, seqs: np.ndarray, target_leaves, lr: float, target_mode: str = "pseudo") -> List[float]:
        seqs = np.ascontiguousarray(seqs, np.int32)
        targets = self.optimal_pseudo_targets(seqs, target_leaves) if target_mode == "pseudo" \
            else self.normal_targets(target_leaves)
        ids, scores, counts = self.beam_search_nodes(seqs)
        losses = []
        for li in range(ids.shape[1]):
            nodes, rows_seq, labels = [], [], []
            for u in range(len(seqs)):
                c = int(counts[u, li])
                nd = ids[u, li, :c]
                tg = targets[li][u]
                nodes.append(nd)
                rows_seq.append(np.repeat(seqs[u][None], c, 0))
                labels.append(np.array([tg.get(int(n), 0.0) for n in nd]))       # MiniBatch.scala:27-34
            nodes = np.concatenate(nodes); rows_seq = np.concatenate(rows_seq); labels = np.concatenate(labels)
            mask = np.flatnonzero((rows_seq == -1).ravel()).astype(np.int32) if self.use_mask else None
            self.step_t += 1
            losses.append(float(self.e.train_step(nodes, rows_seq, mask, labels, lr, self.step_t)))
        return losses
