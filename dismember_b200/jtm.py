"""Host-side mirror of JTM tree learning on top of the C ABI.

jtm/src/main/scala/com/mass/jtm/optim/JTM.scala:22-74 (level loop with step `gap`),
TreeLearning.getChildrenProjection / sortNodeWeights / reBalance
(jtm/.../optim/TreeLearning.scala:48-97,137-150,217-265).  The scorer work -- every
(item sample, candidate node) forward and the in-order weight sums -- runs in the CUDA engine
(dmg_jtm_item_weights, one call per level step for ALL parents); the per-parent greedy re-balance
is sequential by construction and stays on the host, written exactly like the Scala.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from ._capi import Engine


def stable_desc_order(weights: np.ndarray) -> np.ndarray:
    """sortBy(_._2)(Ordering[Float].reverse): stable, Float.compare total order."""
    w = np.ascontiguousarray(weights, np.float32)
    u = w.view(np.uint32).astype(np.int64)
    u = np.where(np.isnan(w), 0x7fc00000, u)
    key = np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
    return np.argsort(-key, kind="stable")


def re_balance(items: Sequence[int], cand_nodes: np.ndarray, cand_weights: np.ndarray, old_node: Dict[int, int],
               children: Sequence[int], max_assign: int) -> Dict[int, List[int]]:
    """TreeLearning.reBalance.  items[i] has its children sorted by weight desc in
    cand_nodes[i] / cand_weights[i].  Returns child node -> item ids."""
    idx_of = {int(it): i for i, it in enumerate(items)}
    res: Dict[int, List[tuple]] = {}
    for i, it in enumerate(items):                                   # groupMap keeps array order
        res.setdefault(int(cand_nodes[i, 0]), []).append((int(it), float(cand_weights[i, 0]), 1))
    processed = set()
    while True:
        best_cnt, best_node = -1, 0
        for n in children:                                           # getMaxNode: first maximum wins
            cnt = len(res[n]) if (n not in processed and n in res) else -1
            if cnt > best_cnt:
                best_cnt, best_node = cnt, n
        if best_cnt <= max_assign:
            break
        processed.add(best_node)
        lst = res[best_node]
        # sortBy(i => (oldItemNodeMap(i.id) != node, i.weight)) under (Boolean asc, Float desc), stable
        w = np.array([x[1] for x in lst], np.float32)
        moved = np.array([old_node[x[0]] != best_node for x in lst])
        order = stable_desc_order(w)
        order = order[np.argsort(moved[order], kind="stable")]
        lst = [lst[k] for k in order]
        res[best_node] = lst[:max_assign]
        for it, _, nxt in lst[max_assign:]:
            i = idx_of[it]
            k = nxt
            while k < cand_nodes.shape[1]:
                node, weight = int(cand_nodes[i, k]), float(cand_weights[i, k])
                if node not in processed:
                    res.setdefault(node, []).append((it, weight, k + 1))
                    break
                k += 1
    return {n: [x[0] for x in v] for n, v in res.items()}


class JTM:
    def __init__(self, engine: Engine, max_level: int, item_codes: Dict[int, int], item_samples: Dict[int, np.ndarray],
                 gap: int, seq_len: int, hierarchical: bool = False, min_level: int = 0, use_mask: bool = True,
                 native: bool = True):
        self.e, self.max_level, self.gap, self.T = engine, max_level, gap, seq_len
        self.item_codes = item_codes                      # current tree: item id -> leaf code
        self.item_samples = item_samples                  # itemSequenceMap: item -> [n_samples, T] item ids
        self.hier, self.min_level, self.use_mask = hierarchical, min_level, use_mask
        self.native = native                              # reBalance through dmg_jtm_assign_level (False: the Python mirror below)

    def _ancestor_at_level(self, item: int, level: int) -> int:      # JTMTree.getAncestorAtLevel
        lim = (1 << (level + 1)) - 1
        c = self.item_codes[item]
        while c >= lim:
            c = (c - 1) >> 1
        return c

    def level_step(self, projection: Dict[int, int], old_level: int) -> Dict[int, int]:
        level = min(self.max_level, old_level + self.gap)
        items = list(projection.keys())                              # caller's order (Scala: Map order)
        parents = np.array([projection[i] for i in items], np.int32)
        counts = [len(self.item_samples.get(i, ())) for i in items]
        off = np.zeros(len(items) + 1, np.int64)
        off[1:] = np.cumsum(counts)
        seqs = np.concatenate([np.asarray(self.item_samples[i], np.int32).reshape(-1, self.T)
                               for i in items if i in self.item_samples] or [np.zeros((0, self.T), np.int32)])
        w = self.e.jtm_item_weights(off, seqs, parents, old_level, level, self.hier, self.min_level, self.use_mask)
        n_child = w.shape[1]
        max_assign = 1 << (self.max_level - level)
        old = np.array([self._ancestor_at_level(it, level) for it in items], np.int32)
        if self.native:
            nodes = self.e.jtm_assign_level(parents, old, w, max_assign)          # dmg_jtm_assign_level: reBalance in the library
            return {it: int(n) for it, n in zip(items, nodes)}
        order = np.stack([stable_desc_order(w[i]) for i in range(len(items))])
        new_proj = dict(projection)
        by_parent: Dict[int, List[int]] = {}
        for k, it in enumerate(items):
            by_parent.setdefault(int(parents[k]), []).append(k)
        for par, rows in by_parent.items():
            first = (par + 1) * n_child - 1                          # getChildrenAtLevel: left to right
            children = [first + c for c in range(n_child)]
            its = [items[k] for k in rows]
            cn = first + order[rows]
            cw = np.take_along_axis(w[rows], order[rows], 1)
            old_map = {it: int(old[k]) for it, k in zip(its, rows)}
            balanced = re_balance(its, cn, cw, old_map, children, max_assign)
            for node, assigned in balanced.items():
                assert len(assigned) <= max_assign
                for it in assigned:
                    new_proj[it] = node
        return new_proj

    def optimize(self) -> Dict[int, int]:                            # JTM.optimize
        proj = {it: 0 for it in self.item_codes}
        for old_level in range(0, self.max_level, self.gap):
            proj = self.level_step(proj, old_level)
        return proj
