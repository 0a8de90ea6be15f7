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


def _improve(h: np.ndarray) -> np.ndarray:
    """scala.collection.Hashing.improve on 32-bit ints (the hash every Scala 2.13 immutable.HashMap mixes keys with)."""
    h = h.astype(np.uint32)
    h = (h + ~(h << np.uint32(9))).astype(np.uint32)
    h = h ^ (h >> np.uint32(14))
    h = (h + (h << np.uint32(4))).astype(np.uint32)
    return h ^ (h >> np.uint32(10))


def champ_order(keys) -> np.ndarray:
    """Int keys in the iteration order of a Scala 2.13 immutable.HashMap (the reference builds with Scala 2.13.8, build.sbt:4).

    The map is a CHAMP trie over improve(key.##) (Int: the value itself), 5 bits per level from the least significant end;
    the structure is canonical (a sub-tree with one entry is inlined into its parent), a node's iterator yields its own
    entries in ascending slot order and then its child nodes in ascending slot order (ChampBaseIterator), depth first.
    Hence the order is lexicographic in (continues below this level?, 5-bit digit) per level.  Known answer:
    (1 to 10).toMap iterates 5, 10, 1, 6, 9, 2, 7, 3, 8, 4.
    """
    keys = np.asarray(keys, np.int64)
    h = _improve(keys.astype(np.int32).view(np.uint32)).astype(np.uint64)
    sortkey = np.zeros(len(keys), np.uint64)
    alive = np.ones(len(keys), bool)
    for lvl in range(7):
        idx = np.flatnonzero(alive)
        if len(idx) == 0:
            break
        prefix = h[idx] & np.uint64((1 << min(5 * (lvl + 1), 32)) - 1)
        _, inv, cnt = np.unique(prefix, return_inverse=True, return_counts=True)
        single = (cnt[inv] == 1) | (lvl == 6)                  # alone under this prefix: an entry of the node at this level
        digit = (h[idx] >> np.uint64(5 * lvl)) & np.uint64(31)
        sortkey[idx] |= ((~single).astype(np.uint64) * np.uint64(32) + digit) << np.uint64(6 * (6 - lvl))
        alive[idx[single]] = False
    return keys[np.argsort(sortkey, kind="stable")]


def build_path_csr(ids, paths, K: int, keep_all_items: bool = False):
    """MappingOp.pathToItems as a CSR over path keys sum_d c_d K^(D-1-d).

    Default = the reference: `itemPathMapping.flatMap { case (item, paths) => paths.map((_, item)) }` builds a
    Map[Path, Int], so a path shared by several items keeps ONE of them -- the last one the flatMap visits, i.e. the last in
    the immutable.HashMap's iteration order over the ids (MappingOp.scala:23-28; `champ_order` reproduces that order).
    keep_all_items=True lists every item of a path in ascending id order instead (what the paper's structure means; not what
    the reference code does).
    """
    ids = np.asarray(ids, np.int64)
    paths = np.asarray(paths, np.int64)                     # [n, J, D]
    n, J, D = paths.shape
    keys = np.zeros((n, J), np.int64)
    for d in range(D):
        keys = keys * K + paths[:, :, d]
    flat_keys = keys.ravel()
    flat_items = np.repeat(ids, J)
    if keep_all_items:
        order = np.lexsort((flat_items, flat_keys))
        flat_keys, flat_items = flat_keys[order], flat_items[order]
        keep = np.ones(len(flat_keys), bool)
        keep[1:] = (flat_keys[1:] != flat_keys[:-1]) | (flat_items[1:] != flat_items[:-1])   # an item listed once per path
        flat_keys, flat_items = flat_keys[keep], flat_items[keep]
    else:
        visit = np.empty(int(ids.max()) + 1 if len(ids) else 0, np.int64)
        visit[champ_order(ids)] = np.arange(len(ids))       # position of every id in the HashMap's iteration
        order = np.lexsort((visit[flat_items], flat_keys))
        flat_keys, flat_items = flat_keys[order], flat_items[order]
        last = np.ones(len(flat_keys), bool)
        last[:-1] = flat_keys[1:] != flat_keys[:-1]          # the last visitor of a path wins
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

    def load_mapping(self, path: str, keep_all_items: bool = False):
        items, ids, paths = tree_file.read_dr_mapping(path)
        return self.set_mapping(items, ids, paths, keep_all_items)

    def set_mapping(self, items, ids, paths, keep_all_items: bool = False):
        self.item_id_mapping = {int(a): int(b) for a, b in zip(items, ids)}
        self.id_item_mapping = {v: k for k, v in self.item_id_mapping.items()}
        off, flat = build_path_csr(ids, paths, self.shape[1], keep_all_items)
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

    def get_parameters(self):
        """LayerModel / RerankModel.getParameters + the softmax weights, as a dict of arrays (DeepRetrieval.saveModel reads these)"""
        return self.engine.dr_download()


class LocalOptimizer:
    """deep-retrieval/.../optim/LocalOptimizer.scala:18-120: the mini-batch loop of one epoch over (sequence, target) samples already
    mapped to item indices.  Every iteration is ONE engine call (dmg_dr_train_step: layer model, then -- while epoch <=
    reRankStoppingEpoch -- the rerank model); shuffling, epochs and evaluation stay with the caller as in the Scala class."""

    def __init__(self, model: DeepRetrieval, item_paths, learning_rate: float, num_sampled: int, batch_size: int,
                 re_rank_epoch: Optional[int] = None, num_thread: int = 1, seed: int = 0):
        self.model, self.lr, self.num_sampled, self.batch_size = model, learning_rate, num_sampled, batch_size
        self.re_rank_epoch, self.num_thread, self.seed = re_rank_epoch, num_thread, seed
        self.layer_t = self.rerank_t = 0                      # Adam's trainCounter / ParameterOptimizer.timestep
        model.engine.dr_load_item_paths(item_paths)           # dataset.itemPathMapping

    def set_item_paths(self, item_paths):
        """after an M-step (CoordinateDescent.optimize) the items' paths change"""
        self.model.engine.dr_load_item_paths(item_paths)

    def train_epoch(self, epoch: int, sequences, targets):
        """-> list of (layer losses [D], rerank loss) per mini-batch"""
        seqs = np.asarray(sequences, np.int32)
        tg = np.asarray(targets, np.int32)
        out = []
        train_rerank = self.re_rank_epoch is None or epoch <= self.re_rank_epoch
        for b0 in range(0, len(seqs), self.batch_size):
            self.layer_t += 1
            if train_rerank:
                self.rerank_t += 1
            out.append(self.model.engine.dr_train_step(seqs[b0:b0 + self.batch_size], tg[b0:b0 + self.batch_size], self.lr, self.layer_t,
                                                       rerank_step_t=self.rerank_t if train_rerank else 0, num_sampled=self.num_sampled,
                                                       seed=self.seed + self.layer_t, parallelism=self.num_thread))
        return out
