"""Synthetic catalogues for tests and bench: same tree shape the reference's
own tools produce, no reference files needed at run time.

* TDM/JTM tree: items coded by TreeInit's halving rule (tdm/.../tree/TreeInit.scala:204-213,
  right half -> 2c+1, left half -> 2c+2) and sunk to the deepest level by
  TreeBuilder.flattenLeaves (TreeBuilder.scala:133-140) => a sparse last level.
* OTM tree: complete, N leaf ids drawn without replacement from the 2^L slots and sorted
  (otm/.../dataset/LocalDataSet.scala:189-197).
* Queries: T item ids, left-padded with 0 (TreeInit.scala:253), history length 2..T, Zipf items.
"""
from __future__ import annotations

import numpy as np

from .formats.tree_file import TreeFile


def halving_codes(n_items: int) -> np.ndarray:
    """code of sorted item index i under genCode(0, n, 0) -- vectorised, no recursion."""
    start = np.zeros(1, np.int64)
    end = np.full(1, n_items, np.int64)
    code = np.zeros(1, np.int64)
    out = np.empty(n_items, np.int64)
    while len(start):
        size = end - start
        done = size == 1
        out[start[done]] = code[done]
        keep = size > 1
        s, e, c = start[keep], end[keep], code[keep]
        mid = (s + e) >> 1
        start = np.concatenate([mid, s])
        end = np.concatenate([e, mid])
        code = np.concatenate([2 * c + 1, 2 * c + 2])
    return out


def tdm_tree(n_items: int, seed: int = 1) -> TreeFile:
    """Item ids 1..N shuffled over the sorted positions, TreeBuilder.build layout."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = rng.permutation(n_items).astype(np.int64) + 1
    codes = halving_codes(n_items)
    max_level = int(np.floor(np.log2(codes.max() + 1)))
    min_leaf = (1 << max_level) - 1
    leaf_codes = codes.copy()
    while True:                                   # flattenLeaves: sink(code) = 2*code+1 until >= minCode
        m = leaf_codes < min_leaf
        if not m.any():
            break
        leaf_codes[m] = leaf_codes[m] * 2 + 1
    order = np.argsort(leaf_codes, kind="stable")
    ids, leaf_codes = ids[order], leaf_codes[order]
    anc = []
    cur = leaf_codes
    for _ in range(max_level):                    # leaf_codes is sorted, so every parent list is too: unique by diff, no sort
        par = (cur - 1) >> 1
        keep = np.ones(len(par), bool)
        keep[1:] = par[1:] != par[:-1]
        cur = par[keep]
        anc.append(cur)
    anc = np.concatenate(anc) if anc else np.zeros(0, np.int64)
    offset = int(ids.max()) + 1
    codes_all = np.concatenate([leaf_codes, anc])
    node_ids = np.concatenate([ids, anc + offset])
    is_leaf = np.concatenate([np.ones(len(ids), np.uint8), np.zeros(len(anc), np.uint8)])
    return TreeFile(max_level, codes_all.astype(np.int32), node_ids.astype(np.int32), is_leaf,
                    np.ones(len(codes_all), np.float32), ids.astype(np.int32), leaf_codes.astype(np.int32))


def otm_mapping(n_items: int, seed: int = 42):
    """-> (item_ids 1..N, sorted leaf ids on level ceil(log2 N))"""
    rng = np.random.Generator(np.random.PCG64(seed))
    leaf_level = int(np.ceil(np.log(n_items) / np.log(2)))
    leaf_start = (1 << leaf_level) - 1
    slots = np.sort(rng.choice(1 << leaf_level, size=n_items, replace=False))
    return np.arange(1, n_items + 1, dtype=np.int32), (slots + leaf_start).astype(np.int32), leaf_level


def din_params(rows: int, E: int, seed: int = 2, dtype=np.float32, structured: bool = True) -> np.ndarray:
    """Compact DIN vector with N(0, 0.05^2) entries, biases 0 (EmbeddingShare.scala:21, Linear.scala:12-13).
    structured=True adds a rank-8 component to the table so sibling scores differ by far more than
    rounding noise (a "trained-like" table)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    emb = rng.normal(0.0, 0.05, size=(rows, E))
    if structured:
        u = rng.normal(0.0, 1.0, size=(rows, 8))
        v = rng.normal(0.0, 0.05, size=(8, E))
        emb = emb + u @ v
    watt = rng.normal(0.0, 0.05 if not structured else 0.2, size=(E, E))
    w1 = rng.normal(0.0, 0.05 if not structured else 0.2, size=(E, 2 * E))
    b1 = np.zeros(E) if not structured else rng.normal(0.0, 0.05, size=E)
    w2 = rng.normal(0.0, 0.05 if not structured else 0.3, size=E)
    b2 = np.zeros(1)
    return np.concatenate([emb.ravel(), watt.ravel(), w1.ravel(), b1, w2, b2]).astype(dtype)


def queries(n_users: int, T: int, n_items: int, seed: int = 4, zipf_a: float = 1.1) -> np.ndarray:
    """B x T item ids in 1..n_items, left-padded with 0; u in [2, T] real entries per user."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((n_users, T), np.int32)
    lens = rng.integers(2, T + 1, size=n_users)
    z = rng.zipf(zipf_a, size=(n_users, T))
    items = ((z - 1) % n_items + 1).astype(np.int32)
    for u in range(n_users):
        out[u, T - lens[u]:] = items[u, :lens[u]]
    return out


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def dr_synthetic_path_csr(num_item: int, K: int, D: int, J: int, seed: int):
    """Host mirror of dmg_dr_init_synthetic's path assignment (csrc/dr.cu): key(item, j) = splitmix64(seed' ^ splitmix64(item J + j))
    mod K^D with seed' = seed ^ splitmix64(6); a path keeps its largest item.  -> (path_off[K^D + 1], path_items)"""
    with np.errstate(over="ignore"):
        n_keys = K ** D
        sd = np.uint64(seed) ^ _splitmix64(np.array([6], np.uint64))[0]
        q = np.arange(num_item * J, dtype=np.uint64)
        key = (_splitmix64(sd ^ _splitmix64(q)) % np.uint64(n_keys)).astype(np.int64)
    winner = np.full(n_keys, -1, np.int64)
    np.maximum.at(winner, key, (q // np.uint64(J)).astype(np.int64))
    occ = winner >= 0
    off = np.zeros(n_keys + 1, np.int64)
    off[1:] = np.cumsum(occ)
    return off, winner[occ].astype(np.int32)
