"""On-disk artefacts the Scala tasks and this engine exchange.

* KV-protobuf tree file — written by TreeBuilder.build
  (tdm/src/main/scala/com/mass/tdm/tree/TreeBuilder.scala:23-101) and
  JTMTree.writeTree (jtm/src/main/scala/com/mass/jtm/tree/JTMTree.scala:115-182),
  read by DistTree.loadData/loadItems (tdm/.../tree/DistTree.scala:25-87).
  Record = big-endian int32 length + ``KVItem{key: bytes, value: bytes}``.
  key ``"tree_meta"`` -> TreeMeta, key ``"Part_<n>"`` -> IdCodePart, any other
  key is the decimal node code -> Node{id, probality, leaf_cate_id, is_leaf}.
* OTM mapping text (``item leafId`` per line) — Serialization.saveMapping /
  loadMapping (tdm/.../utils/Serialization.scala:103-120).
* DR mapping — MappingOp.writeMapping/loadMapping
  (deep-retrieval/.../model/MappingOp.scala:45-94): int32 BE length + ItemSet.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Tuple

import math

import numpy as np

from . import pbwire as pb


@dataclass
class TreeFile:
    """Flat arrays equivalent to DistTree's maps (one entry per stored node)."""
    max_level: int
    codes: np.ndarray      # int32 [n_nodes]   node code (key)
    node_ids: np.ndarray   # int32 [n_nodes]   Node.id (item id for leaves, code+offset for ancestors)
    is_leaf: np.ndarray    # uint8 [n_nodes]
    prob: np.ndarray       # float32 [n_nodes] Node.probality
    leaf_ids: np.ndarray   # int32 [n_items]   from the Part_* id/code lists, file order
    leaf_codes: np.ndarray  # int32 [n_items]

    @property
    def non_leaf_offset(self) -> int:      # DistTree.scala:35
        return int(self.leaf_ids.max()) + 1

    @property
    def max_code(self) -> int:             # DistTree.scala:36
        return int(self.leaf_codes.max())


def _records(data: bytes):
    p = 0
    n = len(data)
    while p + 4 <= n:
        (ln,) = struct.unpack(">i", data[p:p + 4])
        p += 4
        rec = data[p:p + ln]
        if len(rec) != ln:
            raise ValueError("truncated KV record")
        p += ln
        yield rec


def read_tree(path: str) -> TreeFile:
    with open(path, "rb") as f:
        data = f.read()
    codes, ids, leaf, prob = [], [], [], []
    leaf_ids: List[int] = []
    leaf_codes: List[int] = []
    max_level = None
    for rec in _records(data):
        kv = pb.decode(rec)
        key = kv[1][0].decode()
        val = kv.get(2, [b""])[0]
        if key.startswith("tree_meta"):
            meta = pb.decode(val)
            max_level = pb.to_int32(meta.get(1, [0])[0])
        elif key.startswith("Part_"):
            part = pb.decode(val)
            for pair in part.get(2, []):
                m = pb.decode(pair)
                leaf_ids.append(pb.to_int32(m.get(1, [0])[0]))
                leaf_codes.append(pb.to_int32(m.get(2, [0])[0]))
        else:
            node = pb.decode(val)
            codes.append(int(key))
            ids.append(pb.to_int32(node.get(1, [0])[0]))
            prob.append(struct.unpack("<f", node[2][0])[0] if 2 in node else 0.0)
            leaf.append(1 if node.get(4, [0])[0] else 0)
    if max_level is None:
        raise ValueError("tree file has no tree_meta record")
    return TreeFile(max_level, np.array(codes, np.int32), np.array(ids, np.int32),
                    np.array(leaf, np.uint8), np.array(prob, np.float32),
                    np.array(leaf_ids, np.int32), np.array(leaf_codes, np.int32))


def _kv(key: str, value: bytes) -> bytes:
    msg = pb.field_bytes(1, key.encode()) + pb.field_bytes(2, value)
    return struct.pack(">i", len(msg)) + msg


def _node(node_id: int, prob: float, is_leaf: bool) -> bytes:
    out = b""
    if node_id:
        out += pb.field_varint(1, node_id)
    if prob != 0.0:
        out += pb.field_float(2, prob)
    if is_leaf:
        out += pb.field_varint(4, 1)
    return out


def write_tree(path: str, leaf_ids, leaf_codes, max_level: int, leaf_prob=None,
               non_leaf_offset: int | None = None, stat: Dict[int, int] | None = None) -> None:
    """Emit the same record sequence as TreeBuilder.build (TreeBuilder.scala:24-96) / JTMTree.writeTree: leaves sorted by
    code (`sortBy(_.code)`, stable), per leaf the leaf node then its not-yet-written ancestors, then the Part_* chunks of
    512 pairs, then tree_meta.  `stat` = TreeBuilder's `stat: Option[Map[Int, Int]]` (item id -> count): a leaf's probability is
    its count (1.0 when the id is missing), an ancestor's the Float sum of the counts below it (`computeNodeOccurrence`, only ids
    present in `stat` contribute; an ancestor nobody contributes to gets 1.0).  Without `stat` every probability is 1.0
    (`pstat.getOrElse(ancCode, 1.0f)` over an empty map).  `leaf_prob` is the array form of `stat` with every leaf present."""
    if stat is not None:
        ids_l = np.asarray(leaf_ids, np.int64).tolist()
        leaf_prob = np.array([float(stat.get(i, 1)) for i in ids_l], np.float32)
        in_stat = np.array([i in stat for i in ids_l], bool)
    else:
        in_stat = np.ones(len(leaf_ids), bool)
    leaf_ids = np.asarray(leaf_ids, np.int64)
    leaf_codes = np.asarray(leaf_codes, np.int64)
    order = np.argsort(leaf_codes, kind="stable")
    leaf_ids, leaf_codes = leaf_ids[order], leaf_codes[order]
    pstat: Dict[int, float] = {}
    if leaf_prob is None:
        leaf_prob = np.ones(len(leaf_ids), np.float32)
    else:
        leaf_prob = np.asarray(leaf_prob, np.float32)[order]
        for c, pr, ins in zip(leaf_codes.tolist(), leaf_prob.tolist(), in_stat[order].tolist()):
            if not ins:
                continue
            a = c
            for _ in range(max_level):
                a = (a - 1) // 2
                pstat[a] = float(np.float32(pstat.get(a, 0.0) + pr))
    offset = int(max(0, leaf_ids.max()) + 1) if non_leaf_offset is None else non_leaf_offset
    saved = set()
    parts: List[Tuple[str, bytes]] = []
    tmp = b""
    ntmp = 0
    n = len(leaf_ids)
    with open(path, "wb") as f:
        for i in range(n):
            iid, code = int(leaf_ids[i]), int(leaf_codes[i])
            f.write(_kv(str(code), _node(iid, float(leaf_prob[i]), True)))
            pair = b""
            if iid:
                pair += pb.field_varint(1, iid)
            if code:
                pair += pb.field_varint(2, code)
            tmp += pb.field_bytes(2, pair)
            ntmp += 1
            if i == n - 1 or ntmp == 512:
                pid = f"Part_{len(parts) + 1}"
                parts.append((pid, pb.field_bytes(1, pid.encode()) + tmp))
                tmp, ntmp = b"", 0
            a = code
            for _ in range(max_level):
                a = (a - 1) // 2
                if a not in saved:
                    f.write(_kv(str(a), _node(a + offset, pstat.get(a, 1.0), False)))
                    saved.add(a)
        for pid, body in parts:
            f.write(_kv(pid, body))
        meta = pb.field_varint(1, max_level) + b"".join(pb.field_bytes(2, pid.encode()) for pid, _ in parts)
        f.write(_kv("tree_meta", meta))


def flatten_leaves(codes, min_code: int) -> np.ndarray:
    """TreeBuilder.flattenLeaves (TreeBuilder.scala:131-139): sink every code to the leaf level (code * 2 + 1 until >= minCode)."""
    c = np.asarray(codes, np.int64).copy()
    while (c < min_code).any():
        c = np.where(c < min_code, c * 2 + 1, c)
    return c


def build_tree(path: str, tree_ids, tree_codes, stat: Dict[int, int] | None = None) -> Tuple[np.ndarray, int]:
    """TreeBuilder.build (TreeBuilder.scala:24-96): offset = max id + 1, maxLevel = floor(log2(max code + 1)), leaves flattened to
    that level, records written by write_tree.  -> (leaf codes, maxLevel)"""
    ids = np.asarray(tree_ids, np.int64)
    codes = np.asarray(tree_codes, np.int64)
    max_level = int(math.floor(math.log(int(codes.max()) + 1) / math.log(2)))
    leaf_codes = flatten_leaves(codes, (1 << max_level) - 1)
    write_tree(path, ids, leaf_codes, max_level, stat=stat)
    return leaf_codes, max_level


def read_otm_mapping(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """-> (item_ids, leaf_ids), file order."""
    items, leaves = [], []
    with open(path) as f:
        for line in f:
            kv = line.split()
            if not kv:
                continue
            items.append(int(kv[0]))
            leaves.append(int(kv[-1]))
    return np.array(items, np.int32), np.array(leaves, np.int32)


def write_otm_mapping(path: str, item_ids, leaf_ids) -> None:
    with open(path, "w") as f:
        for a, b in zip(item_ids, leaf_ids):
            f.write(f"{int(a)} {int(b)}\n")


def read_dr_mapping(path: str):
    """-> (items int32[n], ids int32[n], paths int32[n, J, D]) in file order."""
    with open(path, "rb") as f:
        data = f.read()
    (ln,) = struct.unpack(">i", data[:4])
    msg = pb.decode(data[4:4 + ln])
    items, ids, paths = [], [], []
    for raw in msg.get(1, []):
        it = pb.decode(raw)
        items.append(pb.to_int32(it.get(1, [0])[0]))
        ids.append(pb.to_int32(it.get(2, [0])[0]))
        pp = []
        for praw in it.get(3, []):
            pm = pb.decode(praw)
            idx: List[int] = []
            for v in pm.get(1, []):
                if isinstance(v, (bytes, bytearray)):
                    idx.extend(pb.packed_varints(v))
                else:
                    idx.append(pb.to_int32(v))
            pp.append(idx)
        paths.append(pp)
    return np.array(items, np.int32), np.array(ids, np.int32), np.array(paths, np.int32)


def write_dr_mapping(path: str, items, ids, paths) -> None:
    body = b""
    for item, iid, pp in zip(items, ids, paths):
        m = b""
        if int(item):
            m += pb.field_varint(1, int(item))
        if int(iid):
            m += pb.field_varint(2, int(iid))
        for p in pp:
            packed = b"".join(pb.write_varint(int(v)) for v in p)
            m += pb.field_bytes(3, pb.field_bytes(1, packed) if len(p) else b"")
        body += pb.field_bytes(1, m)
    with open(path, "wb") as f:
        f.write(struct.pack(">i", len(body)))
        f.write(body)
