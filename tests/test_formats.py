"""On-disk formats: writers and readers round-trip without the reference."""
import os

import numpy as np
import pytest

from dismember_b200 import synth
from dismember_b200.formats import javaser, pbwire, tree_file


def test_varint_roundtrip():
    for v in [0, 1, 127, 128, 300, 2 ** 31 - 1, -1, -5]:
        b = pbwire.write_varint(v)
        got, p = pbwire.read_varint(b, 0)
        assert p == len(b) and pbwire.to_int32(got) == v


def test_tree_roundtrip(tmp_path, jtm_fix):
    p = str(tmp_path / "tree.bin")
    tree_file.write_tree(p, jtm_fix["leaf_ids"], jtm_fix["leaf_codes"], int(jtm_fix["max_level"]))
    t = tree_file.read_tree(p)
    assert t.max_level == 12 and len(t.codes) == len(jtm_fix["codes"])
    assert set(t.codes.tolist()) == set(jtm_fix["codes"].tolist())
    assert (t.leaf_ids == jtm_fix["leaf_ids"]).all() and (t.leaf_codes == jtm_fix["leaf_codes"]).all()
    # ancestors carry id = code + nonLeafOffset (TreeBuilder.scala:66-68)
    anc = t.is_leaf == 0
    assert (t.node_ids[anc] == t.codes[anc] + t.non_leaf_offset).all()


def test_tree_writer_record_order_and_probabilities(tmp_path, jtm_fix):
    """TreeBuilder.build (TreeBuilder.scala:24-96): leaves sortBy(_.code) whatever the caller's order, ancestors right after the
    first leaf below them -> the record sequence of the reference's own data/jtm/example_tree.bin; without `stat` every node
    carries 1.0; with `stat` a leaf carries its count and an ancestor the Float sum of the counts of the ids present in stat."""
    p = str(tmp_path / "tree.bin")
    rev = slice(None, None, -1)
    tree_file.write_tree(p, jtm_fix["leaf_ids"][rev], jtm_fix["leaf_codes"][rev], int(jtm_fix["max_level"]))
    t = tree_file.read_tree(p)
    assert (t.codes == jtm_fix["codes"]).all() and (t.node_ids == jtm_fix["node_ids"]).all() and (t.is_leaf == jtm_fix["is_leaf"]).all()
    assert (t.prob == 1.0).all()
    ids = jtm_fix["leaf_ids"].tolist()
    stat = {i: 1 + (i % 5) for i in ids[::2]}                      # every second item has a count
    tree_file.write_tree(p, jtm_fix["leaf_ids"], jtm_fix["leaf_codes"], int(jtm_fix["max_level"]), stat=stat)
    t = tree_file.read_tree(p)
    prob = dict(zip(t.codes.tolist(), t.prob.tolist()))
    want = {}
    for i, c in zip(ids, jtm_fix["leaf_codes"].tolist()):
        assert prob[c] == float(stat.get(i, 1))
        while c > 0 and i in stat:
            c = (c - 1) // 2
            want[c] = want.get(c, 0) + stat[i]
    for c, is_leaf in zip(t.codes.tolist(), t.is_leaf.tolist()):
        if not is_leaf:
            assert prob[c] == float(np.float32(want.get(c, 1)))    # sums stay below 2^24: exact in Float
    assert prob[0] == float(sum(stat.values()))


def test_synthetic_tree_shape():
    t = synth.tdm_tree(1000, seed=3)
    assert t.max_level == 10 and t.is_leaf.sum() == 1000
    lv = np.floor(np.log2(t.codes.astype(np.int64) + 1)).astype(int)
    assert (lv[t.is_leaf == 1] == 10).all()
    assert np.bincount(lv)[9] == 512           # halving rule fills level 9, sinks the rest


def test_mapping_roundtrips(tmp_path, dr_fix, otm_fix):
    p = str(tmp_path / "m.bin")
    tree_file.write_dr_mapping(p, dr_fix["map_items"], dr_fix["map_ids"], dr_fix["map_paths"])
    a, b, c = tree_file.read_dr_mapping(p)
    assert (a == dr_fix["map_items"]).all() and (b == dr_fix["map_ids"]).all() and (c == dr_fix["map_paths"]).all()
    p = str(tmp_path / "m.txt")
    tree_file.write_otm_mapping(p, otm_fix["items"], otm_fix["leaf_ids"])
    i, l = tree_file.read_otm_mapping(p)
    assert (i == otm_fix["items"]).all() and (l == otm_fix["leaf_ids"]).all()


def test_javaser_minimal_stream():
    # hand-built stream: a float[3] array  (TC_ARRAY, classdesc "[F", no fields)
    import struct
    cd = b"\x72" + struct.pack(">H", 2) + b"[F" + b"\x0b\x9c\x81\x89\x22\xe0\x0c\x42" + b"\x02" + struct.pack(">H", 0) + b"\x78\x70"
    data = b"\xac\xed\x00\x05" + b"\x75" + cd + struct.pack(">i", 3) + struct.pack(">fff", 1.0, -2.5, 3.25)
    objs = javaser.load(data)
    assert len(objs) == 1 and objs[0].dtype == np.float32 and objs[0].tolist() == [1.0, -2.5, 3.25]


def test_javaser_inject_weights_round_trip():
    """Weights back into a Java-serialised model: only the payload of the parameter array changes."""
    import struct
    cd = b"\x72" + struct.pack(">H", 2) + b"[F" + b"\x0b\x9c\x81\x89\x22\xe0\x0c\x42" + b"\x02" + struct.pack(">H", 0) + b"\x78\x70"
    first = struct.pack(">ffff", 1.0, -2.5, 3.25, 0.5)
    # two float[4] arrays: parameters, then the gradient buffer (second one through a class-descriptor reference)
    data = (b"\xac\xed\x00\x05" + b"\x75" + cd + struct.pack(">i", 4) + first +
            b"\x75" + b"\x71" + struct.pack(">i", 0x7E0000) + struct.pack(">i", 4) + struct.pack(">ffff", 9, 9, 9, 9))
    spans = javaser.primitive_array_spans(data)
    assert [(t, n) for _, t, n in spans] == [("F", 4), ("F", 4)]
    new = np.array([7.0, 8.0, -1.0, 0.25], np.float32)
    out = javaser.inject_weights(data, new, min_len=4)
    assert len(out) == len(data)
    objs = javaser.load(out)
    assert objs[0].tolist() == new.tolist() and objs[1].tolist() == [9, 9, 9, 9]
    off = spans[0][0]
    assert out[:off] == data[:off] and out[off + 16:] == data[off + 16:]
    with pytest.raises(javaser.JavaSerError):
        javaser.inject_weights(data, np.zeros(3, np.float32), min_len=4)
    with pytest.raises(javaser.JavaSerError):
        javaser.inject_weights(data, np.zeros(4, np.float64), min_len=4)


def test_javaser_inject_into_the_reference_model(jtm_fix):
    """On the reference's own saved model (present in the build container only): inject, reload, compare."""
    path = "/root/reference/data/jtm/example_model.bin"
    if not os.path.exists(path):
        pytest.skip("reference fixture not on this machine")
    data = open(path, "rb").read()
    params = jtm_fix["params"]                                  # the vector tools/make_golden.py took from this file
    new = (params * np.float32(0.5) + np.float32(0.125)).astype(np.float32)
    out = javaser.inject_weights(data, new)
    assert len(out) == len(data)
    arrs = javaser.primitive_arrays(javaser.load(out), 1000)
    assert arrs[0][1].size == new.size and (arrs[0][1] == new).all()
    back = javaser.inject_weights(out, params)
    assert back == data                                          # and back again: byte-identical to the original file


def test_path_to_items_is_the_reference_map_not_a_multimap():
    """MappingOp.pathToItems (MappingOp.scala:23-28) flatMaps a Map into (path -> item) pairs: a path shared by several
    items keeps ONE, the last in the immutable.HashMap's iteration order over the ids (Scala 2.13.8, build.sbt:4)."""
    import os
    import numpy as np
    from dismember_b200.dr import build_path_csr, champ_order
    # known answers of Scala 2.13: (1 to 10).toMap / (0 until 20).toSet iterate in these orders
    assert champ_order(np.arange(1, 11)).tolist() == [5, 10, 1, 6, 9, 2, 7, 3, 8, 4]
    assert champ_order(np.arange(20)).tolist() == [0, 5, 10, 14, 1, 6, 9, 13, 2, 17, 12, 7, 3, 18, 16, 11, 8, 19, 4, 15]
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "dr_fixture.npz"))       # data/dr/example_mapping.bin
    ids, paths, K = f["map_ids"], f["map_paths"], int(f["K"])
    off, flat = build_path_csr(ids, paths, K)
    assert (np.diff(off) <= 1).all()                                  # a Map: at most one item per path
    off_all, flat_all = build_path_csr(ids, paths, K, keep_all_items=True)
    assert (np.diff(off_all) > 0).sum() == (np.diff(off) > 0).sum()   # the same set of paths
    pos = np.empty(ids.max() + 1, np.int64)
    pos[champ_order(ids)] = np.arange(len(ids))
    shared = np.flatnonzero(np.diff(off_all) > 1)
    assert len(shared) > 0                                            # the fixture does have shared paths
    for k in shared[:200]:
        items = flat_all[off_all[k]:off_all[k + 1]]
        assert flat[off[k]] == items[np.argmax(pos[items])]           # the survivor is the last visitor


def test_build_tree_is_treebuilder_build(tmp_path):
    """TreeBuilder.build (TreeBuilder.scala:24-96) on cluster codes of uneven depth: leaves are sunk to the deepest level
    (flattenLeaves), maxLevel = floor(log2(max code + 1)), ancestors carry code + offset."""
    ids = np.array([10, 11, 12, 13, 14], np.int32)
    codes = np.array([3, 9, 10, 5, 6], np.int64)               # depths 2, 3, 3, 2, 2
    p = str(tmp_path / "t.bin")
    leaf_codes, max_level = tree_file.build_tree(p, ids, codes)
    assert max_level == 3 and leaf_codes.tolist() == [7, 9, 10, 11, 13]
    t = tree_file.read_tree(p)
    assert t.max_level == 3 and sorted(t.leaf_codes.tolist()) == [7, 9, 10, 11, 13]
    assert dict(zip(t.leaf_codes.tolist(), t.leaf_ids.tolist())) == {7: 10, 9: 11, 10: 12, 11: 13, 13: 14}
    anc = t.is_leaf == 0
    assert (t.node_ids[anc] == t.codes[anc] + 15).all() and set(t.codes[anc].tolist()) == {0, 1, 2, 3, 4, 5, 6}
