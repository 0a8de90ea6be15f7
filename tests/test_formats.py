"""On-disk formats: writers and readers round-trip without the reference."""
import numpy as np

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
