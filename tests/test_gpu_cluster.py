"""k-means tree rebuild on the GPU (csrc/cluster.cu) against the oracle (oracle/oracle_cluster.c):
tdm/.../cluster/RecursiveCluster.scala:34-214.  Same generator, same summation order => identical codes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,E,iters", [(2, 4, 1), (3, 4, 2), (37, 8, 3), (300, 16, 2), (1024, 8, 2), (1025, 8, 2), (5000, 24, 3), (20000, 64, 2)])
def test_kmeans_tree_matches_oracle(engine, orc, n, E, iters):
    rng = np.random.default_rng(n)
    emb = rng.random((n, E))
    if n == 300:
        emb[40:60] = emb[40]                                   # identical points: zero distances, ties in argPartition
    want = orc.kmeans_tree(emb, iters, 123)
    got = engine.kmeans_tree(emb, iters, 123)
    assert (got == want).all()
    assert len(set(got.tolist())) == n
    assert (engine.kmeans_tree(emb, iters, 124) != want).any() or n < 4


def test_recursive_cluster_mirror_writes_a_loadable_tree(engine, tmp_path):
    """RecursiveCluster(...).run(path) -> TreeBuilder.build -> the file loads as a TDM tree; clustered items share subtrees."""
    from dismember_b200.cluster import RecursiveCluster
    from dismember_b200.formats import tree_file
    rng = np.random.default_rng(5)
    centers = rng.normal(0.0, 4.0, (8, 16))
    emb = np.concatenate([c + rng.normal(0.0, 0.05, (50, 16)) for c in centers])
    ids = np.arange(1, 401, dtype=np.int32)
    p = str(tmp_path / "tree.bin")
    (tmp_path / "emb.csv").write_text("\n".join(f"{i}," + ",".join(repr(float(x)) for x in row) for i, row in zip(ids, emb)))
    rc = RecursiveCluster.from_file(str(tmp_path / "emb.csv"), cluster_iter_num=3, engine=engine, seed=9)
    got_ids, codes = rc.run(p)
    assert (got_ids == ids).all() and len(set(codes.tolist())) == 400
    t = tree_file.read_tree(p)
    assert t.max_level == 9 and len(t.leaf_ids) == 400 and sorted(t.leaf_ids.tolist()) == ids.tolist()
    engine.load_tree_tdm(t.max_level, t.codes, t.node_ids, t.is_leaf, t.leaf_ids, t.leaf_codes)
    # the 8 blobs of 50 points: at depth 3 every blob sits in one subtree (50 = 400 / 8)
    anc = codes.astype(np.int64)
    while (anc > 14).any():
        anc = np.where(anc > 14, (anc - 1) // 2, anc)
    for b in range(8):
        assert len(set(anc[50 * b:50 * (b + 1)].tolist())) == 1
