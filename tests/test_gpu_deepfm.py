"""DeepFM scorer (the other `model.deep_model`, tdm/.../model/DeepFM.scala:11-44): retrieval and model.forward against
the oracle's restatement, bit for bit."""
import os

import numpy as np
import pytest

from conftest import new_engine

pytestmark = pytest.mark.gpu


def deepfm_params(rows, E, T, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    F = T + 1
    emb = rng.normal(0.0, 0.3, rows * E)
    w1 = rng.normal(0.0, 0.1, F * F * E)
    b1 = rng.normal(0.0, 0.05, F)
    w2 = rng.normal(0.0, 0.3, F)
    return np.concatenate([emb, w1, b1, w2, [0.02]]).astype(np.float32)


@pytest.mark.parametrize("E,n_items,beam", [(16, 3000, 20), (64, 5000, 200), (24, 700, 7)])
def test_deepfm_retrieve_matches_oracle(orc, E, n_items, beam):
    from dismember_b200 import synth
    T, topk, B = 10, 10, 33
    tf = synth.tdm_tree(n_items, seed=3)
    rows = (1 << (tf.max_level + 1)) - 1
    params = deepfm_params(rows, E, T, seed=5)
    seqs = synth.queries(B, T, n_items, seed=6)
    seqs[0] = 0                                               # a user with an empty history
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_deepfm_weights(params, rows, E, T)
    gi, gl, gc = e.tdm_retrieve(seqs, beam, topk)
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T, deepfm=True)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, topk, n_threads=os.cpu_count() or 1)
    assert (gc == oc).all() and (gi == oi).all()
    assert (gl.view(np.uint32) == ol.view(np.uint32)).all()
    # eval variant: consumed items are filtered from the result (Recommender.scala:103-106)
    cons = [list(oi[u, :3][oi[u, :3] >= 0]) for u in range(B)]
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum([len(c) for c in cons])
    flat = np.array([x for c in cons for x in c], np.int32)
    gi2, gl2, gc2 = e.tdm_retrieve(seqs, beam, topk, consumed_off=off, consumed=flat)
    oi2, ol2, oc2 = model.retrieve_batch(tree, seqs, beam, topk, cons_off=off, cons=flat, n_threads=2)
    assert (gi2 == oi2).all() and (gl2.view(np.uint32) == ol2.view(np.uint32)).all() and (gc2 == oc2).all()
    e.close()


def test_deepfm_score_pairs_is_model_forward(orc):
    rows, E, T = 1023, 32, 10
    params = deepfm_params(rows, E, T, seed=9)
    rng = np.random.default_rng(1)
    n = 1000
    node = rng.integers(-1, rows, n).astype(np.int32)
    seq = rng.integers(-1, rows, (n, T)).astype(np.int32)
    e = new_engine()
    e.load_tree_complete(9, np.arange(1, 513, dtype=np.int32), np.arange(511, 1023, dtype=np.int32))
    e.load_deepfm_weights(params, rows, E, T)
    got = e.score_pairs(node, seq)
    want = orc.TdmModel(params, rows, E, T, deepfm=True).forward(node, seq)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    from dismember_b200._capi import DmgIndexError
    with pytest.raises(DmgIndexError):
        e.score_pairs(np.array([rows], np.int32), seq[:1])
    e.close()
