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
    # Recommender.recommendItems widens the beam per user for EVERY model (Recommender.scala:27-33): users with many consumed items
    # start at a deeper level with a wider beam; users with few keep the configured one
    rng = np.random.default_rng(E)
    all_items = tf.leaf_ids
    cons = [rng.choice(all_items, int(k), replace=False).tolist() for k in rng.choice([0, 3, 2 * beam + 5, min(5 * beam, 490)], B)]
    cons[1] = cons[1] + list(oi[1, :4][oi[1, :4] >= 0])
    off[1:] = np.cumsum([len(c) for c in cons])
    flat = np.array([x for c in cons for x in c], np.int32)
    gi3, gl3, gc3 = e.tdm_retrieve(seqs, beam, topk, consumed_off=off, consumed=flat, widen_beam=True)
    oi3, ol3, oc3 = model.retrieve_batch(tree, seqs, beam, topk, cons_off=off, cons=flat, widen_beam=True, n_threads=2)
    assert (gc3 == oc3).all() and (gi3 == oi3).all() and (gl3.view(np.uint32) == ol3.view(np.uint32)).all()
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


# ------------------------------------------------------------------ OTM: DeepModel[Double] = DeepFM
def deepfm_params64(rows, E, T, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    F = T + 1
    return np.concatenate([rng.normal(0.0, 0.3, rows * E), rng.normal(0.0, 0.1, F * F * E), rng.normal(0.0, 0.05, F),
                           rng.normal(0.0, 0.3, F), [0.02]])


@pytest.mark.parametrize("E,leaf_level,beam", [(16, 11, 20), (64, 12, 200), (24, 9, 7), (8, 5, 50)])
def test_otm_deepfm_matches_oracle(orc, E, leaf_level, beam):
    """otm/.../model/DeepFM.scala:12-48 behind dmg_otm_*: batchBeamSearch dump, per-level dump and recommend, fp64 bits."""
    T, topk, B = 10, 10, 19
    rows = (1 << (leaf_level + 1)) - 1
    n_leaf = 1 << leaf_level
    rng = np.random.default_rng(4)
    n_items = int(n_leaf * 0.8)
    leaf_ids = np.sort(rng.choice(n_leaf, n_items, replace=False)).astype(np.int32) + (n_leaf - 1)
    items = np.arange(1, n_items + 1, dtype=np.int32)
    leaf_item = np.full(n_leaf, -1, np.int32)
    leaf_item[leaf_ids - (n_leaf - 1)] = items
    params = deepfm_params64(rows, E, T, seed=5)
    seqs = leaf_ids[rng.integers(0, n_items, (B, T))].astype(np.int32)
    seqs[rng.random((B, T)) < 0.3] = -1
    seqs[0] = -1                                              # a user with an empty history
    e = new_engine()
    e.load_tree_complete(leaf_level, items, leaf_ids)
    e.load_deepfm_weights(params, rows, E, T)
    model = orc.OtmModel(params, rows, E, T, deepfm=True)
    ids, sc, cnt = e.otm_beam_search(seqs, beam)
    gi, gs, gc = e.otm_retrieve(seqs, beam, topk)
    for u in range(B):
        oi, os_ = model.beam_search(seqs[u], leaf_level, beam)
        assert cnt[u] == len(oi) and (ids[u, :cnt[u]] == oi).all() and (sc[u, :cnt[u]].view(np.uint64) == os_.view(np.uint64)).all()
        assert (ids[u, cnt[u]:] == -1).all()
        ri, rs, _ = model.recommend(seqs[u], leaf_level, topk, beam, leaf_item)
        assert gc[u] == len(ri) and (gi[u, :gc[u]] == ri).all() and (gs[u, :gc[u]].view(np.uint64) == rs.view(np.uint64)).all()
        assert (gi[u, gc[u]:] == -1).all()
    # per-level candidates (OTM training reads them): the last level equals the dump
    li, ls, lc = e.otm_beam_search_levels(seqs, beam, leaf_level)
    if li.shape[1] > 0:
        assert (li[:, -1, :] == ids).all() and (ls[:, -1, :].view(np.uint64) == sc.view(np.uint64)).all() and (lc[:, -1] == cnt).all()
    # the host mirror of OTM.scala reaches the same scorer: OTM(modelName = "DeepFM"), useMask = false (OTM.scala:33)
    from dismember_b200.otm import OTM
    m = OTM(engine=e, model_name="DeepFM").set_mapping(items, leaf_ids).set_parameters(params, E, T)
    assert m.use_mask is False and m.leaf_level == leaf_level
    hist = [int(items[np.searchsorted(leaf_ids, x)]) if x >= 0 else 0 for x in seqs[3]]      # item ids (0 is not an item: padding)
    rec = m.recommend(hist, topk, beam)
    assert [i for i, _ in rec] == gi[3, :gc[3]].tolist()
    e.close()


def test_otm_deepfm_score_pairs_is_model_forward(orc):
    rows, E, T = 1023, 32, 10
    params = deepfm_params64(rows, E, T, seed=9)
    rng = np.random.default_rng(1)
    n = 1000
    node = rng.integers(-1, rows, n).astype(np.int32)
    seq = rng.integers(-1, rows, (n, T)).astype(np.int32)
    e = new_engine()
    e.load_tree_complete(9, np.arange(1, 513, dtype=np.int32), np.arange(511, 1023, dtype=np.int32))
    e.load_deepfm_weights(params, rows, E, T)
    got = e.score_pairs(node, seq)
    want = orc.OtmModel(params, rows, E, T, deepfm=True).forward(node, seq)
    assert got.dtype == np.float64 and (got.view(np.uint64) == want.view(np.uint64)).all()
    from dismember_b200._capi import DmgError, DmgIndexError
    with pytest.raises(DmgIndexError):
        e.score_pairs(np.array([rows], np.int32), seq[:1])
    with pytest.raises(DmgIndexError):
        e.otm_retrieve(np.full((1, T), rows, np.int32), 20, 10)
    with pytest.raises(DmgError):                             # the TDM/JTM scorer is Module[Float]
        e.tdm_retrieve(np.zeros((1, T), np.int32), 20, 10)
    e.close()


def test_deepfm_training_matches_oracle(orc):
    """DeepFM in the training loop (tdm/.../model/DeepFM.scala:11-44 behind LocalOptimizer.trainBatch): gradients of the compact vector
    and three Adam steps against the oracle (atomics reorder the sums: 1e-5 relative, as for the DIN step)."""
    from dismember_b200 import synth
    tf = synth.tdm_tree(2000, seed=3)
    rows, E, T = (1 << (tf.max_level + 1)) - 1, 16, 10
    params = deepfm_params(rows, E, T, seed=8)
    rng = np.random.default_rng(9)
    n = 500
    node = rng.integers(0, rows, n).astype(np.int32)
    seq = rng.integers(0, rows, (n, T)).astype(np.int32)
    seq[rng.random((n, T)) < 0.3] = -1
    seq[0] = -1
    labels = (rng.random(n) < 0.3).astype(np.float32)
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_deepfm_weights(params, rows, E, T)
    g, loss = e.din_gradients(node, seq, None, labels)
    og, oloss = orc.deepfm_gradients(params, rows, E, T, node, seq, labels)
    assert abs(loss - oloss) <= 1e-5 * max(1.0, abs(oloss))
    assert np.abs(g - og).max() <= 2e-5 * np.abs(og).max()
    w = params.copy()
    s, r = np.zeros_like(w), np.zeros_like(w)
    for t in (1, 2, 3):
        og, _ = orc.deepfm_gradients(w, rows, E, T, node, seq, labels)
        orc.adam_step(w, og, s, r, 1e-2, t)
        e.train_step(node, seq, None, labels, 1e-2, t)
    got = e.download_din_weights()
    assert np.abs(got - w).max() <= 2e-4 and np.abs(got - params).max() > 1e-3
    # the trained scorer keeps serving: model.forward on the updated weights == the oracle on the same weights
    want = orc.TdmModel(got, rows, E, T, deepfm=True).forward(node[:64], seq[:64])
    assert (e.score_pairs(node[:64], seq[:64]).view(np.uint32) == want.view(np.uint32)).all()
    e.close()


def test_deepfm_fast_path_runs_and_hands_ties_back(orc):
    """The certified fast path (csrc/beam_wave_dfm.cuh) is what serves E = 64: rows_fast > 0, ids and logits == the strict arithmetic and
    the oracle; an all-zero model (every score ties exactly at every cut) is handed back to the strict level-synchronous path and
    still matches; the eval variant with consumed items and widened beams goes through the same path."""
    from dismember_b200 import synth
    E, T, n_items, beam, topk, B = 64, 10, 20000, 200, 10, 96
    tf = synth.tdm_tree(n_items, seed=5)
    rows = (1 << (tf.max_level + 1)) - 1
    params = deepfm_params(rows, E, T, seed=11)
    seqs = synth.queries(B, T, n_items, seed=12)
    seqs[0] = 0
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_deepfm_weights(params, rows, E, T)
    e.fast_stats()                                            # resets the counters
    fi, fl, fc = e.tdm_retrieve(seqs, beam, topk)
    st = e.fast_stats()
    assert st["rows_fast"] > B * 1000 and st["max_err_over_bound"] < 1.0
    e.set_arithmetic("strict")
    si, sl, sc = e.tdm_retrieve(seqs, beam, topk)
    assert (fc == sc).all() and (fi == si).all() and (fl.view(np.uint32) == sl.view(np.uint32)).all()
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T, deepfm=True)
    oi, ol, oc = model.retrieve_batch(tree, seqs[:32], beam, topk, n_threads=os.cpu_count() or 1)
    assert (fc[:32] == oc).all() and (fi[:32] == oi).all() and (fl[:32].view(np.uint32) == ol.view(np.uint32)).all()
    # consumed items + widened beams
    rng = np.random.default_rng(3)
    cons = [rng.choice(tf.leaf_ids, int(k), replace=False).tolist() for k in rng.choice([0, 5, 420, 480], B)]
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum([len(c) for c in cons])
    flat = np.array([x for c in cons for x in c], np.int32)
    s2 = e.tdm_retrieve(seqs, beam, topk, consumed_off=off, consumed=flat, widen_beam=True)
    e.set_arithmetic("fast")
    f2 = e.tdm_retrieve(seqs, beam, topk, consumed_off=off, consumed=flat, widen_beam=True)
    assert (f2[2] == s2[2]).all() and (f2[0] == s2[0]).all() and (f2[1].view(np.uint32) == s2[1].view(np.uint32)).all()
    # one handle per host thread over one copy of the tables (dmg_clone), two batches in flight
    import threading
    twin = e.clone()
    got = [None, None]

    def work(k, eng, sl):
        got[k] = eng.tdm_retrieve(seqs[sl], beam, topk)
    th = [threading.Thread(target=work, args=(0, e, slice(0, B // 2))), threading.Thread(target=work, args=(1, twin, slice(B // 2, B)))]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    for k, sl in enumerate((slice(0, B // 2), slice(B // 2, B))):
        assert (got[k][0] == fi[sl]).all() and (got[k][1].view(np.uint32) == fl[sl].view(np.uint32)).all() and (got[k][2] == fc[sl]).all()
    assert twin.fast_stats()["rows_fast"] > 0
    twin.close()
    # all-zero model: exact ties everywhere
    zero = np.zeros_like(params)
    e.load_deepfm_weights(zero, rows, E, T)
    fz = e.tdm_retrieve(seqs[:24], beam, topk)
    twin = e.clone()                                          # the clone's own fall-back to the strict level-synchronous path
    tz = twin.tdm_retrieve(seqs[:24], beam, topk)
    twin.close()
    e.set_arithmetic("strict")
    sz = e.tdm_retrieve(seqs[:24], beam, topk)
    assert (fz[2] == sz[2]).all() and (fz[0] == sz[0]).all() and (fz[1].view(np.uint32) == sz[1].view(np.uint32)).all()
    assert (tz[2] == sz[2]).all() and (tz[0] == sz[0]).all() and (tz[1].view(np.uint32) == sz[1].view(np.uint32)).all()
    e.close()


@pytest.mark.parametrize("E", [16, 32])
def test_deepfm_narrow_models_take_the_fast_path_on_a_padded_copy(orc, E):
    """DeepFM with embed_size 16 / 32: the certified fast path runs on a zero-padded E = 64 copy (features [x | 0], W1 with zero
    columns): the FM sums and the Linear chains only gain exact zeros, so ids and logits are the narrow model's bits."""
    from dismember_b200 import synth
    T, n_items, beam, topk, B = 10, 8000, 60, 10, 48
    tf = synth.tdm_tree(n_items, seed=4)
    rows = (1 << (tf.max_level + 1)) - 1
    params = deepfm_params(rows, E, T, seed=14)
    seqs = synth.queries(B, T, n_items, seed=15)
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_deepfm_weights(params, rows, E, T)
    e.fast_stats()
    fi, fl, fc = e.tdm_retrieve(seqs, beam, topk)
    assert e.fast_stats()["rows_fast"] > B * 200
    tree = orc.Tree.from_treefile(tf)
    oi, ol, oc = orc.TdmModel(params, rows, E, T, deepfm=True).retrieve_batch(tree, seqs, beam, topk, n_threads=8)
    assert (fc == oc).all() and (fi == oi).all() and (fl.view(np.uint32) == ol.view(np.uint32)).all()
    assert (e.download_din_weights() == params).all()
    e.close()
