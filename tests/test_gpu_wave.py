"""The level-synchronous tensor-core path (csrc/beam_wave.cuh) piece by piece: the tile scorer alone against
model.forward (oracle) within the certified bound, with both row-gather engines (TMA tile::gather4, cp.async), then
whole searches against the oracle bit for bit."""
import os

import numpy as np
import pytest

from dismember_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32)


def _setup(engine, n_items, E=64, seed=17, structured=True):
    tf = synth.tdm_tree(n_items, seed=seed)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=23, structured=structured)
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    engine.load_din_weights(params, rows, E, 10)
    return tf, rows, params


@pytest.mark.parametrize("gather", ["tma"])
def test_wave_scorer_within_bound_of_model_forward(engine, orc, gather):
    """Fast scores of every candidate of a level = model.forward (Recommender.scala:93-94) within eps; the
    candidates of the first scored level are all children of the start level in code order."""
    os.environ["DMG_WAVE_GATHER"] = gather
    try:
        n_items, beam = 20000, 200
        tf, rows, params = _setup(engine, n_items)
        tree = orc.Tree.from_treefile(tf)
        model = orc.TdmModel(params, rows, 64, 10)
        seqs = synth.queries(48, 10, n_items, seed=29)
        seqs[0] = 0
        engine.set_arithmetic("fast")
        try:
            worst = 0.0
            for level in (8, 9, tf.max_level):
                codes, scores, counts, eps = engine.wave_probe(seqs, beam, level)
                assert (counts[eps >= 0] > 0).all()
                for u in range(len(seqs)):
                    if eps[u] < 0:
                        continue
                    n = counts[u]
                    node = codes[u, :n]
                    lv = np.floor(np.log2(node.astype(np.int64) + 1)).astype(int)
                    assert (lv == level).all()
                    hc, hm = tree.id_to_code(seqs[u])
                    seq = np.tile(hc, (n, 1))
                    mask = (np.arange(n)[:, None] * 10 + np.flatnonzero(hm)[None, :]).ravel().astype(np.int32)
                    want = model.forward(node, seq, mask)
                    err = np.abs(scores[u, :n].astype(np.float64) - want.astype(np.float64))
                    assert np.isfinite(scores[u, :n]).all()
                    assert (err <= eps[u]).all(), (gather, level, u, err.max(), eps[u], scores[u, :4], want[:4])
                    worst = max(worst, float(err.max() / eps[u]))
            print("wave scorer", gather, "max |fast-strict|/eps =", worst)
            assert worst < 0.05
        finally:
            engine.set_arithmetic("strict")
    finally:
        os.environ.pop("DMG_WAVE_GATHER", None)


@pytest.mark.parametrize("gather", ["tma"])
@pytest.mark.parametrize("n_items,beam,B", [(20000, 200, 160), (300, 7, 33), (70000, 256, 97), (5000, 64, 1)])
def test_wave_search_matches_oracle(engine, orc, gather, n_items, beam, B):
    os.environ["DMG_WAVE_GATHER"] = gather
    try:
        tf, rows, params = _setup(engine, n_items)
        seqs = synth.queries(B, 10, n_items, seed=31)
        engine.set_arithmetic("fast")
        try:
            items, logits, counts = engine.tdm_retrieve(seqs, beam, 10)
            stats = engine.fast_stats()
        finally:
            engine.set_arithmetic("strict")
        tree = orc.Tree.from_treefile(tf)
        model = orc.TdmModel(params, rows, 64, 10)
        oi, ol, oc = model.retrieve_batch(tree, seqs, beam, 10, n_threads=8)
        assert (counts == oc).all()
        assert (items == oi).all(), f"ids differ, stats={stats}"
        assert (bits(logits) == bits(ol)).all()
        assert stats["rows_fast"] > 0
        print("wave stats", gather, n_items, beam, stats)
    finally:
        os.environ.pop("DMG_WAVE_GATHER", None)


def test_repeated_steps_replay_a_captured_graph(orc):
    """A serving loop repeats one batch shape on one handle: from the third call on the step is ONE cudaGraphLaunch of the captured
    level-synchronous chain (capi.cu: tdm_enqueue).  Every call -- plain launches, the capture, the replays -- returns the oracle's
    ids and logit bits for ITS queries (host-buffer and device-buffer entry points, with a per-thread clone, with consumed items)."""
    import torch
    from conftest import new_engine
    from dismember_b200 import synth
    n_items, E, T, beam, topk, B = 30000, 64, 10, 200, 10, 64
    tf = synth.tdm_tree(n_items, seed=2)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=3, structured=True)
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_din_weights(params, rows, E, T)
    twin = e.clone()
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T)
    dev = torch.device("cuda", 0)
    d_items = torch.empty((B, topk), dtype=torch.int32, device=dev)
    d_log = torch.empty((B, topk), dtype=torch.float32, device=dev)
    d_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    l0 = e.launch_count
    for it in range(6):
        q = synth.queries(B, T, n_items, seed=50 + it)
        oi, ol, oc = model.retrieve_batch(tree, q, beam, topk, n_threads=8)
        gi, gl, gc = e.tdm_retrieve(q, beam, topk)
        assert (gc == oc).all() and (gi == oi).all() and (gl.view(np.uint32) == ol.view(np.uint32)).all(), it
        dq = torch.from_numpy(q).to(dev)                          # a fresh device buffer every call: the handle stages it
        torch.cuda.synchronize()
        twin.tdm_retrieve_dev_sync(B, dq.data_ptr(), beam, topk, True, d_items.data_ptr(), d_log.data_ptr(), d_cnt.data_ptr())
        assert (d_cnt.cpu().numpy() == oc).all() and (d_items.cpu().numpy() == oi).all()
        assert (d_log.cpu().numpy().view(np.uint32) == ol.view(np.uint32)).all(), it
    assert e.launch_count - l0 >= 6 * 15                           # replays count the kernels inside the graph
    # the eval variant (consumed items, widened beams) is another key of the same cache
    rng = np.random.default_rng(1)
    for it in range(4):
        q = synth.queries(B, T, n_items, seed=70 + it)
        cons = [rng.choice(tf.leaf_ids, int(k), replace=False).tolist() for k in rng.choice([0, 4, 30], B)]
        off = np.zeros(B + 1, np.int64)
        off[1:] = np.cumsum([len(c) for c in cons])
        flat = np.array([x for c in cons for x in c], np.int32)
        ob = model.retrieve_batch(tree, q, beam, topk, cons_off=off, cons=flat, widen_beam=True, n_threads=8)
        gb = e.tdm_retrieve(q, beam, topk, consumed_off=off, consumed=flat, widen_beam=True)
        assert (gb[2] == ob[2]).all() and (gb[0] == ob[0]).all() and (gb[1].view(np.uint32) == ob[1].view(np.uint32)).all(), it
    twin.close()
    e.close()


@pytest.mark.parametrize("E", [16, 32])
def test_narrow_models_run_the_tensor_core_path_on_a_zero_padded_copy(orc, E):
    """embed_size 16 / 32 (every configuration the reference ships uses 16): FAST arithmetic builds a zero-padded E = 64 copy with the
    ORIGINAL attention scale 1 / sqrt(E) and runs the tensor-core path on it.  Exact zeros appended to every sequential-k chain change
    no bit, so ids and logits equal the strict E-wide kernel and the oracle; the fast scorer did score rows; clones and the eval
    variant go through the same copy; score_pairs / download still see the narrow model."""
    from conftest import new_engine
    n_items, T, beam, topk, B = 20000, 10, 200, 10, 80
    tf = synth.tdm_tree(n_items, seed=6)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=7, structured=True)
    seqs = synth.queries(B, T, n_items, seed=8)
    seqs[0] = 0
    e = new_engine()
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.load_din_weights(params, rows, E, T)
    e.fast_stats()
    fi, fl, fc = e.tdm_retrieve(seqs, beam, topk)
    st = e.fast_stats()
    assert st["rows_fast"] > B * 1000 and st["max_err_over_bound"] < 1.0
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, topk, n_threads=8)
    assert (fc == oc).all() and (fi == oi).all() and (fl.view(np.uint32) == ol.view(np.uint32)).all()
    twin = e.clone()
    ti, tl, tc = twin.tdm_retrieve(seqs, beam, topk)
    assert (tc == oc).all() and (ti == oi).all() and (tl.view(np.uint32) == ol.view(np.uint32)).all()
    assert twin.fast_stats()["rows_fast"] > 0
    twin.close()
    e.set_arithmetic("strict")
    si, sl, sc = e.tdm_retrieve(seqs, beam, topk)
    assert (sc == oc).all() and (si == oi).all() and (sl.view(np.uint32) == ol.view(np.uint32)).all()
    e.set_arithmetic("fast")
    assert (e.download_din_weights() == params).all()
    node = np.arange(100, 164, dtype=np.int32)
    assert (e.score_pairs(node, seqs[:64] * 0 + node[:, None]).view(np.uint32) == model.forward(node, seqs[:64] * 0 + node[:, None]).view(np.uint32)).all()
    # new weights: the copy follows
    params2 = synth.din_params(rows, E, seed=9, structured=True)
    e.load_din_weights(params2, rows, E, T)
    o2 = orc.TdmModel(params2, rows, E, T).retrieve_batch(tree, seqs[:32], beam, topk, n_threads=8)
    f2 = e.tdm_retrieve(seqs[:32], beam, topk)
    assert (f2[2] == o2[2]).all() and (f2[0] == o2[0]).all() and (f2[1].view(np.uint32) == o2[1].view(np.uint32)).all()
    e.close()
