"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs.  Bar: item ids identical (ties broken identically), scores BIT-identical -- the
engine's strict arithmetic spec equals the oracle's (tolerance stated where it is not zero)."""
import numpy as np
import pytest

from dismember_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def load_jtm(engine, f):
    engine.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    engine.load_din_weights(f["params"], 8191, int(f["E"]), int(f["T"]))


def load_otm(engine, f):
    n = len(f["items"])
    leaf_level = int(np.ceil(np.log(n) / np.log(2)))
    engine.load_tree_complete(leaf_level, f["items"], f["leaf_ids"])
    engine.load_din_weights(f["params"], 8191, int(f["E"]), int(f["T"]))
    return leaf_level


# ------------------------------------------------------------------ fixtures of the reference
@pytest.mark.parametrize("beam", [20, 200])
def test_tdm_fixture_matches_golden_and_oracle(engine, jtm_fix, jtm_oracle, queries, golden_out, beam):
    load_jtm(engine, jtm_fix)
    items, logits, counts = engine.tdm_retrieve(queries["seqs"], beam, 10)
    assert (items == golden_out[f"tdm_items_b{beam}"]).all()
    assert (bits(logits) == bits(golden_out[f"tdm_logits_b{beam}"])).all()
    assert (counts == golden_out[f"tdm_counts_b{beam}"]).all()
    tree, model = jtm_oracle
    oi, ol, oc = model.retrieve_batch(tree, queries["seqs"], beam, 10, n_threads=4)
    assert (items == oi).all() and (bits(logits) == bits(ol)).all() and (counts == oc).all()


def test_tdm_eval_variant_with_consumed(engine, jtm_fix, queries, golden_out):
    load_jtm(engine, jtm_fix)
    items, logits, counts = engine.tdm_retrieve(queries["seqs"], 20, 10, True, queries["cons_off"], queries["cons"], True)
    assert (items == golden_out["tdm_eval_items"]).all()
    assert (bits(logits) == bits(golden_out["tdm_eval_logits"])).all()
    assert (counts == golden_out["tdm_eval_counts"]).all()


@pytest.mark.parametrize("beam", [20, 200])
def test_otm_fixture(engine, otm_fix, otm_oracle, queries, golden_out, beam):
    load_otm(engine, otm_fix)
    model, leaf_level, leaf_item, item_leaf = otm_oracle
    seqs = np.array([[item_leaf.get(int(x), -1) for x in s] for s in queries["seqs"]], np.int32)
    items, scores, counts = engine.otm_retrieve(seqs, beam, 10)
    assert (items == golden_out[f"otm_items_b{beam}"]).all()
    assert (bits(scores) == bits(golden_out[f"otm_scores_b{beam}"])).all()
    ids, sc, cnt = engine.otm_beam_search(seqs[:32], beam)
    for u in range(32):
        oi, os_ = model.beam_search(seqs[u], leaf_level, beam)
        assert cnt[u] == len(oi) and (ids[u, :cnt[u]] == oi).all() and (bits(sc[u, :cnt[u]]) == bits(os_)).all()


@pytest.mark.parametrize("which", ["f32", "f64"])
def test_score_pairs_is_model_forward(engine, orc, jtm_fix, otm_fix, which):
    fix = jtm_fix if which == "f32" else otm_fix
    engine.load_din_weights(fix["params"], 8191, 16, 10)
    model = (orc.TdmModel if which == "f32" else orc.OtmModel)(fix["params"], 8191, 16, 10)
    rng = np.random.default_rng(1)
    n = 1000
    node = rng.integers(0, 8191, n).astype(np.int32)
    seq = rng.integers(0, 8191, (n, 10)).astype(np.int32)
    seq[rng.random((n, 10)) < 0.3] = -1
    seq[:7] = -1
    mask = np.flatnonzero((seq == -1).ravel()).astype(np.int32)
    got = engine.score_pairs(node, seq, mask)
    want = model.forward(node, seq, mask)
    assert (bits(got) == bits(want)).all()
    # unmasked padding (useMask = false path): zero rows still enter the softmax
    got = engine.score_pairs(node[:64], seq[:64], None)
    assert (bits(got) == bits(model.forward(node[:64], seq[:64], None))).all()


# ------------------------------------------------------------------ synthetic catalogues
@pytest.mark.parametrize("E,n_items,beam", [(64, 20000, 200), (32, 5000, 50), (16, 1000, 200), (64, 300, 7)])
def test_tdm_synthetic(engine, orc, E, n_items, beam):
    tf = synth.tdm_tree(n_items, seed=E)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=7)
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    engine.load_din_weights(params, rows, E, 10)
    seqs = synth.queries(96, 10, n_items, seed=11)
    seqs[0] = 0                                               # all padding
    items, logits, counts = engine.tdm_retrieve(seqs, beam, 10)
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, 10)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, 10, n_threads=8)
    assert (counts == oc).all()
    assert (items == oi).all()
    assert (bits(logits) == bits(ol)).all()


def test_otm_synthetic_e64(engine, orc):
    items, leaves, leaf_level = synth.otm_mapping(3000, seed=5)
    rows = (1 << (leaf_level + 1)) - 1
    params = synth.din_params(rows, 64, seed=9, dtype=np.float64)
    engine.load_tree_complete(leaf_level, items, leaves)
    engine.load_din_weights(params, rows, 64, 10)
    model = orc.OtmModel(params, rows, 64, 10)
    leaf_item = np.full(1 << leaf_level, -1, np.int32)
    leaf_item[leaves - ((1 << leaf_level) - 1)] = items
    rng = np.random.default_rng(3)
    seqs = leaves[rng.integers(0, len(leaves), (24, 10))].astype(np.int32)
    seqs[rng.random(seqs.shape) < 0.2] = -1
    got_i, got_s, got_c = engine.otm_retrieve(seqs, 100, 10)
    oi, os_, oc = model.retrieve_batch(seqs, leaf_level, 100, 10, leaf_item, n_threads=8)
    assert (got_c == oc).all() and (got_i == oi).all() and (bits(got_s) == bits(os_)).all()


# ------------------------------------------------------------------ ties, edges, errors
def test_ties_break_like_stable_sort(engine, orc, jtm_fix, jtm_oracle):
    """A zero model makes every score equal: the result is decided by tie-breaking alone."""
    f = jtm_fix
    engine.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    zero = np.zeros(131857, np.float32)
    engine.load_din_weights(zero, 8191, 16, 10)
    tree, _ = jtm_oracle
    model = orc.TdmModel(zero, 8191, 16, 10)
    seqs = np.zeros((3, 10), np.int32)
    seqs[1, 5:] = [2126, 204, 3257, 3439, 996]
    for beam in (5, 20, 200):
        items, logits, counts = engine.tdm_retrieve(seqs, beam, 10)
        oi, ol, oc = model.retrieve_batch(tree, seqs, beam, 10)
        assert (items == oi).all() and (counts == oc).all() and (logits == 0).all()


def test_beam_edges(engine, orc, jtm_fix, jtm_oracle):
    load_jtm(engine, jtm_fix)
    tree, model = jtm_oracle
    seqs = np.zeros((2, 10), np.int32)
    seqs[1] = [0, 0, 2126, 204, 3257, 3439, 996, 1681, 3438, 1882]
    for beam, topk in [(1, 1), (2, 5), (3, 10), (4096, 10), (5000, 10), (9000, 3), (63, 200)]:
        try:
            items, logits, counts = engine.tdm_retrieve(seqs, beam, topk)
        except Exception as e:                      # very wide beams may exceed shared memory: must be a clean error
            assert "beam too large" in str(e)
            continue
        oi, ol, oc = model.retrieve_batch(tree, seqs, beam, topk)
        assert (counts == oc).all() and (items == oi).all() and (bits(logits) == bits(ol)).all(), (beam, topk)


def test_single_user_and_odd_batches(engine, jtm_fix, golden_out, queries):
    load_jtm(engine, jtm_fix)
    for B in (1, 3, 149, 256):
        items, logits, counts = engine.tdm_retrieve(queries["seqs"][:B], 20, 10)
        assert (items == golden_out["tdm_items_b20"][:B]).all()


def test_invalid_index_raises(engine, jtm_fix):
    from dismember_b200 import DmgIndexError, DmgArgumentError
    load_jtm(engine, jtm_fix)
    seqs = np.zeros((2, 10), np.int32)
    seqs[0, 3] = -7                                  # idToCode yields a negative embedding index
    with pytest.raises(DmgIndexError):
        engine.tdm_retrieve(seqs, 20, 10)
    # engine stays usable afterwards
    seqs[0, 3] = 0
    engine.tdm_retrieve(seqs, 20, 10)
    with pytest.raises(DmgIndexError):
        engine.score_pairs(np.array([9000], np.int32), np.zeros((1, 10), np.int32))
    with pytest.raises(DmgArgumentError):
        engine.tdm_retrieve(seqs, 0, 10)             # require(candidateNum > 0)


def test_unknown_item_quirk(engine, jtm_fix, jtm_oracle):
    """ids above nonLeafOffset address ancestors (id - offset); too-large ones are masked."""
    load_jtm(engine, jtm_fix)
    tree, model = jtm_oracle
    off = int(jtm_fix["leaf_ids"].max()) + 1
    seqs = np.zeros((3, 10), np.int32)
    seqs[0, -3:] = [off + 5, off + 100, off + 8189]          # ancestor pseudo-ids
    seqs[1, -2:] = [off + 8190, off + 100000]                # > maxCode -> masked
    items, logits, counts = engine.tdm_retrieve(seqs, 20, 10)
    oi, ol, oc = model.retrieve_batch(tree, seqs, 20, 10)
    assert (items == oi).all() and (bits(logits) == bits(ol)).all()
    # an id below nonLeafOffset that is not a leaf becomes a negative index: the reference throws
    from dismember_b200 import DmgIndexError
    known = set(jtm_fix["leaf_ids"].tolist())
    missing = next(i for i in range(1, off - 1) if i not in known)
    seqs[2, -1] = missing
    with pytest.raises(DmgIndexError):
        engine.tdm_retrieve(seqs, 20, 10)
    with pytest.raises(IndexError):
        model.retrieve_batch(tree, seqs, 20, 10)


# ------------------------------------------------------------------ host mirrors read like the Scala specs
def test_tdm_recommend_api(jtm_fix, golden_out):
    """TdmModelTrainSpec.scala:85-96: recommend(sequence, topk = 3, candidateNum = 20) has length 3 and is
    identical after the model is loaded again."""
    from dismember_b200.formats.tree_file import TreeFile
    from dismember_b200.tdm import TDM
    f = jtm_fix
    tf = TreeFile(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["prob"], f["leaf_ids"], f["leaf_codes"])
    sequence = [0, 0, 2126, 204, 3257, 3439, 996, 1681, 3438, 1882]
    m1 = TDM().set_tree(tf).set_parameters(f["params"], 16, 10)
    rec1 = m1.recommend(sequence, topk=3, candidate_num=20)
    assert len(rec1) == 3
    assert [r[0] for r in rec1] == golden_out["tdm_items_b20"][0, :3].tolist()
    assert all(0.0 < r[1] < 1.0 for r in rec1)
    params_back = m1.engine.download_din_weights()
    m2 = TDM().set_tree(tf).set_parameters(params_back, 16, 10)
    assert m2.recommend(sequence, topk=3, candidate_num=20) == rec1
    m1.engine.close(); m2.engine.close()


def test_otm_recommend_api(otm_fix, golden_out):
    from dismember_b200.otm import OTM
    f = otm_fix
    m = OTM().set_mapping(f["items"], f["leaf_ids"]).set_parameters(f["params"], 16, 10)
    rec = m.recommend([0, 0, 2126, 204, 3257, 3439, 996, 1681, 3438, 1882], topk=3, beam_size=20)
    assert len(rec) == 3 and [r[0] for r in rec] == golden_out["otm_items_b20"][0, :3].tolist()
    m.engine.close()


# ------------------------------------------------------------------ BASELINE-size properties
def test_full_size_properties(engine):
    """1M-item catalogue (BASELINE configs[1]): size-independent properties instead of an oracle run:
    determinism, sortedness, leaf range, beam monotonicity of the best score."""
    n_items = 1_000_000
    tf = synth.tdm_tree(n_items, seed=1)
    rows = (1 << (tf.max_level + 1)) - 1
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    engine.init_din_weights(np.float32, rows, 64, 10, seed=2)
    seqs = synth.queries(1024, 10, n_items, seed=4)
    a = engine.tdm_retrieve(seqs, 200, 10)
    b = engine.tdm_retrieve(seqs, 200, 10)
    assert all((x == y).all() for x, y in zip(a, b))                       # idempotent / deterministic
    items, logits, counts = a
    assert (counts == 10).all() and items.min() >= 1 and items.max() <= n_items
    assert (np.diff(logits, axis=1) <= 0).all()                            # sorted descending
    assert all(len(set(r.tolist())) == 10 for r in items[:64])             # no duplicates
    # a user's result does not depend on who else is in the batch
    c = engine.tdm_retrieve(seqs[100:164], 200, 10)
    assert (c[0] == items[100:164]).all() and (bits(c[1]) == bits(logits[100:164])).all()
    # the retrieved logits equal model.forward on (leaf code, history)
    item_code = np.zeros(n_items + 1, np.int64)
    item_code[tf.leaf_ids] = tf.leaf_codes
    u = 5
    hist = np.where(seqs[u] > 0, item_code[np.maximum(seqs[u], 0)], -1).astype(np.int32)
    node = item_code[items[u]].astype(np.int32)
    fw = engine.score_pairs(node, np.tile(hist, (10, 1)), np.flatnonzero(np.tile(hist == -1, 10)).astype(np.int32))
    assert (bits(fw) == bits(logits[u])).all()


# ------------------------------------------------------------------ Deep Retrieval
def _dr_models(orc, f):
    D = int(f["D"])
    args = (int(f["num_item"]), int(f["K"]), D, int(f["T"]), int(f["E"]), f["layer_emb"],
            [f[f"layer_w{d}"] for d in range(D)], [f[f"layer_b{d}"] for d in range(D)],
            f["rr_emb"], f["rr_w"], f["rr_b"], f["sm_w"], f["sm_b"])
    return args, orc.DrModel(*args)


@pytest.mark.parametrize("beam", [20, 50, 1, 300])
def test_dr_beam_search_fixture(engine, orc, dr_fix, queries, golden_out, beam):
    args, model = _dr_models(orc, dr_fix)
    engine.dr_load(*args)
    item_id = {int(a): int(b) for a, b in zip(dr_fix["map_items"], dr_fix["map_ids"])}
    seqs = np.array([[item_id.get(int(x), -1) for x in s] for s in queries["seqs"][:48]], np.int32)
    seqs[1] = -1
    paths, probs, counts = engine.dr_beam_search(seqs, beam)
    for u in range(len(seqs)):
        op, opr = model.beam_search(seqs[u], beam)
        assert counts[u] == len(op)
        assert (paths[u, :counts[u]] == op).all(), (u, beam)
        assert (bits(probs[u, :counts[u]]) == bits(opr)).all()
    if beam in (20, 50):
        assert (paths[0] == golden_out[f"dr_paths_b{beam}"][0]).all()


def test_dr_retrieve_fixture(engine, orc, dr_fix, queries):
    from dismember_b200.dr import build_path_csr
    args, model = _dr_models(orc, dr_fix)
    engine.dr_load(*args)
    off, flat = build_path_csr(dr_fix["map_ids"], dr_fix["map_paths"], int(dr_fix["K"]))
    engine.dr_load_paths(off, flat)
    item_id = {int(a): int(b) for a, b in zip(dr_fix["map_items"], dr_fix["map_ids"])}
    seqs = np.array([[item_id.get(int(x), -1) for x in s] for s in queries["seqs"][:32]], np.int32)
    for beam, topk in [(50, 10), (300, 20)]:
        items, scores, counts = engine.dr_retrieve(seqs, beam, topk)
        for u in range(len(seqs)):
            oi, os_, _ = model.recommend(seqs[u], topk, beam, off, flat)
            assert counts[u] == len(oi) and (items[u, :counts[u]] == oi).all()
            assert (bits(scores[u, :counts[u]]) == bits(os_)).all()


def test_dr_synthetic_wide(engine, orc):
    """K=64, D=3, E=32 with many ties-free random weights and a dense path->items map."""
    rng = np.random.default_rng(12)
    num_item, K, D, T, E = 500, 64, 3, 10, 32
    layer_emb = rng.normal(0, 0.3, (num_item + K * (D - 1), E))
    layer_w = [rng.normal(0, 0.3, (K, (T + d) * E)) for d in range(D)]
    layer_b = [rng.normal(0, 0.1, K) for d in range(D)]
    rr_emb = rng.normal(0, 0.3, (num_item, E)); rr_w = rng.normal(0, 0.2, (E, T * E)); rr_b = rng.normal(0, 0.1, E)
    sm_w = rng.normal(0, 0.3, (num_item, E)); sm_b = rng.normal(0, 0.1, num_item)
    args = (num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
    model = orc.DrModel(*args)
    engine.dr_load(*args)
    from dismember_b200.dr import build_path_csr
    paths = rng.integers(0, 4, (num_item, 3, D))            # few distinct paths -> long item lists
    off, flat = build_path_csr(np.arange(num_item), paths, K, keep_all_items=True)     # many items per path: long rerank lists
    engine.dr_load_paths(off, flat)
    seqs = rng.integers(-1, num_item, (20, T)).astype(np.int32)
    p, pr, c = engine.dr_beam_search(seqs, 100)
    items, scores, counts = engine.dr_retrieve(seqs, 100, 25)
    for u in range(len(seqs)):
        op, opr = model.beam_search(seqs[u], 100)
        assert (p[u, :c[u]] == op).all() and (bits(pr[u, :c[u]]) == bits(opr)).all()
        oi, os_, _ = model.recommend(seqs[u], 25, 100, off, flat)
        assert counts[u] == len(oi) and (items[u, :counts[u]] == oi).all() and (bits(scores[u, :counts[u]]) == bits(os_)).all()


# ------------------------------------------------------------------ tensor-core scorer with certified cuts
@pytest.mark.parametrize("n_items,beam,structured", [(20000, 200, True), (20000, 200, False), (300, 7, True), (5000, 64, True), (70000, 256, True), (70000, 333, True)])
def test_fast_mode_matches_oracle(engine, orc, n_items, beam, structured):
    """DMG_ARITH_FAST (tcgen05 bf16x3 + certified cuts) must return the oracle's ids AND logit bits."""
    E = 64
    tf = synth.tdm_tree(n_items, seed=17)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=23, structured=structured)
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    engine.load_din_weights(params, rows, E, 10)
    seqs = synth.queries(160, 10, n_items, seed=29)
    seqs[0] = 0
    engine.set_arithmetic("fast")
    try:
        items, logits, counts = engine.tdm_retrieve(seqs, beam, 10)
        stats = engine.fast_stats()
    finally:
        engine.set_arithmetic("strict")
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, 10)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, 10, n_threads=8)
    assert (counts == oc).all()
    assert (items == oi).all(), f"ids differ, stats={stats}"
    assert (bits(logits) == bits(ol)).all()
    if beam <= 256:                                       # wider beams do not fit the fast kernel's shared memory
        assert stats["rows_fast"] > 0                     # and fall back to the strict kernel (same results)
        assert stats["max_err_over_bound"] < 0.05         # observed |fast - strict| stays far inside the bound
    print("fast stats", n_items, beam, stats)


def test_fast_mode_ties_and_consumed(engine, orc):
    """all-zero model (every score ties exactly) and the eval variant with consumed items, fast mode."""
    E, n_items = 64, 3000
    tf = synth.tdm_tree(n_items, seed=3)
    rows = (1 << (tf.max_level + 1)) - 1
    tree = orc.Tree.from_treefile(tf)
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    seqs = synth.queries(20, 10, n_items, seed=31)
    rng = np.random.default_rng(2)
    cons_off = np.zeros(21, np.int64)
    cons = []
    for u in range(20):
        c = sorted(set(rng.choice(tf.leaf_ids, int(rng.integers(0, 500)), replace=False).tolist()))
        cons.extend(c)
        cons_off[u + 1] = len(cons)
    cons = np.array(cons, np.int32)
    for params in (np.zeros(rows * E + 3 * E * E + 2 * E + 1, np.float32), synth.din_params(rows, E, seed=5)):
        engine.load_din_weights(params, rows, E, 10)
        model = orc.TdmModel(params, rows, E, 10)
        engine.set_arithmetic("fast")
        try:
            a = engine.tdm_retrieve(seqs, 50, 10)
            b = engine.tdm_retrieve(seqs, 40, 10, True, cons_off, cons, True)
        finally:
            engine.set_arithmetic("strict")
        oa = model.retrieve_batch(tree, seqs, 50, 10)
        ob = model.retrieve_batch(tree, seqs, 40, 10, cons_off=cons_off, cons=cons, widen_beam=True)
        for got, want in ((a, oa), (b, ob)):
            assert (got[2] == want[2]).all() and (got[0] == want[0]).all() and (bits(got[1]) == bits(want[1])).all()


def test_fast_equals_strict_at_full_size(engine):
    """BASELINE configs[1] (1M items, beam 200, 1024 users): the two arithmetic modes agree bit for bit."""
    n_items = 1_000_000
    tf = synth.tdm_tree(n_items, seed=1)
    rows = (1 << (tf.max_level + 1)) - 1
    engine.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    engine.init_din_weights(np.float32, rows, 64, 10, seed=2)
    seqs = synth.queries(1024, 10, n_items, seed=4)
    strict = engine.tdm_retrieve(seqs, 200, 10)
    engine.set_arithmetic("fast")
    try:
        fast = engine.tdm_retrieve(seqs, 200, 10)
        stats = engine.fast_stats()
    finally:
        engine.set_arithmetic("strict")
    assert (fast[2] == strict[2]).all() and (fast[0] == strict[0]).all() and (bits(fast[1]) == bits(strict[1])).all()
    print("fast stats at 1M:", stats)


# ------------------------------------------------------------------ clones: one handle per host thread over one model
def test_clones_share_the_model_across_threads(orc):
    """dmg_clone: per-thread handles over the parent's tables (LocalOptimizer.scala:35-40, Evaluator.scala:29-37; the
    reference pins the sharing in otm/src/test/scala/CloneModelSpec.scala: "cloned models share same weights storage").
    Three host threads retrieve their slices of the users at the same time; every slice equals the oracle bit for bit."""
    import threading
    from dismember_b200 import DmgError, Engine
    E, n_items, beam = 64, 20000, 200
    tf = synth.tdm_tree(n_items, seed=17)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=23)
    parent = Engine(0)
    parent.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    parent.load_din_weights(params, rows, E, 10)
    parent.set_arithmetic("fast")
    clones = [parent.clone(), parent.clone()]
    engines = [parent] + clones
    seqs = synth.queries(900, 10, n_items, seed=41)
    slices = [slice(0, 300), slice(300, 600), slice(600, 900)]
    out, errs = [None] * 3, []

    def work(k):
        try:
            for _ in range(3):                                   # several batches per thread, all in flight together
                out[k] = engines[k].tdm_retrieve(seqs[slices[k]], beam, 10)
        except Exception as ex:                                  # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, 10)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, 10, n_threads=8)
    for k in range(3):
        items, logits, counts = out[k]
        assert (counts == oc[slices[k]]).all() and (items == oi[slices[k]]).all() and (bits(logits) == bits(ol[slices[k]])).all()
    assert clones[0].fast_stats()["rows_fast"] > 0              # the clone ran the tensor-core kernel with its own bound tables
    # device-buffer forms on a clone: asynchronous and synchronous (host-gated strict redo) agree with the oracle too
    import torch
    dq = torch.from_numpy(seqs[:300]).cuda()
    for fn in (clones[1].tdm_retrieve_dev, clones[1].tdm_retrieve_dev_sync):
        di = torch.full((300, 10), -7, dtype=torch.int32, device="cuda")
        dl = torch.zeros((300, 10), dtype=torch.float32, device="cuda")
        dc = torch.zeros((300,), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        fn(300, dq.data_ptr(), beam, 10, True, di.data_ptr(), dl.data_ptr(), dc.data_ptr())
        clones[1].synchronize()
        assert (di.cpu().numpy() == oi[:300]).all() and (bits(dl.cpu().numpy()) == bits(ol[:300])).all() and (dc.cpu().numpy() == oc[:300]).all()
    # the shared model is frozen while clones live
    with pytest.raises(DmgError):
        clones[0].load_din_weights(params, rows, E, 10)
    with pytest.raises(DmgError):
        parent.load_din_weights(params, rows, E, 10)
    with pytest.raises(DmgError):
        parent.close()
    for c in clones:
        c.close()
    parent.load_din_weights(params, rows, E, 10)                 # free again
    parent.close()


def test_clones_run_otm_and_deep_retrieval(orc, otm_fix, otm_oracle, dr_fix, queries, golden_out):
    """A clone shares the OTM (fp64) tables and the Deep Retrieval tables of its parent: same results as the parent's
    golden outputs, from two threads at once."""
    import threading
    from dismember_b200 import Engine
    from dismember_b200.dr import build_path_csr
    parent = Engine(0)
    load_otm(parent, otm_fix)
    args, dr_model = _dr_models(orc, dr_fix)
    parent.dr_load(*args)
    off, flat = build_path_csr(dr_fix["map_ids"], dr_fix["map_paths"], int(dr_fix["K"]))
    parent.dr_load_paths(off, flat)
    twin = parent.clone()
    _, _, _, item_leaf = otm_oracle
    oseqs = np.array([[item_leaf.get(int(x), -1) for x in s] for s in queries["seqs"]], np.int32)
    item_id = {int(a): int(b) for a, b in zip(dr_fix["map_items"], dr_fix["map_ids"])}
    dseqs = np.array([[item_id.get(int(x), -1) for x in s] for s in queries["seqs"][:32]], np.int32)
    out, errs = {}, []

    def work(name, e):
        try:
            out[name] = (e.otm_retrieve(oseqs, 20, 10), e.dr_retrieve(dseqs, 50, 10))
        except Exception as ex:                                  # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=work, args=(n, e)) for n, e in (("parent", parent), ("twin", twin))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for name in ("parent", "twin"):
        (oi, osc, _), (di, dsc, dc) = out[name]
        assert (oi == golden_out["otm_items_b20"]).all() and (bits(osc) == bits(golden_out["otm_scores_b20"])).all()
        for u in range(len(dseqs)):
            ri, rs, _ = dr_model.recommend(dseqs[u], 10, 50, off, flat)
            assert dc[u] == len(ri) and (di[u, :dc[u]] == ri).all() and (bits(dsc[u, :dc[u]]) == bits(rs)).all()
    twin.close()
    parent.close()
