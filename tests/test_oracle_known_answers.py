"""Pin the oracle's numerics on everything the reference's own tests pin, plus independent
float64 restatements in numpy.  (scalann/src/test/scala/SoftMaxTest.scala)"""
import numpy as np
import pytest


def test_softmax_known_answer(orc):
    # SoftMaxTest.scala:8-27 (values from PyTorch), tolerance 1e-4 as in the Scala test
    x = np.array([[5, 2, 0.8], [0.3, 0.4, 1.0]], np.float32)
    want = np.array([[0.9392, 0.0468, 0.0141], [0.2428, 0.2683, 0.4889]], np.float32)
    got = orc.softmax_f32(x)
    assert np.abs(got - want).max() < 1e-4
    go = np.array([[0.5, 0.1, 0.6], [0.1, 0.9, 0.6]], np.float32)
    want_g = np.array([[0.0162, -0.0179, 0.0017], [-0.1115, 0.0915, 0.0200]], np.float32)
    assert np.abs(orc.softmax_grad_f32(got, go) - want_g).max() < 1e-4


def test_exp_accuracy(orc):
    L = orc.lib()
    xs = np.concatenate([np.linspace(-104, 88.7, 20001), -np.logspace(-6, 2, 500), [0.0, -0.0, 1.0, -1.0]])
    got = np.array([L.orc_expf_api(float(np.float32(x))) for x in xs], np.float64)
    ref = np.exp(xs.astype(np.float32).astype(np.float64))
    normal = ref > 1e-37
    rel = np.abs(got[normal] - ref[normal]) / ref[normal]
    assert rel.max() < 2.5e-7                       # <= ~2 ulp of fp32
    assert L.orc_expf_api(0.0) == 1.0 and L.orc_expf_api(-3.4e38) == 0.0
    xs = np.concatenate([np.linspace(-745, 709, 20001), [0.0, 1.0, -1.0]])
    got = np.array([L.orc_exp_api(float(x)) for x in xs])
    ref = np.exp(xs)
    normal = ref > 1e-300
    assert (np.abs(got[normal] - ref[normal]) / ref[normal]).max() < 4.5e-16
    assert L.orc_exp_api(0.0) == 1.0


def _din_numpy(params, rows, E, T, node, seq, masked):
    """independent float64 restatement (different code path: dense numpy ops)"""
    p = np.asarray(params, np.float64)
    o = 0
    emb = p[o:o + rows * E].reshape(rows, E); o += rows * E
    watt = p[o:o + E * E].reshape(E, E); o += E * E
    w1 = p[o:o + 2 * E * E].reshape(E, 2 * E); o += 2 * E * E
    b1 = p[o:o + E]; o += E
    w2 = p[o:o + E]; o += E
    b2 = p[o]
    q = np.where(node[:, None] >= 0, emb[np.maximum(node, 0)], 0.0)
    K = np.where(seq[:, :, None] >= 0, emb[np.maximum(seq, 0)], 0.0)
    s = np.einsum("ne,nte->nt", q, K) / np.sqrt(E)
    s = np.where(masked, -3.4028234663852886e38, s)
    s = s - s.max(1, keepdims=True)
    pr = np.exp(s)
    pr /= pr.sum(1, keepdims=True)
    a = np.einsum("nt,nte->ne", pr, K)
    att = a @ watt.T
    h = np.maximum(np.concatenate([q, att], 1) @ w1.T + b1, 0.0)
    return h @ w2 + b2


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_din_forward_matches_numpy(orc, jtm_fix, otm_fix, dtype):
    rng = np.random.default_rng(0)
    fix = jtm_fix if dtype == np.float32 else otm_fix
    model = (orc.TdmModel if dtype == np.float32 else orc.OtmModel)(fix["params"], 8191, 16, 10)
    n = 300
    node = rng.integers(0, 8191, n).astype(np.int32)
    seq = rng.integers(0, 8191, (n, 10)).astype(np.int32)
    seq[rng.random((n, 10)) < 0.3] = -1
    seq[:5] = -1                                    # fully padded histories
    masked = seq == -1
    got = model.forward(node, seq, np.flatnonzero(masked.ravel()).astype(np.int32))
    want = _din_numpy(fix["params"], 8191, 16, 10, node, seq, masked)
    tol = 2e-5 if dtype == np.float32 else 1e-12
    assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max())


def test_din_forward_rejects_bad_index(orc, jtm_fix):
    model = orc.TdmModel(jtm_fix["params"], 8191, 16, 10)
    node = np.array([8191], np.int32)
    seq = np.full((1, 10), -1, np.int32)
    with pytest.raises(IndexError):
        model.forward(node, seq)


def test_sort_is_stable_total_order(orc, jtm_oracle):
    # all-equal scores: a zero model keeps the first `beam` candidates in code order
    tree, _ = jtm_oracle
    zero = np.zeros(131857, np.float32)
    m = orc.TdmModel(zero, 8191, 16, 10)
    items, logits = m.recommend_raw(tree, np.zeros(10, np.int32), 20)
    assert (logits == 0).all() and len(items) > 0


def test_deepfm_oracle_against_independent_float64(orc):
    """DeepFM restatement (tdm/.../model/DeepFM.scala:11-44, nn/FM.scala:14-44) vs a float64 numpy evaluation of the
    same graph written from the layer definitions: FM = (|sum F|^2 - sum |F|^2) / 2, DNN = W2.relu(W1.Fflat + b1) + b2."""
    rows, E, T = 63, 8, 4
    F = T + 1
    rng = np.random.default_rng(0)
    params = rng.normal(0, 0.3, rows * E + F * F * E + 2 * F + 1).astype(np.float32)
    m = orc.TdmModel(params, rows, E, T, deepfm=True)
    node = np.array([5, 7, -1, 62], np.int32)
    seq = np.array([[1, 2, -1, 3], [4, 4, 4, 4], [0, 1, 2, 3], [-1, -1, -1, -1]], np.int32)
    out = m.forward(node, seq)
    p = params.astype(np.float64)
    emb = p[:rows * E].reshape(rows, E)
    w1 = p[rows * E:rows * E + F * F * E].reshape(F, F * E)
    b1 = p[rows * E + F * F * E:rows * E + F * F * E + F]
    w2 = p[rows * E + F * F * E + F:rows * E + F * F * E + 2 * F]
    b2 = p[-1]
    row = lambda c: np.zeros(E) if c < 0 else emb[c]
    for r in range(len(node)):
        Fm = np.stack([row(node[r])] + [row(c) for c in seq[r]])
        fm = (np.sum(Fm.sum(0) ** 2) - np.sum(Fm ** 2)) / 2
        ref = fm + np.maximum(w1 @ Fm.ravel() + b1, 0) @ w2 + b2
        assert abs(out[r] - ref) <= 2e-6 * max(1.0, abs(ref))
    with pytest.raises(IndexError):
        m.forward(np.array([rows], np.int32), seq[:1])


def test_otm_deepfm_oracle_against_independent_float64(orc):
    """OTM's DeepFM (otm/.../model/DeepFM.scala:12-48: the same graph for Double) vs a numpy evaluation of the layer
    definitions (different summation order: 1e-12), and its beam search against a search written with numpy sorts."""
    leaf_level, E, T = 6, 8, 4
    rows, F = (1 << (leaf_level + 1)) - 1, T + 1
    rng = np.random.default_rng(3)
    params = rng.normal(0, 0.3, rows * E + F * F * E + 2 * F + 1)
    m = orc.OtmModel(params, rows, E, T, deepfm=True)
    emb = params[:rows * E].reshape(rows, E)
    w1 = params[rows * E:rows * E + F * F * E].reshape(F, F * E)
    b1 = params[rows * E + F * F * E:rows * E + F * F * E + F]
    w2 = params[rows * E + F * F * E + F:rows * E + F * F * E + 2 * F]
    b2 = params[-1]
    row = lambda c: np.zeros(E) if c < 0 else emb[c]

    def ref(c, s):
        Fm = np.stack([row(c)] + [row(x) for x in s])
        return (np.sum(Fm.sum(0) ** 2) - np.sum(Fm ** 2)) / 2 + np.maximum(w1 @ Fm.ravel() + b1, 0) @ w2 + b2
    node = np.array([5, 7, -1, rows - 1], np.int32)
    seq = np.array([[70, 80, -1, 90], [64, 64, 64, 64], [100, 101, 102, 103], [-1, -1, -1, -1]], np.int32)
    out = m.forward(node, seq)
    for r in range(len(node)):
        assert abs(out[r] - ref(node[r], seq[r])) <= 1e-12 * max(1.0, abs(ref(node[r], seq[r])))
    # CandidateSearcher.beamSearch: level floor(log2 beam) whole, then stable top-beam and both children per level
    beam, s = 5, seq[0]
    ids, sc = m.beam_search(s, leaf_level, beam)
    cur = list(range(3, 7))
    scores = None
    for level in range(2, leaf_level):
        if scores is not None:
            order = sorted(range(len(cur)), key=lambda i: (-scores[i], i))[:beam]
            cur = [cur[i] for i in order]
        cur = [c for p in cur for c in (2 * p + 1, 2 * p + 2)]
        scores = [out_ for out_ in m.forward(np.array(cur, np.int32), np.tile(s, (len(cur), 1)))]
    assert list(ids) == cur and (sc == np.array(scores)).all()


def _bce_numpy(params, rows, E, T, node, seq, masked, labels):
    """BCECriterionWithLogits (scalann/.../nn/BCECriterionWithLogits.scala:28-91): mean(max(x,0) - x z + log(1 + exp(-|x|)))"""
    x = _din_numpy(params, rows, E, T, node, seq, masked)
    return float(np.mean(np.maximum(x, 0.0) - x * labels + np.log1p(np.exp(-np.abs(x)))))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_gradients_against_finite_differences(orc, dtype):
    """The backward of every DIN layer + the criterion (orc_train.inc restates EmbeddingShare / LookupTable.updateEmbeddings,
    Attention, Mask, SoftMax, MatMul, Concat, Linear, ReLU updateGradInput / accGradParameters) checked independently:
    central differences of an independent float64 loss, for every dense parameter class and for touched / untouched table rows."""
    rng = np.random.default_rng(5)
    rows, E, T, n = 63, 16, 6, 24
    params = rng.normal(0, 0.3, rows * E + 3 * E * E + 2 * E + 1)
    node = rng.integers(0, rows, n).astype(np.int32)
    seq = rng.integers(0, rows, (n, T)).astype(np.int32)
    seq[rng.random((n, T)) < 0.25] = -1
    seq[0] = -1                                      # a fully padded history
    node[1] = seq[1, 2] = 7                          # the same row as item and as history of one sample
    masked = seq == -1
    mask_flat = np.flatnonzero(masked.ravel()).astype(np.int32)
    labels = (rng.random(n) < 0.4).astype(np.float64)
    grad, loss = orc.din_gradients(params.astype(dtype), rows, E, T, node, seq, mask_flat, labels.astype(dtype))
    assert abs(float(loss) - _bce_numpy(params, rows, E, T, node, seq, masked, labels)) < (1e-12 if dtype == np.float64 else 2e-6)
    o_w = rows * E
    probe = list(rng.choice(rows * E, 60, replace=False)) + list(o_w + rng.choice(3 * E * E, 60, replace=False)) + \
        list(range(o_w + 3 * E * E, o_w + 3 * E * E + 2 * E + 1, 3)) + [7 * E + 3, int(node[0]) * E]
    h = 1e-6
    worst = 0.0
    for i in probe:
        pp, pm = params.copy(), params.copy()
        pp[i] += h; pm[i] -= h
        fd = (_bce_numpy(pp, rows, E, T, node, seq, masked, labels) - _bce_numpy(pm, rows, E, T, node, seq, masked, labels)) / (2 * h)
        worst = max(worst, abs(fd - float(grad[i])))
    assert worst < (2e-8 if dtype == np.float64 else 3e-6), worst
    used = np.zeros(rows, bool)
    used[node] = True
    used[seq[seq >= 0]] = True
    g_emb = np.asarray(grad[:rows * E]).reshape(rows, E)
    assert (g_emb[~used] == 0).all() and np.abs(g_emb[used]).sum() > 0          # scatter-add touches gathered rows only


def test_oracle_adam_closed_form(orc):
    """Adam.optimize (scalann/.../optim/Adam.scala:54-65): s = b1 s + (1-b1) g; r = b2 r + (1-b2) g^2; denom = sqrt(r) + eps;
    w -= lr sqrt(1-b2^t)/(1-b1^t) s / denom, dense over every parameter (zero gradients still decay the moments)."""
    rng = np.random.default_rng(9)
    n = 1000
    w = rng.normal(size=n); g = rng.normal(size=n); g[::7] = 0.0
    s = rng.normal(size=n) * 0.1; r = np.abs(rng.normal(size=n)) * 0.01
    for dtype, tol in ((np.float64, 1e-15), (np.float32, 2e-7)):
        for t in (1, 2, 50):
            wv, sv, rv = w.astype(dtype), s.astype(dtype), r.astype(dtype)
            orc.adam_step(wv, g.astype(dtype), sv, rv, 1e-3, t)
            s2 = 0.9 * s + 0.1 * g
            r2 = 0.999 * r + 0.001 * g * g
            w2 = w - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * s2 / (np.sqrt(r2) + 1e-8)
            assert np.abs(sv - s2).max() < tol * 10 and np.abs(rv - r2).max() < tol * 10
            assert np.abs(wv - w2).max() < tol * 10 * max(1.0, np.abs(w2).max())
            assert (sv[::7] == (0.9 * s.astype(dtype)[::7]).astype(dtype)).all() or dtype == np.float32   # zero gradient: pure decay


def _pseudo_targets_python(model, seqs, targets, leaf_level, start_level, use_mask):
    """Independent restatement of OTMTree.optimalPseudoTargets with the Scala driver's List / Map bookkeeping (dicts)."""
    T = seqs.shape[1]
    out = [[{int(i): 1.0 for i in t} for t in targets]]
    for _ in range(leaf_level - 1, start_level, -1):
        children = out[0]
        parents = []
        for u, ch in enumerate(children):
            nodes = list(ch.keys())
            if not nodes:
                parents.append({})
                continue
            pos = np.array(nodes, np.int32)
            neg = np.where(pos % 2 == 0, pos - 1, pos + 1).astype(np.int32)
            seq = np.tile(seqs[u], (len(nodes), 1))
            mask = np.flatnonzero((seq == -1).ravel()).astype(np.int32) if use_mask else None
            pp = model.forward(pos if use_mask else neg, seq, mask)
            pn = model.forward(neg, seq, mask)
            acc = {}
            for k, n in enumerate(nodes):
                label = ch[n] if pp[k] >= pn[k] else ch.get(int(neg[k]), 0.0)
                acc[(n - 1) >> 1] = acc.get((n - 1) >> 1, 0.0) + label
            parents.append({p: min(1.0, max(0.0, v)) for p, v in acc.items()})
        out.insert(0, parents)
    return out


@pytest.mark.parametrize("use_mask", [True, False])
def test_otm_pseudo_targets_oracle(orc, otm_fix, use_mask):
    """orc_otm_pseudo_targets (OTMTree.scala:27-46, 104-172) against the dict-based restatement: users with several targets,
    siblings both listed, shared ancestors (sums above 1 that must clip), padded histories, both useMask settings."""
    f = otm_fix
    n = len(f["items"])
    leaf_level = int(np.ceil(np.log(n) / np.log(2)))
    model = orc.OtmModel(f["params"], 8191, int(f["E"]), int(f["T"]))
    rng = np.random.default_rng(11)
    B, T, start_level = 24, int(f["T"]), 4
    leaf_ids = f["leaf_ids"].astype(np.int32)
    seqs = rng.choice(leaf_ids, (B, T)).astype(np.int32)
    seqs[rng.random((B, T)) < 0.3] = -1
    seqs[0] = -1
    targets = []
    for u in range(B):
        k = int(rng.integers(1, 9))
        t = rng.choice(leaf_ids, k, replace=False).tolist()
        if u % 3 == 0:                                   # both children of one parent, and a cousin pair
            base = int(t[0]) - (1 if int(t[0]) % 2 == 0 else 0)
            t += [base, base + 1]
        if u % 5 == 0:
            t.append(t[0])                               # a duplicated target
        targets.append([int(x) for x in t if (1 << leaf_level) - 1 <= x < (2 << leaf_level) - 1])
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum([len(t) for t in targets])
    flat = np.concatenate([np.array(t, np.int32) for t in targets])
    ids, vals, cnt = model.pseudo_targets(seqs, off, flat, leaf_level, start_level, use_mask, M=12)
    want = _pseudo_targets_python(model, seqs, targets, leaf_level, start_level, use_mask)
    assert ids.shape[0] == len(want) == leaf_level - start_level
    clipped = 0
    for li in range(ids.shape[0]):
        for u in range(B):
            got = {int(i): float(v) for i, v in zip(ids[li, u, :cnt[li, u]], vals[li, u, :cnt[li, u]])}
            assert got == want[li][u], (li, u)
            assert (np.diff(ids[li, u, :cnt[li, u]]) > 0).all() and (ids[li, u, cnt[li, u]:] == -1).all()
            clipped += sum(1 for v in got.values() if v == 1.0) + sum(1 for v in got.values() if v == 0.0)
    assert clipped > 0


# ---- Deep Retrieval training (oracle/oracle_dr_train.c) -------------------------------------------------------------------
def test_cross_entropy_reference_vectors(orc):
    """scalann/src/test/scala/CrossEntropyTest.scala:26-43: CrossEntropyCriterion forward 1.3200 and backward, the Scala test's 1e-4."""
    logits = np.array([[2.3, -10.2], [0.8, -3.1], [-2.2, 1.0]])
    loss, grad = orc.cross_entropy(logits, [0, 1, 1])
    assert abs(loss - 1.3200) < 1e-4
    want = np.array([-1.2219e-06, 1.2422e-06, 3.2672e-01, -3.2672e-01, 1.3055e-02, -1.3055e-02]).reshape(3, 2)
    assert grad.shape == logits.shape and np.abs(grad - want).max() < 1e-4
    # float64 closed form: mean over rows of logsumexp - logit[target]; gradient (softmax - onehot) / R
    z = logits - logits.max(1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(1, keepdims=True)
    oh = np.eye(2)[[0, 1, 1]]
    assert abs(loss - (-np.log(p[np.arange(3), [0, 1, 1]]).mean())) < 1e-14
    assert np.abs(grad - (p - oh) / 3).max() < 1e-15


def test_sampled_softmax_reference_property(orc):
    """scalann/src/test/scala/SampledSoftmaxLossTest.scala:8-52: with fixed sampledValues the loss decreases over 7 forward / backward
    rounds (the criterion updates its own weights with Adam, lr 7e-3, on gradients it never zeroes), gradInput has inputVecs' shape."""
    rng = np.random.default_rng(2022)
    B, E, S, n_cls, lr = 6, 10, 4, 200, 7e-3
    u = rng.uniform(-0.05, 0.05, (B, E))
    w = rng.normal(0.0, 0.01, (n_cls, E))
    b = np.zeros(n_cls)
    pos = [0, 1, 3, 2, 77, 101]
    neg = [[19, 3, 66, 190], [33, 4, 88, 111], [2, 48, 92, 129], [1, 66, 34, 167], [53, 11, 0, 123], [8, 99, 100, 12]]
    sampled = np.array([[p] + n for p, n in zip(pos, neg)], np.int32)
    gw, gb = np.zeros_like(w), np.zeros_like(b)
    st = [np.zeros_like(w), np.zeros_like(w), np.zeros_like(b), np.zeros_like(b)]
    losses = []
    for t in range(1, 8):
        loss, gu = orc.sampled_softmax(u, w, b, sampled, gw, gb)
        assert gu.shape == u.shape
        orc.adam_eps(w.reshape(-1), gw.reshape(-1), st[0].reshape(-1), st[1].reshape(-1), lr, 1e-7, t)
        orc.adam_eps(b, gb, st[2], st[3], lr, 1e-7, t)
        losses.append(loss)
    assert all(a > b_ for a, b_ in zip(losses, losses[1:])), losses
    assert abs(losses[0] - np.log(S + 1)) < 1e-3                       # near-zero logits: uniform over the 5 sampled classes


def _dr_toy(seed, num_item=40, K=7, D=3, T=4, E=6, P=2):
    rng = np.random.default_rng(seed)
    mk = lambda *s: rng.normal(0.0, 0.3, s)
    m = dict(num_item=num_item, K=K, D=D, T=T, E=E, layer_emb=mk(num_item + K * (D - 1), E), layer_w=[mk(K, (T + d) * E) for d in range(D)],
             layer_b=[mk(K) for _ in range(D)], rr_emb=mk(num_item, E), rr_w=mk(E, T * E), rr_b=mk(E), sm_w=mk(num_item, E), sm_b=mk(num_item))
    item_paths = rng.integers(0, K, (num_item, P, D)).astype(np.int32)
    return m, item_paths


def _dr_layer_loss_f64(m, seq, target, item_paths, emb=None, w=None, b=None):
    """independent float64 numpy restatement of the layer model's mean cross entropies (sum over layers)"""
    emb = m["layer_emb"] if emb is None else emb
    w = m["layer_w"] if w is None else w
    b = m["layer_b"] if b is None else b
    tot = 0.0
    rows = [(s, p) for s in range(len(seq)) for p in range(item_paths.shape[1])]
    for d in range(m["D"]):
        ls = []
        for s, p in rows:
            path = item_paths[target[s], p]
            idx = list(seq[s]) + [path[i] + m["num_item"] + i * m["K"] for i in range(d)]
            x = np.concatenate([emb[c] if c >= 0 else np.zeros(m["E"]) for c in idx])
            z = w[d] @ x + b[d]
            ls.append(np.log(np.exp(z - z.max()).sum()) + z.max() - z[path[d]])
        tot += np.mean(ls)
    return tot


def test_dr_layer_gradients_by_finite_differences(orc):
    """orc_dr_layer_grad (trainLayerBatch, LocalOptimizer.scala:139-168) against central differences of an independent float64
    numpy loss: embedding rows (history and path-node rows), Linear weights and biases of every layer; padded histories."""
    m, item_paths = _dr_toy(3)
    rng = np.random.default_rng(5)
    n = 9
    seq = rng.integers(0, m["num_item"], (n, m["T"])).astype(np.int32)
    seq[rng.random(seq.shape) < 0.25] = -1
    target = rng.integers(0, m["num_item"], n).astype(np.int32)
    model = orc.DrModel(**m)
    tr = orc.DrTrainer(model, 1e-3)
    g, loss = tr.layer_grad(seq, target, item_paths, item_paths.shape[1])
    assert abs(loss.sum() - _dr_layer_loss_f64(m, seq, target, item_paths)) < 1e-12
    h = 1e-6
    checks = [("layer_emb", None, g[0])] + [("layer_w", d, g[1 + 2 * d]) for d in range(m["D"])] + [("layer_b", d, g[2 + 2 * d]) for d in range(m["D"])]
    for name, d, grad in checks:
        arr = m[name] if d is None else m[name][d]
        flat = arr.reshape(-1)
        for k in rng.choice(flat.size, min(12, flat.size), replace=False):
            old = flat[k]
            flat[k] = old + h
            up = _dr_layer_loss_f64(m, seq, target, item_paths)
            flat[k] = old - h
            dn = _dr_layer_loss_f64(m, seq, target, item_paths)
            flat[k] = old
            assert abs((up - dn) / (2 * h) - grad.reshape(-1)[k]) < 1e-7, (name, d, k)
    # thread chunks (syncGradients): equal chunk sizes -> the same mean gradient up to rounding; unequal -> mean of chunk means
    g3, loss3 = tr.layer_grad(seq, target, item_paths, item_paths.shape[1], parallelism=3)
    assert all(np.abs(a - b_).max() < 1e-14 for a, b_ in zip(g, g3)) and np.abs(loss - loss3).max() < 1e-14
    g2, loss2 = tr.layer_grad(seq, target, item_paths, item_paths.shape[1], parallelism=2)       # chunks of 5 and 4 samples
    ga, la = tr.layer_grad(seq[:5], target[:5], item_paths, item_paths.shape[1])
    gb_, lb = tr.layer_grad(seq[5:], target[5:], item_paths, item_paths.shape[1])
    assert all(np.abs(x - (a + b_) / 2).max() < 1e-15 for x, a, b_ in zip(g2, ga, gb_)) and np.abs(loss2 - (la + lb) / 2).max() < 1e-15


def test_dr_rerank_gradients_by_finite_differences(orc):
    """orc_dr_rerank_grad (trainRerank, LocalOptimizer.scala:122-137): model gradients against central differences of a float64
    numpy sampled-softmax loss; softmax-parameter gradients accumulate across calls (ParameterOptimizer never zeroes them)."""
    m, _ = _dr_toy(4)
    rng = np.random.default_rng(6)
    n, S = 8, 5
    seq = rng.integers(0, m["num_item"], (n, m["T"])).astype(np.int32)
    seq[rng.random(seq.shape) < 0.25] = -1
    target = rng.integers(0, m["num_item"], n).astype(np.int32)
    sampled = np.array([[t] + sorted(rng.choice([x for x in range(m["num_item"]) if x != t], S, replace=False).tolist()) for t in target], np.int32)

    def loss_f64():
        ls = []
        for i in range(n):
            x = np.concatenate([m["rr_emb"][c] if c >= 0 else np.zeros(m["E"]) for c in seq[i]])
            u = m["rr_w"] @ x + m["rr_b"]
            z = m["sm_w"][sampled[i]] @ u + m["sm_b"][sampled[i]]
            ls.append(np.log(np.exp(z - z.max()).sum()) + z.max() - z[0])
        return np.mean(ls)
    tr = orc.DrTrainer(orc.DrModel(**m), 1e-3)
    g, loss = tr.rerank_grad(seq, sampled)
    assert abs(loss - loss_f64()) < 1e-13
    h = 1e-6
    for name, grad in [("rr_emb", g[0]), ("rr_w", g[1]), ("rr_b", g[2]), ("sm_w", tr.sm_g[0]), ("sm_b", tr.sm_g[1])]:
        flat = m[name].reshape(-1)
        for k in rng.choice(flat.size, min(12, flat.size), replace=False):
            old = flat[k]
            flat[k] = old + h
            up = loss_f64()
            flat[k] = old - h
            dn = loss_f64()
            flat[k] = old
            assert abs((up - dn) / (2 * h) - grad.reshape(-1)[k]) < 1e-7, (name, k)
    once = [x.copy() for x in tr.sm_g]
    tr.rerank_grad(seq, sampled)
    assert all(np.abs(x - 2 * o).max() < 1e-15 for x, o in zip(tr.sm_g, once))


# ---- k-means tree rebuild (oracle/oracle_cluster.c) -----------------------------------------------------------------------------
def _arg_partition_python(elems, position):
    """line-by-line port of Utils.argPartition (tdm/.../utils/Utils.scala:130-199)"""
    e = list(elems)
    ix = list(range(len(e)))

    def swap(a, b):
        e[a], e[b] = e[b], e[a]
        ix[a], ix[b] = ix[b], ix[a]

    def med(p1, p2, p3):
        if e[p1] < e[p2]:
            return p2 if e[p2] < e[p3] else (p3 if e[p1] < e[p3] else p1)
        return p2 if e[p2] > e[p3] else (p3 if e[p1] > e[p3] else p1)
    left, right = 0, len(e) - 1
    while left < right:
        pvt = med(left, right, (left + right) // 2)
        pv = e[pvt]
        swap(pvt, left)
        i, lt, gt = left, left, right
        while i <= gt:
            if e[i] < pv:
                swap(lt, i); lt += 1; i += 1
            elif e[i] > pv:
                swap(gt, i); gt -= 1
            else:
                i += 1
        if lt <= position <= gt:
            left = right
        elif position < lt:
            right = lt - 1
        else:
            left = gt + 1
    return e, ix


def test_arg_partition_matches_the_scala_algorithm(orc):
    rng = np.random.default_rng(31)
    for n in [2, 3, 4, 5, 8, 17, 64, 257, 1000]:
        for trial in range(4):
            d = rng.random(n) if trial < 2 else rng.integers(0, 4, n).astype(np.float64)       # duplicates exercise the == branch
            mid = n // 2
            got_e, got_ix = orc.arg_partition(d, mid)
            want_e, want_ix = _arg_partition_python(d.tolist(), mid)
            assert got_ix.tolist() == want_ix and got_e.tolist() == want_e
            # balanceTree's contract: the left half holds the mid smallest distances
            assert sorted(got_ix.tolist()) == list(range(n)) and (d[got_ix] == got_e).all()
            assert got_e[:mid].max() <= got_e[mid:].min()
    with pytest.raises(ValueError):
        orc.arg_partition(np.array([1.0, np.nan, 0.5, 2.0]), 2)


def test_kmeans_tree_oracle_properties(orc):
    """RecursiveCluster.run: every point gets its own code, the tree is balanced (sibling subtrees differ by at most one leaf, depth
    floor/ceil(log2 n)), two well separated blobs are split at the root, the result is a function of the seed."""
    rng = np.random.default_rng(32)
    for n in [2, 3, 7, 100, 257, 1500]:
        emb = rng.random((n, 6))
        codes = orc.kmeans_tree(emb, 3, 7)
        assert len(set(codes.tolist())) == n and (codes > 0).all()
        level = np.floor(np.log2(codes + 1)).astype(int)
        assert level.min() >= int(np.floor(np.log2(n))) and level.max() <= int(np.ceil(np.log2(n)))
        anc = codes.copy()                                     # leaf counts under the two children of the root differ by <= 1
        while (anc > 2).any():
            anc = np.where(anc > 2, (anc - 1) // 2, anc)
        assert abs(int((anc == 1).sum()) - int((anc == 2).sum())) <= 1
        assert (orc.kmeans_tree(emb, 3, 7) == codes).all()
    blobs = np.concatenate([rng.normal(0.0, 0.1, (64, 4)), rng.normal(5.0, 0.1, (64, 4))])
    codes = orc.kmeans_tree(blobs, 2, 1)
    anc = codes.copy()
    while (anc > 2).any():
        anc = np.where(anc > 2, (anc - 1) // 2, anc)
    assert len(set(anc[:64].tolist())) == 1 and len(set(anc[64:].tolist())) == 1 and anc[0] != anc[64]


def test_deepfm_gradients_by_finite_differences(orc):
    """orc_deepfm_gradients_f32 (DeepFM.scala:11-44 + FM.scala:46-72 backward, BCE mean) against central differences of an independent
    float64 numpy loss on every parameter group; padded history slots and a padded item row get no gradient."""
    rng = np.random.default_rng(41)
    rows, E, T, n = 63, 8, 4, 13
    F, IN = T + 1, (T + 1) * E
    params = np.concatenate([rng.normal(0, 0.4, rows * E), rng.normal(0, 0.3, F * IN), rng.normal(0, 0.2, F), rng.normal(0, 0.5, F),
                             [0.1]]).astype(np.float32)
    node = rng.integers(0, rows, n).astype(np.int32)
    seq = rng.integers(0, rows, (n, T)).astype(np.int32)
    seq[rng.random((n, T)) < 0.3] = -1
    labels = (rng.random(n) < 0.4).astype(np.float32)

    def loss64(p):
        p = p.astype(np.float64)
        emb, w1 = p[:rows * E].reshape(rows, E), p[rows * E:rows * E + F * IN].reshape(F, IN)
        b1, w2, b2 = p[rows * E + F * IN:][:F], p[rows * E + F * IN + F:][:F], p[-1]
        tot = 0.0
        for r in range(n):
            X = np.stack([emb[c] if c >= 0 else np.zeros(E) for c in [node[r]] + seq[r].tolist()])
            fm = ((X.sum(0) ** 2).sum() - (X ** 2).sum()) / 2
            y = fm + np.maximum(w1 @ X.ravel() + b1, 0) @ w2 + b2
            tot += max(y, 0) - y * labels[r] + np.log1p(np.exp(-abs(y)))
        return tot / n
    g, loss = orc.deepfm_gradients(params, rows, E, T, node, seq, labels)
    assert abs(loss - loss64(params)) < 1e-5
    h = 1e-3
    for lo, hi in [(0, rows * E), (rows * E, rows * E + F * IN), (rows * E + F * IN, len(params))]:
        for k in rng.choice(np.arange(lo, hi), min(25, hi - lo), replace=False):
            p = params.astype(np.float64)
            p[k] += h
            up = loss64(p)
            p[k] -= 2 * h
            dn = loss64(p)
            assert abs((up - dn) / (2 * h) - g[k]) < 2e-4 * max(1.0, abs(g[k])), k
    touched = np.zeros(rows, bool)
    touched[node] = True
    touched[seq[seq >= 0]] = True
    assert not g[:rows * E].reshape(rows, E)[~touched].any()


@pytest.mark.parametrize("E,use_mask", [(64, True), (16, True), (32, False)])
def test_tuned_cpu_form_stays_close_to_the_faithful_port(E, use_mask):
    """oracle_tuned.c (bench.py's `cpu_baseline_tuned`) re-associates the arithmetic, so it is only CLOSE to the faithful port: same
    counts, logits within 2e-5 wherever the ids agree, and the same top-k on (nearly) every user; padded / unknown histories included."""
    from oracle import oracle as orc
    from dismember_b200 import synth
    n_items, T, beam, topk, B = 6000, 10, 60, 10, 96
    tf = synth.tdm_tree(n_items, seed=4)
    rows = (1 << (tf.max_level + 1)) - 1
    params = synth.din_params(rows, E, seed=5, structured=True)
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T)
    tuned = orc.TunedTdm(tree, model, n_threads=2)
    q = synth.queries(B, T, n_items, seed=6)
    q[0] = 0                                                        # all padding: every position masked
    q[1, :5] = 0
    fi, fl, fc = model.retrieve_batch(tree, q, beam, topk, use_mask=use_mask, n_threads=2)
    ti, tl, tc = tuned.retrieve_batch(q, beam, topk, use_mask=use_mask, n_threads=3)
    assert (fc == tc).all()
    same = fi == ti
    assert same.all(1).mean() >= 0.97
    assert np.abs(fl[same] - tl[same]).max() < 2e-5
