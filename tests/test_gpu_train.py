"""Training path (K4-K7) and JTM weights (K9) against the oracle.

Gradients are accumulated with atomics on the GPU and row-by-row on the CPU, so the tolerance is
1e-5 relative to the largest gradient entry for fp32 and 1e-11 for fp64; Adam is
elementwise.  The sampler's RNG cannot match the reference's ThreadLocalRandom, so K4 is checked
through the properties NegativeSampler guarantees."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch(rng, rows, T, n):
    node = rng.integers(0, rows, n).astype(np.int32)
    seq = rng.integers(0, rows, (n, T)).astype(np.int32)
    seq[rng.random((n, T)) < 0.3] = -1
    seq[0] = -1
    mask = np.flatnonzero((seq == -1).ravel()).astype(np.int32)
    labels = (rng.random(n) < 0.2).astype(np.float64)
    return node, seq, mask, labels


@pytest.mark.parametrize("which", ["f32", "f64"])
def test_gradients_match_oracle(engine, orc, jtm_fix, otm_fix, which):
    fix = jtm_fix if which == "f32" else otm_fix
    params = fix["params"]
    engine.load_din_weights(params, 8191, 16, 10)
    rng = np.random.default_rng(5)
    node, seq, mask, labels = _batch(rng, 8191, 10, 700)
    g, loss = engine.din_gradients(node, seq, mask, labels)
    og, oloss = orc.din_gradients(params, 8191, 16, 10, node, seq, mask, labels.astype(params.dtype))
    tol = 1e-5 if which == "f32" else 1e-11
    assert abs(loss - oloss) <= tol * max(1.0, abs(oloss))
    assert np.abs(g - og).max() <= tol * np.abs(og).max()
    # untouched rows receive exactly zero gradient
    touched = np.zeros(8191, bool)
    touched[node] = True
    touched[seq[seq >= 0]] = True
    assert not g[:8191 * 16].reshape(8191, 16)[~touched].any()


def test_gradients_e64_structured(engine, orc):
    from dismember_b200 import synth
    rows, E, T = 4095, 64, 10
    params = synth.din_params(rows, E, seed=3)
    engine.load_din_weights(params, rows, E, T)
    rng = np.random.default_rng(6)
    node, seq, mask, labels = _batch(rng, rows, T, 333)
    g, loss = engine.din_gradients(node, seq, mask, labels)
    og, oloss = orc.din_gradients(params, rows, E, T, node, seq, mask, labels.astype(np.float32))
    assert abs(loss - oloss) <= 1e-5 * max(1.0, abs(oloss))
    assert np.abs(g - og).max() <= 2e-5 * np.abs(og).max()


@pytest.mark.parametrize("which", ["f32", "f64"])
def test_train_steps_match_oracle(engine, orc, jtm_fix, otm_fix, which):
    """LocalOptimizer: zeroGrad, fwd, BCE, bwd, dense Adam -- three consecutive steps."""
    fix = jtm_fix if which == "f32" else otm_fix
    params = fix["params"].copy()
    dt = params.dtype
    engine.load_din_weights(params, 8191, 16, 10)
    rng = np.random.default_rng(7)
    w = params.copy()
    s = np.zeros_like(w)
    r = np.zeros_like(w)
    lr = 1e-3
    for t in range(1, 4):
        node, seq, mask, labels = _batch(rng, 8191, 10, 512)
        loss = engine.train_step(node, seq, mask, labels, lr, t)
        og, oloss = orc.din_gradients(w, 8191, 16, 10, node, seq, mask, labels.astype(dt))
        orc.adam_step(w, og, s, r, lr, t)
        assert abs(loss - oloss) <= (1e-5 if which == "f32" else 1e-11) * max(1.0, abs(oloss))
    got = engine.download_din_weights()
    # Adam's first steps move every touched weight by ~lr regardless of the gradient size, so compare
    # the UPDATE (w - w0), where a 1e-5 relative gradient error shows up amplified near g ~ 0
    upd, oupd = got - params, w - params
    close = np.abs(upd - oupd) <= (2e-2 if which == "f32" else 1e-6) * lr + 1e-12
    assert close.mean() > 0.999
    # the refreshed transposed weights are what retrieval uses after a step
    node, seq, mask, _ = _batch(rng, 8191, 10, 64)
    fw = engine.score_pairs(node, seq, mask)
    model = (orc.TdmModel if which == "f32" else orc.OtmModel)(got, 8191, 16, 10)
    assert (fw == model.forward(node, seq, mask)).all()


def test_adam_is_dense_like_the_reference(engine, orc, jtm_fix):
    """Adam.scala updates EVERY parameter each step: a row touched at step 1 keeps moving at step 2
    although its gradient is zero there (moments decay), a lazy/sparse Adam would not."""
    params = jtm_fix["params"].copy()
    engine.load_din_weights(params, 8191, 16, 10)
    seq = np.full((1, 10), -1, np.int32)
    engine.train_step(np.array([5], np.int32), seq, np.arange(10, dtype=np.int32), np.array([1.0]), 1e-2, 1)
    w1 = engine.download_din_weights()
    engine.train_step(np.array([9], np.int32), seq, np.arange(10, dtype=np.int32), np.array([1.0]), 1e-2, 2)
    w2 = engine.download_din_weights()
    row5_1, row5_2 = w1[5 * 16:6 * 16], w2[5 * 16:6 * 16]
    assert (row5_1 != params[5 * 16:6 * 16]).any()
    assert (row5_2 != row5_1).any()            # moved again with zero gradient
    assert (w2[7 * 16:8 * 16] == params[7 * 16:8 * 16]).all()   # never touched: s = r = 0 -> no move


def test_sampler_properties(engine, jtm_fix):
    f = jtm_fix
    engine.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    engine.load_din_weights(f["params"], 8191, 16, 10)
    L = int(f["max_level"])
    layer_neg = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], np.int32)      # configs/tdm.conf prefix
    rng = np.random.default_rng(8)
    targets = rng.choice(f["leaf_ids"], 200).astype(np.int32)
    seqs = rng.choice(f["leaf_ids"], (200, 10)).astype(np.int32)
    seqs[:, :3] = 0
    node, oseq, lab = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=99)
    layer_sum = int(sum(1 + x for x in layer_neg[1:]))                              # NegativeSampler.scala:57
    assert len(node) == 200 * layer_sum
    exists = set(f["codes"].tolist())
    code_of = dict(zip(f["leaf_ids"].tolist(), f["leaf_codes"].tolist()))
    node = node.reshape(200, layer_sum)
    lab = lab.reshape(200, layer_sum)
    for t in range(200):
        off = 0
        anc = code_of[int(targets[t])]
        path = []
        while anc > 0:
            path.append(anc)
            anc = (anc - 1) >> 1
        path = path[::-1]                                                           # TDMTree.pathNodes, root excluded
        for level in range(1, L + 1):
            k = int(layer_neg[level])
            blk, lb = node[t, off:off + 1 + k], lab[t, off:off + 1 + k]
            assert blk[0] == path[level - 1] and lb[0] == 1.0 and (lb[1:] == 0.0).all()
            negs = blk[1:]
            lo, hi = 2 ** level - 1, 2 ** (level + 1) - 1
            assert ((negs >= lo) & (negs < hi)).all() and blk[0] not in negs.tolist()
            assert len(set(negs.tolist())) == k and all(int(c) in exists for c in negs)
            assert (np.diff(negs) > 0).all()                                       # BitSet.toList: ascending
            off += 1 + k
    # histories: TDMTree.idToCode of the target's sequence, repeated layer_sum times
    oseq = oseq.reshape(200, layer_sum, 10)
    assert (oseq[:, :, :3] == -1).all() and (oseq == oseq[:, :1]).all()
    assert oseq[0, 0, 5] == code_of[int(seqs[0, 5])]
    # a different seed draws different negatives; the same seed is reproducible
    n2, _, _ = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=100)
    n3, _, _ = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=99)
    assert (n3.reshape(200, layer_sum) == node).all() and (n2.reshape(200, layer_sum) != node).any()


def test_sample_then_train_reduces_loss(engine, jtm_fix):
    """TdmModelTrainSpec-style smoke: a few Adam iterations on sampled batches reduce the loss."""
    f = jtm_fix
    engine.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    rng = np.random.default_rng(10)
    n_par = 8191 * 16 + 3 * 256 + 33
    params = (rng.normal(0, 0.05, n_par)).astype(np.float32)
    engine.load_din_weights(params, 8191, 16, 10)
    layer_neg = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], np.int32)
    targets = rng.choice(f["leaf_ids"], 64).astype(np.int32)
    seqs = rng.choice(f["leaf_ids"], (64, 10)).astype(np.int32)
    losses = []
    for t in range(1, 31):
        node, oseq, lab = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=t)
        mask = np.flatnonzero((oseq == -1).ravel()).astype(np.int32)
        losses.append(float(engine.train_step(node, oseq, mask, lab, 5e-3, t)))
    assert losses[-1] < losses[0] * 0.9


def test_jtm_item_weights(engine, orc, jtm_fix):
    f = jtm_fix
    engine.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    engine.load_din_weights(f["params"], 8191, 16, 10)
    model = orc.TdmModel(f["params"], 8191, 16, 10)
    tree = orc.Tree(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    rng = np.random.default_rng(11)
    n_items, T = 40, 10
    counts = rng.integers(0, 30, n_items)
    counts[3] = 0                                               # never a target -> -1e6
    off = np.zeros(n_items + 1, np.int64)
    off[1:] = np.cumsum(counts)
    seqs = rng.choice(f["leaf_ids"], (int(off[-1]), T)).astype(np.int32)
    seqs[rng.random(seqs.shape) < 0.2] = 0
    code_of = dict(zip(f["leaf_ids"].tolist(), f["leaf_codes"].tolist()))
    for old_level, level, hier, min_level in [(0, 3, False, 0), (4, 6, True, 5), (9, 12, True, 0)]:
        parents = rng.integers(2 ** old_level - 1, 2 ** (old_level + 1) - 1, n_items).astype(np.int32)
        got = engine.jtm_item_weights(off, seqs, parents, old_level, level, hier, min_level)
        gap = level - old_level
        want = np.zeros((n_items, 2 ** gap), np.float32)
        for i in range(n_items):
            if counts[i] == 0:
                want[i] = -1e6
                continue
            smp = seqs[off[i]:off[i + 1]]
            children = [int(parents[i])]
            for _ in range(gap):
                children = [c for n in children for c in (2 * n + 1, 2 * n + 2)]
            for ci, child in enumerate(children):
                wsum = np.float32(0)
                node, lv = child, level
                while node > parents[i]:
                    codes = np.full(smp.shape, -1, np.int32)
                    for a in range(smp.shape[0]):
                        for b in range(T):
                            iid = int(smp[a, b])
                            if iid == 0:
                                continue
                            c = code_of[iid]
                            if hier and lv >= min_level:
                                lim = 2 ** (lv + 1) - 1
                                while c >= lim:
                                    c = (c - 1) >> 1
                            codes[a, b] = c
                    logits = model.forward(np.full(len(smp), node, np.int32), codes,
                                           np.flatnonzero((smp == 0).ravel()).astype(np.int32))
                    acc = np.float32(0)
                    for v in logits:
                        acc = np.float32(acc + v)
                    wsum = np.float32(wsum + acc)
                    node = (node - 1) // 2
                    lv -= 1
                want[i, ci] = wsum
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), (old_level, level)


def test_categorical_sampler_follows_node_probabilities(engine, jtm_fix):
    """NegativeSampler.sampleFromCategoricalDistribution (tdm/.../utils/NegativeSampler.scala:116-144): negatives of a level are
    drawn from the level's Node.probality weights (levelProbs :59-66), new codes != the positive only, ascending order.  The
    reference seeds a MersenneTwister from nanoTime, so parity is distributional: chi-square of the drawn codes of one level
    against p_c / (1 - p_pos) (one negative per draw, the same target every time)."""
    f = jtm_fix
    L = int(f["max_level"])
    rng = np.random.default_rng(3)
    codes = f["codes"].astype(np.int64)
    prob = rng.gamma(0.6, 1.0, len(codes)).astype(np.float32) + 1e-3          # skewed popularity
    engine.load_tree_tdm(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"], prob=prob)
    engine.load_din_weights(f["params"], 8191, 16, 10)
    n = 40000
    target = int(f["leaf_ids"][17])
    targets = np.full(n, target, np.int32)
    seqs = np.zeros((n, 10), np.int32)
    layer_neg = np.ones(L + 1, np.int32)
    layer_neg[0] = 0
    node, _, lab = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=5, with_prob=True, tolerance=20)
    layer_sum = 2 * L
    node = node.reshape(n, layer_sum)
    assert (lab.reshape(n, layer_sum)[:, 0::2] == 1).all() and (lab.reshape(n, layer_sum)[:, 1::2] == 0).all()
    level = 6                                                                 # 64 codes, all present in the fixture tree
    lo, hi = 2 ** level - 1, 2 ** (level + 1) - 1
    pos = node[0, 2 * (level - 1)]
    negs = node[:, 2 * (level - 1) + 1]
    assert ((negs >= lo) & (negs < hi) & (negs != pos)).all()
    p_of = dict(zip(codes.tolist(), prob.astype(np.float64).tolist()))
    lv_codes = np.array([c for c in range(lo, hi) if c in p_of and c != pos])
    w = np.array([p_of[int(c)] for c in lv_codes])
    expect = n * w / w.sum()
    got = np.array([(negs == c).sum() for c in lv_codes], np.float64)
    chi2 = float(((got - expect) ** 2 / expect).sum())
    dof = len(lv_codes) - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)                     # 5 sigma of a chi-square with dof degrees of freedom
    # uniform draws on the same tree do NOT follow these weights (the test has power)
    node_u, _, _ = engine.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=5)
    got_u = np.array([(node_u.reshape(n, layer_sum)[:, 2 * (level - 1) + 1] == c).sum() for c in lv_codes], np.float64)
    assert float(((got_u - expect) ** 2 / expect).sum()) > 20 * dof
    # several negatives per level: distinct, ascending, never the positive while the tolerance holds
    layer_neg2 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], np.int32)
    node2, _, _ = engine.tdm_sample_expand(targets[:500], seqs[:500], layer_neg2, 1, seed=7, with_prob=True, tolerance=200)
    ls2 = int(sum(1 + x for x in layer_neg2[1:]))
    node2 = node2.reshape(500, ls2)
    off = 0
    for lvl in range(1, L + 1):
        k = int(layer_neg2[lvl])
        blk = node2[:, off:off + 1 + k]
        if k > 1:
            assert (np.diff(blk[:, 1:], axis=1) > 0).all()
        if lvl >= 4:
            assert (blk[:, 1:] != blk[:, :1]).all()
        off += 1 + k
    # without probabilities the withProb sampler is refused
    engine.load_tree_tdm(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    from dismember_b200 import DmgError
    with pytest.raises(DmgError):
        engine.tdm_sample_expand(targets[:4], seqs[:4], layer_neg, 1, seed=5, with_prob=True)


@pytest.mark.parametrize("which", ["f32", "f64"])
def test_dev_variants_match_the_host_entry_points(jtm_fix, otm_fix, which):
    """dmg_score_pairs_dev / dmg_train_step_dev (device buffers, no copy, no synchronisation) == dmg_score_pairs / dmg_train_step bit
    for bit on the forward, within the atomics' reordering on the step; a bad index surfaces at dmg_synchronize."""
    import torch
    from conftest import new_engine
    from dismember_b200._capi import DmgError
    fix = jtm_fix if which == "f32" else otm_fix
    params = fix["params"]
    tdt = torch.float32 if which == "f32" else torch.float64
    rng = np.random.default_rng(15)
    node, seq, mask, labels = _batch(rng, 8191, 10, 900)
    a, b = new_engine(), new_engine()
    for e in (a, b):
        e.load_din_weights(params, 8191, 16, 10)
    dev = torch.device("cuda", 0)
    t_node, t_seq = torch.from_numpy(node).to(dev), torch.from_numpy(seq).to(dev)
    t_mask = torch.from_numpy((seq == -1).astype(np.uint8)).to(dev)
    t_lab = torch.from_numpy(labels.astype(params.dtype)).to(dev)
    t_out = torch.empty(len(node), dtype=tdt, device=dev)
    t_loss = torch.zeros(1, dtype=tdt, device=dev)
    torch.cuda.synchronize()
    a.score_pairs_dev(len(node), t_node.data_ptr(), t_seq.data_ptr(), t_mask.data_ptr(), t_out.data_ptr())
    a.synchronize()
    want = b.score_pairs(node, seq, mask)
    assert (t_out.cpu().numpy().view(np.uint8) == want.view(np.uint8)).all()
    # no mask bytes = useMask false: the host form with an empty mask list
    a.score_pairs_dev(len(node), t_node.data_ptr(), t_seq.data_ptr(), 0, t_out.data_ptr())
    a.synchronize()
    assert (t_out.cpu().numpy().view(np.uint8) == b.score_pairs(node, seq, None).view(np.uint8)).all()
    for t in (1, 2, 3):
        a.train_step_dev(len(node), t_node.data_ptr(), t_seq.data_ptr(), t_mask.data_ptr(), t_lab.data_ptr(), 1e-2, t, t_loss.data_ptr())
        a.synchronize()
        loss_b = b.train_step(node, seq, mask, labels, 1e-2, t)
        tol = 1e-5 if which == "f32" else 1e-11
        assert abs(float(t_loss.item()) - float(loss_b)) <= tol * max(1.0, abs(float(loss_b)))
    wa, wb = a.download_din_weights(), b.download_din_weights()
    assert np.abs(wa - wb).max() <= (2e-4 if which == "f32" else 1e-9)
    assert np.abs(wa - params).max() > 1e-3
    bad = t_node.clone()
    bad[5] = 9000
    a.score_pairs_dev(len(node), bad.data_ptr(), t_seq.data_ptr(), t_mask.data_ptr(), t_out.data_ptr())
    with pytest.raises(DmgError):
        a.synchronize()
    a.synchronize()                                               # the flag is reported once
    a.close()
    b.close()


def test_dp_train_step_world1_is_train_step(jtm_fix):
    """dmg_dp_train_step on a communicator of one rank (LocalOptimizer.syncGradients with a single thread) == dmg_train_step: same
    loss, same weights bit for bit after three steps.  The 2-GPU form runs in tests/test_gpu_shard.py when two GPUs are visible."""
    from conftest import new_engine
    params = jtm_fix["params"]
    rng = np.random.default_rng(16)
    a, b = new_engine(), new_engine()
    a.shard_init(1, 0)
    for e in (a, b):
        e.load_din_weights(params, 8191, 16, 10)
    for t in (1, 2, 3):
        node, seq, mask, labels = _batch(rng, 8191, 10, 256)
        la = a.dp_train_step(node, seq, mask, labels, 1e-2, t)
        lb = b.train_step(node, seq, mask, labels, 1e-2, t)
        assert abs(float(la) - float(lb)) <= 1e-5 * max(1.0, abs(float(lb)))
    wa, wb = a.download_din_weights(), b.download_din_weights()
    assert np.abs(wa - wb).max() <= 2e-4 and np.abs(wa - params).max() > 1e-3
    a.close()
    b.close()


@pytest.mark.parametrize("which", ["f32", "f64"])
def test_adam_kernel_is_bit_exact_given_the_gradient(engine, orc, jtm_fix, otm_fix, which):
    """adam_dense_kernel vs Adam.optimize (scalann/.../optim/Adam.scala:19-73) with the SAME gradient: a one-row batch without repeated
    ids has no summation-order freedom, so dmg_din_gradients and dmg_train_step see identical gradient bits and the updated weights
    must equal the oracle's Adam applied to that gradient bit for bit, three steps in a row (moments included)."""
    fix = jtm_fix if which == "f32" else otm_fix
    params = fix["params"].copy()
    engine.load_din_weights(params, 8191, 16, 10)
    w, s, r = params.copy(), np.zeros_like(params), np.zeros_like(params)
    rng = np.random.default_rng(31)
    for t in (1, 2, 3):
        ids = rng.choice(8191, 11, replace=False).astype(np.int32)
        node, seq = ids[:1], ids[1:].reshape(1, 10).copy()
        seq[0, 7:] = -1
        mask = np.arange(7, 10, dtype=np.int32)
        labels = np.array([1.0 if t % 2 else 0.0])
        g, _ = engine.din_gradients(node, seq, mask, labels)
        orc.adam_step(w, g, s, r, 3e-3, t)
        engine.train_step(node, seq, mask, labels, 3e-3, t)
        got = engine.download_din_weights()
        assert (got.view(np.uint8) == w.view(np.uint8)).all(), t
