"""Deep Retrieval training step on the GPU (csrc/dr_train.cu) against the oracle (oracle/oracle_dr_train.c):
deep-retrieval/.../optim/LocalOptimizer.scala:58-194.  Logits, Linear weight / bias gradients follow the oracle's fma chains; `log`
and the atomic scatter-adds differ in the last bits, hence 1e-12 relative."""
import numpy as np
import pytest

from conftest import new_engine

pytestmark = pytest.mark.gpu


def _model(seed, num_item, K, D, T, E, P):
    rng = np.random.default_rng(seed)
    mk = lambda *s: rng.normal(0.0, 0.3, s)
    m = dict(num_item=num_item, K=K, D=D, T=T, E=E, layer_emb=mk(num_item + K * (D - 1), E), layer_w=[mk(K, (T + d) * E) for d in range(D)],
             layer_b=[mk(K) for _ in range(D)], rr_emb=mk(num_item, E), rr_w=mk(E, T * E), rr_b=mk(E), sm_w=mk(num_item, E) * 0.2, sm_b=np.zeros(num_item))
    return m, rng.integers(0, K, (num_item, P, D)).astype(np.int32)


def _batch(rng, m, n, S):
    seq = rng.integers(0, m["num_item"], (n, m["T"])).astype(np.int32)
    seq[rng.random(seq.shape) < 0.25] = -1
    seq[0] = -1
    target = rng.integers(0, m["num_item"], n).astype(np.int32)
    sampled = np.array([[t] + sorted(rng.choice([x for x in range(m["num_item"]) if x != t], S, replace=False).tolist()) for t in target], np.int32)
    return seq, target, sampled


def _close(a, b, tol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def _flat(dl):
    return [dl["layer_emb"]] + [x for d in range(len(dl["layer_w"])) for x in (dl["layer_w"][d], dl["layer_b"][d])]


@pytest.mark.parametrize("shape,parallelism", [((300, 50, 3, 6, 16, 2), 1), ((500, 100, 2, 10, 24, 1), 3), ((120, 9, 4, 3, 8, 3), 4)])
def test_dr_gradients_match_oracle(orc, shape, parallelism):
    num_item, K, D, T, E, P = shape
    m, item_paths = _model(7, *shape)
    rng = np.random.default_rng(8)
    n, S = 37, 12
    seq, target, sampled = _batch(rng, m, n, S)
    e = new_engine()
    e.dr_load(**m)
    e.dr_load_item_paths(item_paths)
    tr = orc.DrTrainer(orc.DrModel(**{k: (v.copy() if isinstance(v, np.ndarray) else ([x.copy() for x in v] if isinstance(v, list) else v)) for k, v in m.items()}), 1e-3)
    g_l, loss_l = tr.layer_grad(seq, target, item_paths, P, parallelism)
    g_r, loss_r = tr.rerank_grad(seq, sampled)
    loss, rloss = e.dr_train_step(seq, target, 1e-3, 1, sampled=sampled, parallelism=parallelism, apply=False)
    assert _close(loss, loss_l, 1e-13) and abs(rloss - loss_r) < 1e-13
    got = e.dr_download(gradients=True)
    for a, b in zip(_flat(got), g_l):
        assert _close(a, b)
    assert _close(got["rr_emb"], g_r[0]) and _close(got["rr_w"], g_r[1]) and _close(got["rr_b"], g_r[2])
    assert _close(got["sm_w"], tr.sm_g[0]) and _close(got["sm_b"], tr.sm_g[1])
    untouched = np.setdiff1d(np.arange(num_item), np.unique(seq))
    assert (got["rr_emb"][untouched] == 0).all() and (got["layer_emb"][untouched] == 0).all()
    # parameters untouched by apply = False; a second call does not pile layer gradients up (zeroGradParameters) but does pile the
    # softmax-parameter gradients up (ParameterOptimizer.scala:67-88 never zeroes them)
    w0 = e.dr_download()
    assert (w0["layer_w"][0] == m["layer_w"][0]).all() and (w0["rr_w"] == m["rr_w"]).all() and (w0["sm_w"] == m["sm_w"]).all()
    e.dr_train_step(seq, target, 1e-3, 1, sampled=sampled, parallelism=parallelism, apply=False)
    again = e.dr_download(gradients=True)
    assert _close(again["layer_w"][D - 1], g_l[1 + 2 * (D - 1)]) and _close(again["sm_w"], 2 * tr.sm_g[0])
    e.close()


def test_dr_training_steps_match_oracle(orc):
    """four iterations of LocalOptimizer.optimize's loop: layer Adam, SampledSoftmaxLoss's own Adam, rerank Adam; the rerank model
    stops training after iteration 3 (reRankStoppingEpoch); then retrieval runs on the trained tables."""
    shape = (400, 40, 3, 5, 16, 2)
    num_item, K, D, T, E, P = shape
    m, item_paths = _model(11, *shape)
    rng = np.random.default_rng(12)
    lr, S = 3e-3, 10
    e = new_engine()
    e.dr_load(**m)
    e.dr_load_item_paths(item_paths)
    om = orc.DrModel(**{k: (v.copy() if isinstance(v, np.ndarray) else ([x.copy() for x in v] if isinstance(v, list) else v)) for k, v in m.items()})
    tr = orc.DrTrainer(om, lr)
    for t in range(1, 5):
        seq, target, sampled = _batch(rng, m, 48, S)
        rt = t if t <= 3 else 0
        want_l, want_r = tr.step(seq, target, item_paths, P, sampled, t, rt)
        got_l, got_r = e.dr_train_step(seq, target, lr, t, rerank_step_t=rt, sampled=sampled)
        assert _close(got_l, want_l, 1e-12), t
        assert (np.isnan(got_r) and np.isnan(want_r)) or abs(got_r - want_r) < 1e-12, t
    w = e.dr_download()
    # Adam divides by sqrt(r) + eps: a gradient component of 1e-13 relative error next to r ~ g^2 moves the step by the same relative amount
    for a, b in zip(_flat(w), [om.layer_emb] + [x for d in range(D) for x in (om.layer_w[d], om.layer_b[d])]):
        assert _close(a, b, 1e-9)
    for k in ("rr_emb", "rr_w", "rr_b", "sm_w", "sm_b"):
        assert _close(w[k], getattr(om, k), 1e-9), k
    assert np.abs(w["layer_w"][0] - m["layer_w"][0]).max() > 1e-3          # the weights did move
    # the beam search reads the refreshed in-major copies: same paths as the oracle on ITS trained tables
    om2 = orc.DrModel(num_item, K, D, T, E, w["layer_emb"], w["layer_w"], w["layer_b"], w["rr_emb"], w["rr_w"], w["rr_b"], w["sm_w"], w["sm_b"])
    q = rng.integers(0, num_item, (5, T)).astype(np.int32)
    paths, probs, counts = e.dr_beam_search(q, 8)
    for u in range(5):
        op, opr = om2.beam_search(q[u], 8)
        assert counts[u] == len(op) and (paths[u, :counts[u]] == op).all() and (probs[u, :counts[u]].view(np.uint64) == opr.view(np.uint64)).all()
    e.close()


def test_dr_device_sampler_properties():
    """SampledSoftmaxLoss.uniformSampler (:156-178) on the device: positive first, numSampled distinct negatives != positive in
    ascending order, roughly uniform; a training step with it runs and the rerank loss starts near log(S + 1)."""
    shape = (200, 20, 2, 4, 8, 1)
    m, item_paths = _model(13, *shape)
    m["sm_w"] *= 0.0
    e = new_engine()
    e.dr_load(**m)
    e.dr_load_item_paths(item_paths)
    rng = np.random.default_rng(14)
    seq, target, _ = _batch(rng, m, 64, 3)
    S = 20
    loss, rloss = e.dr_train_step(seq, target, 1e-3, 1, num_sampled=S, seed=99)
    assert abs(rloss - np.log(S + 1)) < 1e-9
    g = e.dr_download(gradients=True)["sm_b"]                              # never zeroed: (softmax - onehot) / n summed per item
    assert abs(g.sum()) < 1e-12 and (g != 0).sum() > 64
    with pytest.raises(Exception):
        e.dr_train_step(seq, target, 1e-3, 1, num_sampled=200, seed=1)      # numSampled < numClasses
    bad = target.copy()
    bad[3] = 200
    with pytest.raises(Exception):
        e.dr_train_step(seq, bad, 1e-3, 1, num_sampled=5, seed=1)
    e.close()


def test_dr_local_optimizer_mirror_learns():
    """dismember_b200.dr.LocalOptimizer (the Scala class's mini-batch loop over dmg_dr_train_step): on a learnable toy set (the target's
    paths are a function of the first history item) both losses fall; the rerank model stops after reRankEpoch."""
    from dismember_b200.dr import DeepRetrieval, LocalOptimizer
    shape = (60, 6, 2, 3, 8, 1)
    num_item, K, D, T, E, P = shape
    m, _ = _model(21, *shape)
    for k in ("layer_emb", "rr_emb", "rr_w", "sm_w"):
        m[k] = m[k] * 0.1
    m["layer_w"] = [w * 0.1 for w in m["layer_w"]]
    item_paths = np.stack([np.arange(num_item) % K, (np.arange(num_item) // K) % K], 1).reshape(num_item, 1, 2).astype(np.int32)
    rng = np.random.default_rng(22)
    seqs = rng.integers(0, num_item, (512, T)).astype(np.int32)
    targets = seqs[:, 0].copy()
    e = new_engine()
    dr = DeepRetrieval(engine=e).set_model(**m)
    opt = LocalOptimizer(dr, item_paths, learning_rate=2e-2, num_sampled=8, batch_size=128, re_rank_epoch=6, seed=5)
    hist = [opt.train_epoch(ep, seqs, targets) for ep in range(1, 9)]
    first, last = hist[0][0], hist[5][-1]
    assert last[0].sum() < 0.7 * first[0].sum() and last[1] < 0.9 * first[1]
    assert all(np.isnan(r) for _, r in hist[7]) and opt.layer_t == 32 and opt.rerank_t == 24
    assert np.abs(dr.get_parameters()["layer_w"][0] - m["layer_w"][0]).max() > 1e-2
    e.close()
