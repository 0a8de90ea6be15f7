"""Node-table sharding (csrc/shard.cu): the level-synchronous request/score exchange must return the bits of the
unsharded search.  world = 1 runs everywhere (all owner regions are local, no NCCL); world = 2 needs two GPUs and
is skipped on the one-GPU box (tools/shard_check.py is the same check under torchrun, see profiles/)."""
import os
import sys

import numpy as np
import pytest

from conftest import new_engine

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_fixture_matches_oracle(jtm_fix, queries, golden_out):
    f = jtm_fix
    e = new_engine()
    e.shard_init(1, 0)
    e.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    e.shard_load_din_weights(f["params"], 8191, 16, 10)
    items, logits, counts = e.shard_tdm_retrieve(queries["seqs"][:64], 20, 10)
    assert (items == golden_out["tdm_items_b20"][:64]).all()
    assert (logits.view(np.uint32) == golden_out["tdm_logits_b20"][:64].view(np.uint32)).all()
    assert e.shard_info()[:2] == (8191, 8191)
    e.close()


def test_world1_synthetic_matches_strict_engine_and_oracle(orc):
    from dismember_b200 import synth
    n_items, E, T, beam, topk, B = 5000, 64, 10, 200, 10, 40
    tf = synth.tdm_tree(n_items, seed=1)
    rows = (1 << (tf.max_level + 1)) - 1
    seqs = synth.queries(B, T, n_items, seed=4)
    ref = new_engine()
    ref.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    ref.init_din_weights(np.float32, rows, E, T, seed=2)
    ref.set_arithmetic("strict")
    ri, rl, rc = ref.tdm_retrieve(seqs, beam, topk)
    params = ref.download_din_weights()
    ref.close()
    e = new_engine()
    e.shard_init(1, 0)
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.shard_init_din_weights(rows, E, T, seed=2)              # same counter-based values as init_din_weights
    si, sl, sc = e.shard_tdm_retrieve(seqs, beam, topk)
    e.close()
    assert (si == ri).all() and (sc == rc).all()
    assert (sl.view(np.uint32) == rl.view(np.uint32)).all()
    tree = orc.Tree.from_treefile(tf)
    model = orc.TdmModel(params, rows, E, T)
    oi, ol, oc = model.retrieve_batch(tree, seqs, beam, topk, n_threads=os.cpu_count() or 1)
    assert (si == oi).all() and (sl.view(np.uint32) == ol.view(np.uint32)).all() and (sc == oc).all()


def test_unsharded_entry_points_refuse_a_sharded_table(jtm_fix, queries):
    from dismember_b200._capi import DmgError
    f = jtm_fix
    e = new_engine()
    e.shard_init(1, 0)
    e.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    with pytest.raises(DmgError):
        e.shard_load_din_weights(f["params"], 8191 * 2 + 1, 16, 10)      # table must match the tree
    e.close()


def test_world1_jtm_item_weights_match_the_unsharded_engine(jtm_fix, queries):
    """dmg_shard_jtm_item_weights (scorer rows routed like retrieval candidates) vs dmg_jtm_item_weights, bit for bit;
    hierarchical re-coding, items without samples, several chunks."""
    f = jtm_fix
    L = int(f["max_level"])
    rng = np.random.default_rng(3)
    n_it = 300
    items = rng.choice(f["leaf_ids"], n_it, replace=False)
    counts = rng.integers(0, 5, n_it)
    counts[:5] = 0
    off = np.zeros(n_it + 1, np.int64)
    off[1:] = np.cumsum(counts)
    seqs = rng.choice(f["leaf_ids"], (int(off[-1]), 10)).astype(np.int32)
    seqs[:, :2] = 0
    old_level, level = 3, 6
    par = rng.integers((1 << old_level) - 1, (2 << old_level) - 1, n_it).astype(np.int32)
    ref = new_engine()
    ref.load_tree_tdm(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    ref.load_din_weights(f["params"], 8191, 16, 10)
    e = new_engine()
    e.shard_init(1, 0)
    e.load_tree_tdm(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    e.shard_load_din_weights(f["params"], 8191, 16, 10)
    for hier in (False, True):
        want = ref.jtm_item_weights(off, seqs, par, old_level, level, hierarchical=hier, min_level=0)
        got = e.shard_jtm_item_weights(off, seqs, par, old_level, level, hierarchical=hier, min_level=0)
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
        assert (got[:5] == np.float32(-1e6)).all()
    ref.close()
    e.close()


def test_world1_deep_retrieval_matches_the_unsharded_engine(dr_fix, queries):
    """dmg_shard_dr_retrieve (history tiles, owner-scored rerank candidates) vs dmg_dr_retrieve on the DR fixture, bit for bit."""
    from dismember_b200.dr import build_path_csr
    f = dr_fix
    D = int(f["D"])
    args = (int(f["num_item"]), int(f["K"]), D, int(f["T"]), int(f["E"]), f["layer_emb"],
            [f[f"layer_w{d}"] for d in range(D)], [f[f"layer_b{d}"] for d in range(D)],
            f["rr_emb"], f["rr_w"], f["rr_b"], f["sm_w"], f["sm_b"])
    off, flat = build_path_csr(f["map_ids"], f["map_paths"], int(f["K"]))
    item_id = {int(a): int(b) for a, b in zip(f["map_items"], f["map_ids"])}
    seqs = np.array([[item_id.get(int(x), -1) for x in s] for s in queries["seqs"][:40]], np.int32)
    seqs[1] = -1
    ref = new_engine()
    ref.dr_load(*args)
    ref.dr_load_paths(off, flat)
    e = new_engine()
    e.shard_init(1, 0)
    e.shard_dr_load(*args)
    e.dr_load_paths(off, flat)
    for beam, topk in [(50, 10), (300, 20), (7, 3)]:
        ri, rs, rc = ref.dr_retrieve(seqs, beam, topk)
        si, ss, sc = e.shard_dr_retrieve(seqs, beam, topk)
        assert (sc == rc).all() and (si == ri).all()
        assert (ss.view(np.uint64) == rs.view(np.uint64)).all()
    ref.close()
    e.close()


def _two_gpu_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.argv = ["shard_check.py", "--items", "5000", "--batch", "24", "--train-targets", "40", "--jtm-items", "200", "--dr-items", "3000",
                "--dr-k", "12", "--dr-batch", "16", "--shard-train-targets", "40", "--out", os.path.join(out_dir, f"r{rank}.json")]
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import runpy
    runpy.run_path(os.path.join(ROOT, "tools", "shard_check.py"), run_name="__main__")


def test_world2_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run tools/shard_check.py under torchrun with gpurun --gpus 2)")
    import json
    import torch.multiprocessing as mp
    mp.spawn(_two_gpu_worker, args=(2, 32500 + os.getpid() % 2000, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = json.load(open(tmp_path / f"r{r}.json"))
        assert d["ids_identical"] and d["logits_bit_identical"] and d["rows_scored_for_other_ranks"] > 0
        assert d["jtm_item_weights"]["weights_bit_identical"]
        assert d["deep_retrieval"]["ids_identical"] and d["deep_retrieval"]["scores_bit_identical"]
        # dmg_shard_train_step: owner-applied embedding gradients + all-reduce of the dense scalars == one engine on the whole batch
        st = d["shard_train_step"]
        assert st["max_abs_emb_diff_vs_single_engine"] < 1e-5 and st["max_abs_dense_diff_vs_single_engine"] < 1e-5
        assert st["max_abs_weight_change"] > 1e-3 and st["rows_fetched_for_other_ranks"] > 0
        assert all(abs(x - y) < 1e-5 for x, y in zip(st["loss"], st["loss_single_engine"]))
        # dmg_dp_train_step (LocalOptimizer.syncGradients with GPUs in the place of threads): two steps on two ranks == one engine on
        # the concatenated batch up to the fp32 summation order
        assert d["dp_train_step"]["max_abs_weight_diff_vs_single_engine"] < 1e-5 * max(1.0, d["dp_train_step"]["max_abs_weight"])


def test_dr_synthetic_model_unsharded_sharded_oracle(orc):
    """dmg_dr_init_synthetic: the device-generated model retrieves the same items as the oracle fed with the downloaded tables and
    the host mirror of the path assignment; a world-1 sharded handle (its own item range = everything) gives the same bits."""
    from dismember_b200 import synth
    num_item, K, D, T, E, J = 3000, 12, 3, 5, 16, 2
    e = new_engine()
    e.dr_init_synthetic(num_item, K, D, T, E, J, seed=77)
    w = e.dr_download()
    off, items = synth.dr_synthetic_path_csr(num_item, K, D, J, 77)
    assert off[-1] == len(items) and 0 < len(items) <= K ** D and (np.diff(off) <= 1).all()
    om = orc.DrModel(num_item, K, D, T, E, w["layer_emb"], w["layer_w"], w["layer_b"], w["rr_emb"], w["rr_w"], w["rr_b"], w["sm_w"], w["sm_b"])
    assert abs(w["sm_w"].std() - 0.05) < 0.005 and (w["layer_b"][0] == 0).all()
    rng = np.random.default_rng(3)
    seq = rng.integers(-1, num_item, (9, T)).astype(np.int32)
    gi, gs, gc = e.dr_retrieve(seq, 30, 10)
    for u in range(len(seq)):
        oi, os_, _ = om.recommend(seq[u], 10, 30, off, items)
        assert gc[u] == len(oi) and (gi[u, :gc[u]] == oi).all() and (gs[u, :gc[u]].view(np.uint64) == os_.view(np.uint64)).all()
    assert gc.sum() > 0
    s = new_engine()
    s.shard_init(1, 0)
    s.dr_init_synthetic(num_item, K, D, T, E, J, seed=77)
    si, ss, sc = s.shard_dr_retrieve(seq, 30, 10)
    assert (sc == gc).all() and (si == gi).all() and (ss.view(np.uint64) == gs.view(np.uint64)).all()
    s.close()
    e.close()


def test_shard_train_step_world1_is_train_step(jtm_fix):
    """dmg_shard_train_step on one rank (every row local: no exchange, nothing to all-reduce) == dmg_train_step; bad codes raise."""
    from dismember_b200._capi import DmgError
    f = jtm_fix
    params = f["params"]
    rng = np.random.default_rng(26)
    a, b = new_engine(), new_engine()
    a.shard_init(1, 0)
    a.load_tree_tdm(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    a.shard_load_din_weights(params, 8191, 16, 10)
    b.load_din_weights(params, 8191, 16, 10)
    for t in (1, 2, 3):
        n = 300
        node = rng.integers(0, 8191, n).astype(np.int32)
        seq = rng.integers(0, 8191, (n, 10)).astype(np.int32)
        seq[rng.random((n, 10)) < 0.3] = -1
        mask = np.flatnonzero((seq == -1).ravel()).astype(np.int32)
        labels = (rng.random(n) < 0.2).astype(np.float32)
        la = a.shard_train_step(node, seq, mask, labels, 1e-2, t)
        lb = b.train_step(node, seq, mask, labels, 1e-2, t)
        assert abs(float(la) - float(lb)) <= 1e-5 * max(1.0, abs(float(lb)))
    wa, wb = a.download_din_weights(), b.download_din_weights()
    assert np.abs(wa - wb).max() <= 2e-4 and np.abs(wa - params).max() > 1e-3
    node[3] = 8191
    with pytest.raises(DmgError):
        a.shard_train_step(node, seq, mask, labels, 1e-2, 4)
    a.close()
    b.close()
