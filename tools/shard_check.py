#!/usr/bin/env python3
"""Table-sharded TDM retrieval across the GPUs of one box, checked against the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/shard_check.py --items 1000000 --batch 256

Every rank holds 1/world of the node table (csrc/shard.cu), brings its own `batch` users, and prints one JSON line:
parity of its results with the oracle run on the full table, rows it scored for other ranks, users/s."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=100_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--beam", type=int, default=200)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--oracle-users", type=int, default=64)
    ap.add_argument("--check", default="oracle", choices=["oracle", "engine"],
                    help="oracle: CPU oracle on the full table (needs the table on the host); engine: the unsharded strict "
                         "CUDA engine on this rank's GPU (for catalogues too large to ship to the host)")
    ap.add_argument("--jtm-items", type=int, default=0, help="also check dmg_shard_jtm_item_weights on this many items per rank")
    ap.add_argument("--jtm-gap", type=int, default=4)
    ap.add_argument("--dr-items", type=int, default=0, help="also check dmg_shard_dr_retrieve on a synthetic model with this many items")
    ap.add_argument("--dr-k", type=int, default=100)
    ap.add_argument("--train-targets", type=int, default=0, help="also check dmg_dp_train_step (rows from this many targets per rank)")
    ap.add_argument("--dr-batch", type=int, default=64)
    ap.add_argument("--shard-train-targets", type=int, default=0,
                    help="also check dmg_shard_train_step (training on the sharded table; rows from this many targets per rank)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from dismember_b200 import shard, synth
    from oracle import oracle as orc
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)       # control plane only: the data path is NCCL inside the library
    torch.cuda.set_device(local)
    T = 10
    tf = synth.tdm_tree(a.items, seed=1)
    rows = (1 << (tf.max_level + 1)) - 1
    eng = shard.make_sharded_engine(local)
    box = [eng.shard_unique_id() if rank == 0 else None]           # a second communicator for the data-parallel replica check
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    uid_for_dp = box[0]
    eng.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    eng.shard_init_din_weights(rows, a.dim, T, seed=2)
    seqs = synth.queries(a.batch, T, a.items, seed=100 + rank)
    items, logits, counts = eng.shard_tdm_retrieve(seqs, a.beam, a.topk)           # warm-up + the checked result
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        eng.shard_tdm_retrieve(seqs, a.beam, a.topk)
    dist.barrier()
    dt = time.perf_counter() - t0
    local_rows, global_rows, exchanged = eng.shard_info()
    # oracle on the full table (same counter-based values: an unsharded engine generates them)
    n = min(a.oracle_users, a.batch)
    from dismember_b200 import Engine
    full = Engine(local)
    full.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    full.init_din_weights(np.float32, rows, a.dim, T, seed=2)
    if a.check == "oracle":
        params = full.download_din_weights()
        full.close()
        orc.build()
        tree = orc.Tree.from_treefile(tf)
        model = orc.TdmModel(params, rows, a.dim, T)
        oi, ol, oc = model.retrieve_batch(tree, seqs[:n], a.beam, a.topk, n_threads=max(1, (os.cpu_count() or 2) // world))
    else:
        full.set_arithmetic("strict")
        oi, ol, oc = full.tdm_retrieve(seqs[:n], a.beam, a.topk)
        full.close()
    # JTM item weights over the sharded table (config 4): this rank's slice of the items vs the unsharded engine
    jtm = None
    if a.jtm_items > 0:
        rng = np.random.Generator(np.random.PCG64(50 + rank))
        n_it = a.jtm_items
        n_samples = rng.integers(0, 9, n_it)
        off = np.zeros(n_it + 1, np.int64)
        off[1:] = np.cumsum(n_samples)
        sseq = synth.queries(int(off[-1]), T, a.items, seed=70 + rank)
        old_level, level = 6, 6 + a.jtm_gap
        par = rng.integers((1 << old_level) - 1, (2 << old_level) - 1, n_it).astype(np.int32)
        t0 = time.perf_counter()
        got = eng.shard_jtm_item_weights(off, sseq, par, old_level, level, hierarchical=True, min_level=0)
        jdt = time.perf_counter() - t0
        full = Engine(local)
        full.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        full.init_din_weights(np.float32, rows, a.dim, T, seed=2)
        want = full.jtm_item_weights(off, sseq, par, old_level, level, hierarchical=True, min_level=0)
        full.close()
        jtm = {"items": n_it, "samples": int(off[-1]), "gap": a.jtm_gap, "scorer_rows": int(off[-1]) * ((2 << a.jtm_gap) - 2),
               "weights_bit_identical": bool((got.view(np.uint32) == want.view(np.uint32)).all()), "seconds": jdt}
    # Deep Retrieval with sharded item tables (config 5): this rank's users vs the unsharded engine on the whole tables
    dr = None
    if a.dr_items > 0:
        from dismember_b200.dr import build_path_csr
        n_item, K, D, Ed = a.dr_items, a.dr_k, 3, 32
        rng = np.random.Generator(np.random.PCG64(8))                       # same tables on every rank
        layer_emb = rng.normal(0, 0.05, (n_item + (D - 1) * K, Ed))
        layer_w = [rng.normal(0, 0.05, (K, (T + d) * Ed)) for d in range(D)]
        layer_b = [np.zeros(K) for _ in range(D)]
        rr_emb = rng.normal(0, 0.05, (n_item, Ed)); rr_w = rng.normal(0, 0.05, (Ed, T * Ed)); rr_b = np.zeros(Ed)
        sm_w = rng.normal(0, 0.05, (n_item, Ed)); sm_b = rng.normal(0, 0.01, n_item)
        off, flat = build_path_csr(np.arange(n_item), rng.integers(0, K, (n_item, 2, D)), K)
        args = (n_item, K, D, T, Ed, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
        dseq = np.random.Generator(np.random.PCG64(90 + rank)).integers(-1, n_item, (a.dr_batch, T)).astype(np.int32)
        eng.shard_dr_load(*args)
        eng.dr_load_paths(off, flat)
        eng.shard_dr_retrieve(dseq, 200, 10)
        dist.barrier()
        t0 = time.perf_counter()
        si, ss, sc = eng.shard_dr_retrieve(dseq, 200, 10)
        dist.barrier()
        ddt = time.perf_counter() - t0
        full = Engine(local)
        full.dr_load(*args)
        full.dr_load_paths(off, flat)
        ri, rs, rc = full.dr_retrieve(dseq, 200, 10)
        full.close()
        dr = {"items": n_item, "K": K, "D": D, "E": Ed, "batch_per_rank": a.dr_batch, "beam": 200,
              "ids_identical": bool((si == ri).all() and (sc == rc).all()),
              "scores_bit_identical": bool((ss.view(np.uint64) == rs.view(np.uint64)).all()),
              "users_per_s_whole_job": world * a.dr_batch / ddt}
    # data-parallel training step over replicas (dmg_dp_train_step) vs ONE engine training on the concatenated batch
    dp = None
    if a.train_targets > 0:
        n_it = min(a.items, 50_000)
        ttf = synth.tdm_tree(n_it, seed=5)
        trows = (1 << (ttf.max_level + 1)) - 1
        rep = Engine(local)
        rep.shard_init(world, rank, uid_for_dp)
        rep.load_tree_tdm(ttf.max_level, ttf.codes, ttf.node_ids, ttf.is_leaf, ttf.leaf_ids, ttf.leaf_codes)
        rep.init_din_weights(np.float32, trows, 16, T, seed=3)
        rng = np.random.Generator(np.random.PCG64(200 + rank))
        tg = rng.integers(1, n_it + 1, a.train_targets).astype(np.int32)
        tsq = synth.queries(a.train_targets, T, n_it, seed=300 + rank)
        neg = np.array([0] + [min(2 ** l - 1, 15) for l in range(1, ttf.max_level + 1)], np.int32)
        node, sq, lab = rep.tdm_sample_expand(tg, tsq, neg, 1, seed=400 + rank)
        mask = np.flatnonzero((sq == -1).ravel()).astype(np.int32)
        losses = [float(rep.dp_train_step(node, sq, mask, lab, 1e-2, t)) for t in (1, 2)]
        w_dp = rep.download_din_weights()
        rep.close()
        parts = [None] * world
        dist.all_gather_object(parts, (node, sq, lab))
        one = Engine(local)
        one.load_tree_tdm(ttf.max_level, ttf.codes, ttf.node_ids, ttf.is_leaf, ttf.leaf_ids, ttf.leaf_codes)
        one.init_din_weights(np.float32, trows, 16, T, seed=3)
        an, asq, al = (np.concatenate([p[i] for p in parts]) for i in range(3))
        am = np.flatnonzero((asq == -1).ravel()).astype(np.int32)
        for t in (1, 2):
            one.train_step(an, asq, am, al, 1e-2, t)
        w_one = one.download_din_weights()
        one.close()
        moved = np.abs(w_one - np.float32(0)).max()
        dp = {"rows_per_rank": int(len(node)), "steps": 2, "loss": losses,
              "max_abs_weight_diff_vs_single_engine": float(np.abs(w_dp - w_one).max()),
              "max_abs_weight": float(moved)}
    # training on the SHARDED table (dmg_shard_train_step) vs ONE unsharded engine training on the concatenated batch
    sht = None
    if a.shard_train_targets > 0:
        rng = np.random.Generator(np.random.PCG64(500 + rank))
        tg = rng.integers(1, a.items + 1, a.shard_train_targets + 3 * rank).astype(np.int32)       # unequal row counts per rank
        tsq = synth.queries(len(tg), T, a.items, seed=600 + rank)
        neg = np.array([0] + [min(2 ** l - 1, 7) for l in range(1, tf.max_level + 1)], np.int32)
        sampler = Engine(local)                                          # the sampler needs an unsharded tree + seq_len only
        sampler.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        sampler.init_din_weights(np.float32, rows, a.dim, T, seed=1)
        node, sq, lab = sampler.tdm_sample_expand(tg, tsq, neg, 1, seed=700 + rank)
        sampler.close()
        mask = np.flatnonzero((sq == -1).ravel()).astype(np.int32)
        ex0 = eng.shard_info()[2]
        t0 = time.perf_counter()
        losses = [float(eng.shard_train_step(node, sq, mask, lab, 1e-2, t)) for t in (1, 2)]
        tdt = time.perf_counter() - t0
        w_loc = eng.download_din_weights()
        parts = [None] * world
        dist.all_gather_object(parts, (node, sq, lab))
        one = Engine(local)
        one.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        one.init_din_weights(np.float32, rows, a.dim, T, seed=2)
        an, asq, al = (np.concatenate([p[i] for p in parts]) for i in range(3))
        am = np.flatnonzero((asq == -1).ravel()).astype(np.int32)
        one_losses = [float(one.train_step(an, asq, am, al, 1e-2, t)) for t in (1, 2)]
        w_one = one.download_din_weights()
        one.close()
        n_loc = eng.shard_info()[0]
        gr = shard.global_row(np.arange(n_loc), world, rank)
        emb_loc, emb_one = w_loc[:n_loc * a.dim].reshape(n_loc, a.dim), w_one[:rows * a.dim].reshape(rows, a.dim)
        w0 = Engine(local)
        w0.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        w0.init_din_weights(np.float32, rows, a.dim, T, seed=2)
        moved = float(np.abs(w_one - w0.download_din_weights()).max())
        w0.close()
        sht = {"rows_this_rank": int(len(node)), "rows_global": int(len(an)), "steps": 2, "loss": losses, "loss_single_engine": one_losses,
               "max_abs_emb_diff_vs_single_engine": float(np.abs(emb_loc - emb_one[gr]).max()),
               "max_abs_dense_diff_vs_single_engine": float(np.abs(w_loc[n_loc * a.dim:] - w_one[rows * a.dim:]).max()),
               "max_abs_weight_change": moved, "rows_fetched_for_other_ranks": int(eng.shard_info()[2] - ex0), "seconds": tdt,
               "all_reduced_scalars": int(3 * a.dim * a.dim + 2 * a.dim + 1 + ((world - 1) * a.dim))}
        # the trained shards keep serving: retrieval on them == the single engine that took the same steps is checked by the caller's
        # retrieval legs on fresh weights; here only the optimiser state is compared
    line = {"rank": rank, "world": world, "shard_train_step": sht, "jtm_item_weights": jtm, "deep_retrieval": dr, "dp_train_step": dp, "items": a.items, "levels": tf.max_level, "batch_per_rank": a.batch, "beam": a.beam,
            "table_rows_global": global_rows, "table_rows_local": local_rows,
            "rows_scored_for_other_ranks": exchanged, "users_checked": n, "checked_against": a.check,
            "ids_identical": bool((items[:n] == oi).all() and (counts[:n] == oc).all()),
            "logits_bit_identical": bool((logits[:n].view(np.uint32) == ol.view(np.uint32)).all()),
            "users_per_s_whole_job": world * a.batch * a.steps / dt}
    print(json.dumps(line), flush=True)
    if a.out:
        json.dump(line, open(a.out, "w"))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
