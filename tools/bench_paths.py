#!/usr/bin/env python3
"""Secondary bench lines: every SURVEY 8(a) kernel other than the headline TDM search, each with its
algorithmic figure against the measured B200 peak and the CPU oracle timed beside it.

    python tools/bench_paths.py [--quick] > profiles/rN_paths.jsonl

One JSON line per path.  Timed through the host-buffer C ABI (the call a Scala host would make), wall
clock around synchronous calls, after warm-up.  bench.py stays the headline; these lines explain the
other rows of DESIGN.md section 4."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def otm_deepfm_section(a, E, T, threads):
    """3c. OTM retrieval with DeepModel[Double] = DeepFM: level-synchronous path of otm_deepfm.cu"""
    from dismember_b200 import Engine, synth
    from oracle import oracle as orc
    n_items = 100_000 if a.quick else 1_000_000
    items, leaf_ids, leaf_level = synth.otm_mapping(n_items, seed=42)
    rows_tab = (1 << (leaf_level + 1)) - 1
    F = T + 1
    rng = np.random.Generator(np.random.PCG64(13))
    dparams = np.concatenate([rng.normal(0, 0.05, rows_tab * E), rng.normal(0, 0.05, F * F * E), np.zeros(F),
                              rng.normal(0, 0.3, F), [0.0]])
    eng = Engine(0)
    eng.load_tree_complete(leaf_level, items, leaf_ids)
    eng.load_deepfm_weights(dparams, rows_tab, E, T)
    B = 256
    lseq = leaf_ids[rng.integers(0, n_items, (B, T))].astype(np.int32)
    lseq[:, :3] = -1
    dt = timeit(lambda: eng.otm_retrieve(lseq, 200, 10), warm=1, reps=3)
    rows_u = 256 + 400 * (leaf_level - 8)
    flop_u = rows_u * 2 * (F + 1) * F * E
    emit(path="otm_retrieve with the DeepFM scorer (fp64, level-synchronous)", items=n_items, levels=leaf_level, batch=B, ms=dt * 1e3,
         users_per_s=B / dt, roofline={"bound": "fp64 FMA pipe", "algorithmic_flop_per_user": flop_u,
                                       "achieved_tflops": flop_u * B / dt / 1e12})
    leaf_item = np.full(1 << leaf_level, -1, np.int32)
    leaf_item[leaf_ids - ((1 << leaf_level) - 1)] = items
    om = orc.OtmModel(dparams, rows_tab, E, T, deepfm=True)
    t0 = time.perf_counter()
    oi, osc, oc = om.retrieve_batch(lseq[:2 * threads], leaf_level, 200, 10, leaf_item, n_threads=threads)
    cdt = time.perf_counter() - t0
    gi, gs, gc = eng.otm_retrieve(lseq[:2 * threads], 200, 10)
    emit(path="otm_retrieve DeepFM cpu_baseline", kind="port", cores=threads, users_per_s=2 * threads / cdt,
         parity={"ids_identical": bool((gi == oi).all()), "scores_bit_identical": bool((gs.view(np.uint64) == osc.view(np.uint64)).all())})
    eng.close()


def dr_section(a, T):
    """4. Deep Retrieval beam search + rerank, fp64 (a23)"""
    from dismember_b200 import Engine
    from oracle import oracle as orc
    n_item, K, D, J = (20_000, 100, 3, 2) if a.quick else (200_000, 1000, 3, 2)
    Ed = 16 if a.quick else 64
    rng = np.random.Generator(np.random.PCG64(8))
    layer_emb = rng.normal(0, 0.05, (n_item + (D - 1) * K, Ed))
    layer_w = [rng.normal(0, 0.05, (K, (T + d) * Ed)) for d in range(D)]
    layer_b = [np.zeros(K) for _ in range(D)]
    rr_emb = rng.normal(0, 0.05, (n_item, Ed)); rr_w = rng.normal(0, 0.05, (Ed, T * Ed)); rr_b = np.zeros(Ed)
    sm_w = rng.normal(0, 0.05, (n_item, Ed)); sm_b = np.zeros(n_item)
    from dismember_b200.dr import build_path_csr
    paths = rng.integers(0, K, (n_item, J, D))
    off, flat = build_path_csr(np.arange(n_item), paths, K)
    eng = Engine(0)
    eng.dr_load(n_item, K, D, T, Ed, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
    eng.dr_load_paths(off, flat)
    B = 64 if a.quick else 1024                                # enough users for every resident CTA of the persistent kernel
    dseq = rng.integers(0, n_item, (B, T)).astype(np.int32)
    dt = timeit(lambda: eng.dr_retrieve(dseq, 200, 10), warm=1, reps=3)
    flop_u = 2 * K * T * Ed + sum(200 * 2 * K * (T + d) * Ed for d in range(1, D))
    emit(path="dr_retrieve (Deep Retrieval beam search + rerank, fp64)", items=n_item, K=K, D=D, beam=200, batch=B, ms=dt * 1e3,
         users_per_s=B / dt, roofline={"bound": "fp64 FMA pipe / top-k", "naive_flop_per_user": flop_u,
                                       "achieved_tflops_naive": flop_u * B / dt / 1e12})
    dm = orc.DrModel(n_item, K, D, T, Ed, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
    nchk = 4
    t0 = time.perf_counter()
    ref = [dm.recommend(dseq[i], 10, 200, off, flat) for i in range(nchk)]
    cdt = time.perf_counter() - t0
    gi, gs, gc = eng.dr_retrieve(dseq[:nchk], 200, 10)
    ok = all((gi[i, :gc[i]] == ref[i][0]).all() and len(ref[i][0]) == gc[i] for i in range(nchk))
    emit(path="dr_retrieve cpu_baseline", kind="port", cores=1, users_per_s=nchk / cdt, parity={"ids_identical": bool(ok)})
    eng.close()


def dr_train_section(a, T):
    """5. Deep Retrieval training step (f3): layer-model cross entropy + rerank sampled softmax, fp64, dense Adam"""
    from dismember_b200 import Engine
    from oracle import oracle as orc
    hbm, _ = peaks()
    n_item, K, D, Ed, P, S, n = (20_000, 100, 3, 16, 2, 20, 256) if a.quick else (1_000_000, 1000, 3, 64, 2, 100, 2048)
    eng = Engine(0)
    eng.dr_init_synthetic(n_item, K, D, T, Ed, J=P, seed=8)
    rng = np.random.Generator(np.random.PCG64(9))
    item_paths = rng.integers(0, K, (n_item, P, D)).astype(np.int32)
    eng.dr_load_item_paths(item_paths)
    seq = rng.integers(-1, n_item, (n, T)).astype(np.int32)
    tg = rng.integers(0, n_item, n).astype(np.int32)
    step = [0]

    def one():
        step[0] += 1
        eng.dr_train_step(seq, tg, 1e-3, step[0], num_sampled=S, seed=step[0])
    dt = timeit(one, warm=2, reps=5)
    R = n * P
    gemm_flop = sum(3 * 2 * R * K * (T + d) * Ed for d in range(D)) + 3 * 2 * n * Ed * T * Ed
    n_par = (n_item + K * (D - 1)) * Ed + sum(K * (T + d) * Ed + K for d in range(D)) + n_item * Ed + Ed * T * Ed + Ed + n_item * Ed + n_item
    adam_bytes = 7 * 8 * n_par
    emit(path="dr_train_step (layer-model CE over n*P rows + rerank sampled softmax, fp64, dense Adam on every table)", items=n_item, K=K, D=D,
         E=Ed, samples=n, paths_per_item=P, num_sampled=S, ms_per_step=dt * 1e3, samples_per_s=n / dt,
         roofline={"bound": "fp64 FMA pipe (three GEMMs per Linear) + hbm (dense Adam)", "gemm_flop_per_step": gemm_flop,
                   "adam_bytes_per_step": adam_bytes, "gemm_tflops_if_all_time": gemm_flop / dt / 1e12,
                   "adam_gbs_if_all_time": adam_bytes / dt / 1e9, "hbm_peak": hbm})
    # CPU port on a small slice of the same shapes (single thread)
    if a.quick:
        w = eng.dr_download()
        om = orc.DrModel(n_item, K, D, T, Ed, w["layer_emb"], w["layer_w"], w["layer_b"], w["rr_emb"], w["rr_w"], w["rr_b"], w["sm_w"], w["sm_b"])
        tr = orc.DrTrainer(om, 1e-3)
        t0 = time.perf_counter()
        tr.layer_grad(seq[:32], tg[:32], item_paths, P)
        cdt = time.perf_counter() - t0
        emit(path="dr_train_step cpu_baseline (layer gradients only)", kind="port", cores=1, samples_per_s=32 / cdt)
    eng.close()


def kmeans_section(a):
    """6. k-means tree rebuild (f4): RecursiveCluster.run on item embeddings"""
    from dismember_b200 import Engine
    from oracle import oracle as orc
    n, E, iters = (20_000, 16, 2) if a.quick else (1_000_000, 64, 3)
    rng = np.random.Generator(np.random.PCG64(4))
    emb = rng.random((n, E))
    eng = Engine(0)
    eng.kmeans_tree(emb[:1000], 1, 1)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    codes = eng.kmeans_tree(emb, iters, 7)
    dt = time.perf_counter() - t0
    emit(path="kmeans_tree (RecursiveCluster.run: level-synchronous balanced 2-means + host argPartition)", items=n, E=E, runs=iters,
         seconds=dt, items_per_s=n / dt, launches=eng.launch_count - l0, distinct_codes=int(len(np.unique(codes))))
    m = 20_000 if not a.quick else 5_000
    t0 = time.perf_counter()
    oc = orc.kmeans_tree(emb[:m], iters, 7)
    cdt = time.perf_counter() - t0
    gc = eng.kmeans_tree(emb[:m], iters, 7)
    emit(path="kmeans_tree cpu_baseline", kind="port", cores=1, items=m, seconds=cdt, items_per_s=m / cdt,
         parity={"codes_identical": bool((gc == oc).all())})
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--train-items", type=int, default=10_000_000)
    ap.add_argument("--only", default=None, choices=[None, "otm_deepfm", "dr", "dr_train", "kmeans"], help="run one section only")
    a = ap.parse_args()
    from dismember_b200 import Engine, synth
    from oracle import oracle as orc
    orc.build()
    hbm, hbm_src = peaks()
    threads = os.cpu_count() or 1
    E, T = 64, 10
    if a.only == "otm_deepfm":
        return otm_deepfm_section(a, E, T, threads)
    if a.only == "dr":
        return dr_section(a, T)
    if a.only == "dr_train":
        return dr_train_section(a, T)
    if a.only == "kmeans":
        return kmeans_section(a)

    # ---- 1. training step: fused DIN fwd/bwd + BCE + scatter-add, dense Adam -------------------
    for n_items in ([100_000] if a.quick else [1_000_000, a.train_items]):
        tf = synth.tdm_tree(n_items, seed=1)
        L = tf.max_level
        rows_tab = (1 << (L + 1)) - 1
        eng = Engine(0)
        eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
        eng.init_din_weights(np.float32, rows_tab, E, T, seed=2)
        n_tg = 8                                               # ~8 k rows per step, the reference's batch (configs/tdm.conf)
        rng = np.random.Generator(np.random.PCG64(7))
        targets = rng.integers(1, n_items + 1, n_tg).astype(np.int32)
        seqs = synth.queries(n_tg, T, n_items, seed=9)
        layer_neg = np.array([0] + [min(2 ** l - 1, 63) for l in range(1, L + 1)], np.int32)     # 64 rows per level from level 6 on
        node, sq, lab = eng.tdm_sample_expand(targets, seqs, layer_neg, 1, seed=11)
        rows = len(node)
        mask = np.nonzero((sq.ravel() < 0))[0].astype(np.int32)
        step = [0]

        def one():
            step[0] += 1
            eng.train_step(node, sq, mask, lab, 1e-3, step[0])
        dt = timeit(one, warm=2, reps=5)
        n_par = rows_tab * E + 3 * E * E + 2 * E + 1
        dense = 7 * n_par * 4 + n_par * 4                     # Adam: read w,g,s,r + write w,s,r; grad zeroing
        sparse = 2 * rows * (1 + T) * E * 4
        emit(path="train_step (TDM sampler rows, DIN fwd/bwd, BCE, scatter-add, dense Adam)", items=n_items, levels=L,
             rows_per_step=rows, ms_per_step=dt * 1e3, rows_per_s=rows / dt,
             roofline={"bound": "hbm", "algorithmic_bytes_per_step": dense + sparse, "achieved": (dense + sparse) / dt / 1e9,
                       "peak": hbm, "unit": "GB/s", "frac": (dense + sparse) / dt / 1e9 / hbm, "peak_source": hbm_src},
             note="dense-Adam semantics of scalann Adam.scala:54-65: every parameter is touched every step")
        # the same step through the device-buffer entry point (no copies, no synchronisation inside): CUDA events around 20 calls
        import torch
        dev = torch.device("cuda", 0)
        t_node, t_sq = torch.from_numpy(node).to(dev), torch.from_numpy(sq).to(dev)
        t_mask = torch.from_numpy((sq < 0).astype(np.uint8)).to(dev)
        t_lab, t_loss = torch.from_numpy(lab.astype(np.float32)).to(dev), torch.zeros(1, dtype=torch.float32, device=dev)
        st = torch.cuda.Stream(dev)
        eng.set_stream(st.cuda_stream)
        torch.cuda.synchronize()
        for _ in range(3):
            step[0] += 1
            eng.train_step_dev(rows, t_node.data_ptr(), t_sq.data_ptr(), t_mask.data_ptr(), t_lab.data_ptr(), 1e-3, step[0], t_loss.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            step[0] += 1
            eng.train_step_dev(rows, t_node.data_ptr(), t_sq.data_ptr(), t_mask.data_ptr(), t_lab.data_ptr(), 1e-3, step[0], t_loss.data_ptr())
        e1.record(st)
        eng.synchronize()
        ddt = e0.elapsed_time(e1) / 20 * 1e-3
        eng.set_stream(0)
        emit(path="train_step_dev (device buffers: dmg_train_step_dev)", items=n_items, rows_per_step=rows, ms_per_step=ddt * 1e3, rows_per_s=rows / ddt,
             roofline={"bound": "hbm", "algorithmic_bytes_per_step": dense + sparse, "achieved": (dense + sparse) / ddt / 1e9, "peak": hbm,
                       "unit": "GB/s", "frac": (dense + sparse) / ddt / 1e9 / hbm, "peak_source": hbm_src})
        if n_items <= 1_000_000:
            params = eng.download_din_weights()
            t0 = time.perf_counter()
            g, loss = orc.din_gradients(params, rows_tab, E, T, node, sq, mask, lab)
            s = np.zeros_like(params); r = np.zeros_like(params)
            orc.adam_step(params, g, s, r, 1e-3, 1)
            cdt = time.perf_counter() - t0
            emit(path="train_step cpu_baseline", items=n_items, kind="port", cores=1, ms_per_step=cdt * 1e3, rows_per_s=rows / cdt,
                 sample="one step of the same batch, oracle/ C restatement (single thread)")
        eng.close()

    # ---- 2. model.forward on independent rows (dmg_score_pairs) and JTM item weights (a22) ------------------------
    n_items = 100_000 if a.quick else 1_000_000
    tf = synth.tdm_tree(n_items, seed=1)
    L = tf.max_level
    rows_tab = (1 << (L + 1)) - 1
    eng = Engine(0)
    eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    eng.init_din_weights(np.float32, rows_tab, E, T, seed=2)
    rng = np.random.Generator(np.random.PCG64(5))
    n = 200_000
    node = rng.integers(0, rows_tab, n).astype(np.int32)
    sq = rng.integers(0, rows_tab, (n, T)).astype(np.int32)
    dt = timeit(lambda: eng.score_pairs(node, sq), warm=1, reps=3)
    by = n * (1 + T) * E * 4 + n * 4
    emit(path="score_pairs (model.forward on independent rows)", rows=n, ms=dt * 1e3, rows_per_s=n / dt,
         roofline={"bound": "hbm", "algorithmic_bytes": by, "achieved": by / dt / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / dt / 1e9 / hbm},
         note="host buffers: H2D of the (1+T) indices per row and D2H of the logits are inside")
    import torch
    dev = torch.device("cuda", 0)
    t_node, t_sq, t_out = torch.from_numpy(node).to(dev), torch.from_numpy(sq).to(dev), torch.empty(n, dtype=torch.float32, device=dev)
    st = torch.cuda.Stream(dev)
    eng.set_stream(st.cuda_stream)
    torch.cuda.synchronize()
    for _ in range(2):
        eng.score_pairs_dev(n, t_node.data_ptr(), t_sq.data_ptr(), 0, t_out.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        eng.score_pairs_dev(n, t_node.data_ptr(), t_sq.data_ptr(), 0, t_out.data_ptr())
    e1.record(st)
    eng.synchronize()
    ddt = e0.elapsed_time(e1) / 10 * 1e-3
    eng.set_stream(0)
    emit(path="score_pairs_dev (device buffers: dmg_score_pairs_dev)", rows=n, ms=ddt * 1e3, rows_per_s=n / ddt,
         roofline={"bound": "hbm / fp32 FMA", "algorithmic_bytes": by, "achieved": by / ddt / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / ddt / 1e9 / hbm})
    params = eng.download_din_weights()
    model = orc.TdmModel(params, rows_tab, E, T)
    t0 = time.perf_counter()
    model.forward(node[:20000], sq[:20000])
    cdt = time.perf_counter() - t0
    emit(path="score_pairs cpu_baseline", kind="port", cores=1, rows_per_s=20000 / cdt, sample="20000 rows, oracle forward (single thread)")
    # JTM: 2000 items x 16 samples, gap 4 (16 candidate children, 4 path nodes each)
    n_it, n_s, gap, old = 2000, 16, 4, 8
    off = np.arange(n_it + 1, dtype=np.int64) * n_s
    sseq = synth.queries(n_it * n_s, T, n_items, seed=21)
    par = rng.integers((1 << old) - 1, (2 << old) - 1, n_it).astype(np.int32)
    dt = timeit(lambda: eng.jtm_item_weights(off, sseq, par, old, old + gap), warm=1, reps=3)
    scored = n_it * n_s * ((2 << gap) - 2)
    emit(path="jtm_item_weights (TreeLearning.aggregateWeights for a level step)", items=n_it, samples_per_item=n_s, gap=gap,
         scorer_rows=scored, ms=dt * 1e3, scorer_rows_per_s=scored / dt)
    eng.close()

    # ---- 2b. TDM retrieval with the DeepFM scorer: level-synchronous path of shard.cu, world 1 ----------
    n_items = 100_000 if a.quick else 1_000_000
    tf = synth.tdm_tree(n_items, seed=1)
    L = tf.max_level
    rows_tab = (1 << (L + 1)) - 1
    F = T + 1
    rng = np.random.Generator(np.random.PCG64(13))
    dparams = np.concatenate([rng.normal(0, 0.05, rows_tab * E), rng.normal(0, 0.05, F * F * E), np.zeros(F),
                              rng.normal(0, 0.3, F), [0.0]]).astype(np.float32)
    eng = Engine(0)
    eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    eng.load_deepfm_weights(dparams, rows_tab, E, T)
    B = 1024
    dq = synth.queries(B, T, n_items, seed=31)
    rows_u = 256 + 400 * (L - 8)
    by_u = rows_u * E * 4 + T * E * 4 + 10 * 8
    dt = timeit(lambda: eng.tdm_retrieve(dq, 200, 10), warm=2, reps=5)
    emit(path="tdm_retrieve with the DeepFM scorer (certified fast path: per-user hoisting, 12 dot products per row, strict re-scores; one batch in flight)",
         items=n_items, levels=L, batch=B, ms=dt * 1e3, users_per_s=B / dt, fast_stats=eng.fast_stats(),
         roofline={"bound": "hbm", "algorithmic_bytes_per_user": by_u, "achieved": by_u * B / dt / 1e9, "peak": hbm, "unit": "GB/s",
                   "frac": by_u * B / dt / 1e9 / hbm})
    # four host threads, one handle each over one copy of the tables (dmg_clone): batches in flight
    import threading
    NFd = 4
    engs = [eng] + [eng.clone() for _ in range(NFd - 1)]
    qs = [synth.queries(B, T, n_items, seed=40 + k) for k in range(NFd)]
    for k, e_ in enumerate(engs):
        e_.tdm_retrieve(qs[k], 200, 10)
    reps = 6

    def loop(k):
        for _ in range(reps):
            engs[k].tdm_retrieve(qs[k], 200, 10)
    th = [threading.Thread(target=loop, args=(k,)) for k in range(NFd)]
    t0 = time.perf_counter()
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    dtf = (time.perf_counter() - t0) / (reps * NFd)
    for e_ in engs[1:]:
        e_.close()
    emit(path="tdm_retrieve with the DeepFM scorer (certified fast path, 4 batches in flight)", items=n_items, levels=L, batch=B, ms=dtf * 1e3,
         users_per_s=B / dtf, roofline={"bound": "hbm", "algorithmic_bytes_per_user": by_u, "achieved": by_u * B / dtf / 1e9, "peak": hbm, "unit": "GB/s",
                                        "frac": by_u * B / dtf / 1e9 / hbm})
    eng.set_arithmetic("strict")
    dt = timeit(lambda: eng.tdm_retrieve(dq, 200, 10), warm=1, reps=3)
    emit(path="tdm_retrieve with the DeepFM scorer (level-synchronous, strict fp32)", items=n_items, levels=L, batch=B, ms=dt * 1e3,
         users_per_s=B / dt, roofline={"bound": "fp32 FMA pipe", "algorithmic_flop_per_user": rows_u * 2 * (F + 1) * F * E,
                                       "achieved_tflops": rows_u * 2 * (F + 1) * F * E * B / dt / 1e12,
                                       "hbm_algorithmic_gbs": by_u * B / dt / 1e9})
    eng.set_arithmetic("fast")
    dm = orc.TdmModel(dparams, rows_tab, E, T, deepfm=True)
    tree = orc.Tree.from_treefile(tf)
    t0 = time.perf_counter()
    oi, ol, oc = dm.retrieve_batch(tree, dq[:4 * threads], 200, 10, n_threads=threads)
    cdt = time.perf_counter() - t0
    gi, gl, gc = eng.tdm_retrieve(dq[:4 * threads], 200, 10)
    emit(path="tdm_retrieve DeepFM cpu_baseline", kind="port", cores=threads, users_per_s=4 * threads / cdt,
         parity={"ids_identical": bool((gi == oi).all()), "logits_bit_identical": bool((gl.view(np.uint32) == ol.view(np.uint32)).all())})
    eng.close()

    # ---- 3. OTM retrieval, fp64 (a13-a14) ------------------------------------------------------------------------
    n_items = 100_000 if a.quick else 1_000_000
    items, leaf_ids, leaf_level = synth.otm_mapping(n_items, seed=42)
    rows_tab = (1 << (leaf_level + 1)) - 1
    eng = Engine(0)
    eng.load_tree_complete(leaf_level, items, leaf_ids)
    eng.init_din_weights(np.float64, rows_tab, E, T, seed=2)
    B = 256
    rng = np.random.Generator(np.random.PCG64(3))
    lseq = leaf_ids[rng.integers(0, n_items, (B, T))].astype(np.int32)
    lseq[:, :3] = -1
    dt = timeit(lambda: eng.otm_retrieve(lseq, 200, 10), warm=1, reps=3)
    rows_u = 256 + 400 * (leaf_level - 8)
    emit(path="otm_retrieve (fp64 DIN scorer, complete tree)", items=n_items, levels=leaf_level, batch=B, ms=dt * 1e3, users_per_s=B / dt,
         roofline={"bound": "fp64 FMA pipe", "algorithmic_flop_per_user": rows_u * 27324, "achieved_tflops": rows_u * 27324 * B / dt / 1e12},
         note="strict fp64 chains (the reference's OTM model is Module[Double])")
    params = eng.download_din_weights()
    leaf_item = np.full(1 << leaf_level, -1, np.int32)
    leaf_item[leaf_ids - ((1 << leaf_level) - 1)] = items
    om = orc.OtmModel(params, rows_tab, E, T)
    t0 = time.perf_counter()
    oi, osc, oc = om.retrieve_batch(lseq[:2 * threads], leaf_level, 200, 10, leaf_item, n_threads=threads)
    cdt = time.perf_counter() - t0
    gi, gs, gc = eng.otm_retrieve(lseq[:2 * threads], 200, 10)
    emit(path="otm_retrieve cpu_baseline", kind="port", cores=threads, users_per_s=2 * threads / cdt,
         parity={"ids_identical": bool((gi == oi).all()), "scores_bit_identical": bool((gs.view(np.uint64) == osc.view(np.uint64)).all())})
    eng.close()

    otm_deepfm_section(a, E, T, threads)

    # ---- 3b. BASELINE configs[2]: OTM at 10 M items, fp64 -- one training step (LocalOptimizer per-level step: rows = users x 2 beam,
    #          otm/.../optim/LocalOptimizer.scala:55-109) and retrieval on the same 17 GB table ---------------------------------------
    if not a.quick:
        n_items = a.train_items
        items, leaf_ids, leaf_level = synth.otm_mapping(n_items, seed=42)
        rows_tab = (1 << (leaf_level + 1)) - 1
        eng = Engine(0)
        eng.load_tree_complete(leaf_level, items, leaf_ids)
        eng.init_din_weights(np.float64, rows_tab, E, T, seed=2)
        rng = np.random.Generator(np.random.PCG64(17))
        B = 256
        lseq = leaf_ids[rng.integers(0, n_items, (B, T))].astype(np.int32)
        lseq[:, :3] = -1
        dt = timeit(lambda: eng.otm_retrieve(lseq, 200, 10), warm=1, reps=3)
        rows_u = 256 + 400 * (leaf_level - 8)
        emit(path="otm_retrieve (fp64, BASELINE configs[2] catalogue)", items=n_items, levels=leaf_level, batch=B, ms=dt * 1e3, users_per_s=B / dt,
             roofline={"bound": "fp64 FMA pipe", "algorithmic_flop_per_user": rows_u * 27324, "achieved_tflops": rows_u * 27324 * B / dt / 1e12})
        # one level's training rows: 20 users x 400 beam nodes of the leaf level (batch 8192 / (2 beam)), labels = pseudo targets
        n_rows = 8000
        node = rng.integers((1 << leaf_level) - 1, (2 << leaf_level) - 1, n_rows).astype(np.int32)
        tseq = np.repeat(lseq[:20], 400, axis=0)
        lab = (rng.random(n_rows) < 0.05).astype(np.float64)
        mask = np.nonzero((tseq.ravel() < 0))[0].astype(np.int32)
        step = [0]

        def one64():
            step[0] += 1
            eng.train_step(node, tseq, mask, lab, 1e-3, step[0])
        dt = timeit(one64, warm=2, reps=4)
        n_par = rows_tab * E + 3 * E * E + 2 * E + 1
        by = 8 * n_par * 8 + 2 * n_rows * (1 + T) * E * 8
        emit(path="otm train_step (fp64 DIN fwd/bwd + BCE + scatter-add + dense Adam, BASELINE configs[2])", items=n_items, levels=leaf_level,
             rows_per_step=n_rows, ms_per_step=dt * 1e3,
             roofline={"bound": "hbm", "algorithmic_bytes_per_step": by, "achieved": by / dt / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": by / dt / 1e9 / hbm, "peak_source": hbm_src})
        eng.close()

    dr_section(a, T)
    dr_train_section(a, T)
    kmeans_section(a)


if __name__ == "__main__":
    main()
