"""Probe: does a second batch in flight (second engine, own stream, own host thread) fill the tail of the first?"""
import sys, time, threading
import numpy as np
import torch
sys.path.insert(0, ".")
from dismember_b200 import Engine, synth

n_eng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
items, E, T, B, K, W = 1_000_000, 64, 10, 1024, 64, 8
tf = synth.tdm_tree(items, seed=1)
rows = (1 << (tf.max_level + 1)) - 1
engs = []
for k in range(n_eng):
    e = Engine(0)
    e.load_tree_tdm(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
    e.init_din_weights(np.float32, rows, E, T, seed=2)
    e.set_arithmetic("fast")
    engs.append(e)
qs = [synth.queries(B, T, items, seed=4 + s) for s in range(W + K)]
for mode in ("host", "dev"):
    dq = [torch.from_numpy(q).cuda() for q in qs]
    outs = [(torch.empty((B, 10), dtype=torch.int32, device="cuda"), torch.empty((B, 10), dtype=torch.float32, device="cuda"),
             torch.empty((B,), dtype=torch.int32, device="cuda")) for _ in engs]
    def work(k, lo, hi):
        e = engs[k]
        for i in range(lo + k, hi, n_eng):
            if mode == "host":
                e.tdm_retrieve(qs[i], 200, 10)
            else:
                o = outs[k]
                e.tdm_retrieve_dev(B, dq[i].data_ptr(), 200, 10, True, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
        e.synchronize()
    def run(lo, hi):
        th = [threading.Thread(target=work, args=(k, lo, hi)) for k in range(n_eng)]
        [t.start() for t in th]; [t.join() for t in th]
    run(0, W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(W, W + K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"engines={n_eng} mode={mode}: {B * K / dt:.0f} users/s, {dt / K * 1e3:.4f} ms/step", flush=True)
