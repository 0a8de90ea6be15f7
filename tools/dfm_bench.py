"""DeepFM retrieval, certified fast path vs strict arithmetic on one handle, then four clones in flight (dev aid; the recorded
line lives in tools/bench_paths.py).  python tools/dfm_bench.py [n_items]"""
import sys, time, os, numpy as np
sys.path.insert(0, "/root/repo")
from dismember_b200 import Engine, synth
from oracle import oracle as orc
E, T = 64, 10
n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tf = synth.tdm_tree(n_items, seed=1)
L = tf.max_level
rows_tab = (1 << (L + 1)) - 1
F = T + 1
rng = np.random.Generator(np.random.PCG64(13))
dparams = np.concatenate([rng.normal(0, 0.05, rows_tab * E), rng.normal(0, 0.05, F * F * E), np.zeros(F), rng.normal(0, 0.3, F), [0.0]]).astype(np.float32)
eng = Engine(0)
eng.load_tree_tdm(L, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)
eng.load_deepfm_weights(dparams, rows_tab, E, T)
B = 1024
dq = synth.queries(B, T, n_items, seed=31)
for mode in ("fast", "strict"):
    eng.set_arithmetic(mode)
    eng.tdm_retrieve(dq, 200, 10)
    t0 = time.perf_counter()
    for _ in range(5):
        r = eng.tdm_retrieve(dq, 200, 10)
    dt = (time.perf_counter() - t0) / 5
    print(mode, "ms", dt * 1e3, "users/s", B / dt, eng.fast_stats() if mode == "fast" else "")
    if mode == "fast": rf = r
    else: print("fast == strict:", (rf[0] == r[0]).all(), (rf[1].view(np.uint32) == r[1].view(np.uint32)).all(), (rf[2] == r[2]).all())
# batches in flight: one clone per host thread
import threading
eng.set_arithmetic("fast")
NF = 4
engs = [eng] + [eng.clone() for _ in range(NF - 1)]
qs = [synth.queries(B, T, n_items, seed=40 + k) for k in range(NF)]
for k, e_ in enumerate(engs):
    e_.tdm_retrieve(qs[k], 200, 10)
reps = 8
def loop(k):
    for _ in range(reps):
        engs[k].tdm_retrieve(qs[k], 200, 10)
th = [threading.Thread(target=loop, args=(k,)) for k in range(NF)]
t0 = time.perf_counter()
[t.start() for t in th]; [t.join() for t in th]
dt = (time.perf_counter() - t0) / (reps * NF)
print("fast, %d in flight: ms %.3f users/s %.0f" % (NF, dt * 1e3, B / dt))
