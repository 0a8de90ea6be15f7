#!/usr/bin/env python3
"""Generate tests/golden/* from the reference's bundled fixtures.

Run in the authoring container (needs /root/reference; the GPU box has no copy):

    python tools/make_golden.py

Inputs (read-only, decoded without a JVM -- see dismember_b200/formats):
  data/jtm/example_tree.bin    KV-protobuf tree, 7801 nodes / 3706 leaves, max_level 12
  data/jtm/example_model.bin   Java-serialised DIN(Float, E=16), compact vector of 131857
  data/otm/example_model.bin   Java-serialised DIN(Double, E=16)
  data/otm/example_mapping.txt 3706 "item leafId" lines
  data/dr/example_model.bin    Java-serialised DeepRetrieval (LayerModel + RerankModel)
  data/dr/example_mapping.bin  ItemSet proto (item, id, J=2 paths of D=3)
  data/jtm/train_data.csv      user_x_y,seq*10,target rows -> 256 sampled real histories

Outputs:
  tests/golden/jtm_fixture.npz, otm_fixture.npz, dr_fixture.npz   real inputs (exact bits)
  tests/golden/queries.npz                                         histories, targets, consumed sets
  tests/golden/oracle_outputs.npz   outputs of oracle/ on those inputs.  The JVM
      reference cannot run here, so these pin the ORACLE (regression + the
      values the CUDA path must reproduce), not the Scala program.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.setrecursionlimit(200000)

from dismember_b200.formats import javaser, tree_file  # noqa: E402

REF = os.environ.get("DMG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
CANONICAL = [0, 0, 2126, 204, 3257, 3439, 996, 1681, 3438, 1882]   # TdmModelTrainSpec.scala:85


def _tensor(t):
    st = t["_storage"]["values"]
    off = int(t["_storageOffset"])
    size = [int(x) for x in t["_size"]]
    n = int(np.prod(size))
    return st[off:off + n].reshape(size)


def din_params(path):
    objs = javaser.load_file(path)
    arrs = javaser.primitive_arrays(objs, 1000)
    return arrs[0][1]        # first big array = compact parameter storage (second = gradients)


def find_linear_params(obj):
    """collect DenseTensors out of a scala Vector/List serialisation proxy tree, in order"""
    found = []

    def fn(path, o):
        if isinstance(o, javaser.JavaObject) and o.classname.endswith("DenseTensor"):
            found.append(o)

    javaser.walk(obj, fn)
    return found


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(20251017))

    # ---- JTM / TDM fixture -------------------------------------------------
    tf = tree_file.read_tree(f"{REF}/data/jtm/example_tree.bin")
    p32 = din_params(f"{REF}/data/jtm/example_model.bin")
    assert p32.dtype == np.float32 and p32.size == 131857
    np.savez_compressed(f"{OUT}/jtm_fixture.npz", max_level=tf.max_level, codes=tf.codes, node_ids=tf.node_ids,
                        is_leaf=tf.is_leaf, prob=tf.prob, leaf_ids=tf.leaf_ids, leaf_codes=tf.leaf_codes,
                        params=p32, E=16, T=10)

    # ---- OTM fixture -------------------------------------------------------
    p64 = din_params(f"{REF}/data/otm/example_model.bin")
    assert p64.dtype == np.float64 and p64.size == 131857
    items, leaves = tree_file.read_otm_mapping(f"{REF}/data/otm/example_mapping.txt")
    np.savez_compressed(f"{OUT}/otm_fixture.npz", params=p64, items=items, leaf_ids=leaves, E=16, T=10)

    # ---- DR fixture --------------------------------------------------------
    dr = javaser.load_file(f"{REF}/data/dr/example_model.bin")[0]
    lm, rm = dr["layerModel"], dr["reRankModel"]
    num_item, K, D, T, E = (int(lm[k]) for k in ("numItem", "numNode", "numLayer", "seqLen", "embedSize"))
    layer_emb = _tensor(lm["embedParams"])
    lin = find_linear_params(lm["linearParams"])
    assert len(lin) == 2 * D
    layer_w = [_tensor(lin[2 * d]) for d in range(D)]
    layer_b = [_tensor(lin[2 * d + 1]) for d in range(D)]
    for d in range(D):
        assert layer_w[d].shape == (K, (T + d) * E) and layer_b[d].shape == (K,)
    rr_emb = _tensor(rm["embedParams"])
    rlin = find_linear_params(rm["linearParams"])
    rr_w, rr_b = _tensor(rlin[0]), _tensor(rlin[1])
    sm_w, sm_b = _tensor(rm["softmaxWeights"]), _tensor(rm["softmaxBiases"])
    assert rr_w.shape == (E, T * E) and sm_w.shape == (num_item, E)
    m_items, m_ids, m_paths = tree_file.read_dr_mapping(f"{REF}/data/dr/example_mapping.bin")
    np.savez_compressed(f"{OUT}/dr_fixture.npz", num_item=num_item, K=K, D=D, T=T, E=E, layer_emb=layer_emb,
                        **{f"layer_w{d}": layer_w[d] for d in range(D)},
                        **{f"layer_b{d}": layer_b[d] for d in range(D)},
                        rr_emb=rr_emb, rr_w=rr_w, rr_b=rr_b, sm_w=sm_w, sm_b=sm_b,
                        map_items=m_items, map_ids=m_ids, map_paths=m_paths)

    # ---- queries: canonical + 255 real histories from train_data.csv -------
    rows = []
    with open(f"{REF}/data/jtm/train_data.csv") as f:
        for line in f:
            parts = line.rstrip("\n").split(",")
            rows.append([int(x) for x in parts[1:]])
    rows = np.array(rows, np.int32)                 # [n, 11] = seq(10) + target
    pick = rng.choice(len(rows), size=255, replace=False)
    seqs = np.vstack([np.array(CANONICAL, np.int32)[None], rows[pick, :10]])
    targets = np.concatenate([[0], rows[pick, 10]]).astype(np.int32)
    cons_off = [0]
    cons = []
    all_items = tf.leaf_ids
    for u in range(len(seqs)):
        own = [int(x) for x in seqs[u] if x != 0]
        extra = rng.choice(all_items, size=int(rng.integers(0, 60)), replace=False).tolist()
        c = sorted(set(own + extra))
        cons.extend(c)
        cons_off.append(len(cons))
    np.savez_compressed(f"{OUT}/queries.npz", seqs=seqs, targets=targets,
                        cons_off=np.array(cons_off, np.int64), cons=np.array(cons, np.int32))

    # ---- oracle outputs ----------------------------------------------------
    from oracle import oracle as orc
    orc.build()
    out = {}
    tree = orc.Tree.from_treefile(tf)
    tdm = orc.TdmModel(p32, 8191, 16, 10)
    for beam in (20, 200):
        it, lg, ct = tdm.retrieve_batch(tree, seqs, beam, 10, n_threads=8)
        out[f"tdm_items_b{beam}"], out[f"tdm_logits_b{beam}"], out[f"tdm_counts_b{beam}"] = it, lg, ct
    it, lg, ct = tdm.retrieve_batch(tree, seqs, 20, 10, cons_off=np.array(cons_off, np.int64),
                                    cons=np.array(cons, np.int32), widen_beam=True, n_threads=8)
    out["tdm_eval_items"], out["tdm_eval_logits"], out["tdm_eval_counts"] = it, lg, ct
    raw_i, raw_l = tdm.recommend_raw(tree, seqs[0], 20)
    out["tdm_canon_raw_items"], out["tdm_canon_raw_logits"] = raw_i, raw_l

    otm = orc.OtmModel(p64, 8191, 16, 10)
    leaf_level = int(np.ceil(np.log(len(items)) / np.log(2)))
    leaf_item = np.full(1 << leaf_level, -1, np.int32)
    leaf_item[leaves - ((1 << leaf_level) - 1)] = items
    item_leaf = {int(a): int(b) for a, b in zip(items, leaves)}
    oseqs = np.array([[item_leaf.get(int(x), -1) for x in s] for s in seqs], np.int32)
    for beam in (20, 200):
        it, sc, ct = otm.retrieve_batch(oseqs, leaf_level, beam, 10, leaf_item, n_threads=8)
        out[f"otm_items_b{beam}"], out[f"otm_scores_b{beam}"], out[f"otm_counts_b{beam}"] = it, sc, ct
    bi, bs = otm.beam_search(oseqs[0], leaf_level, 20)
    out["otm_canon_beam_ids"], out["otm_canon_beam_scores"] = bi, bs

    drm = orc.DrModel(num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b)
    item_id = {int(a): int(b) for a, b in zip(m_items, m_ids)}
    dseqs = np.array([[item_id.get(int(x), -1) for x in s] for s in seqs], np.int32)
    for beam in (20, 50):
        P = np.zeros((len(dseqs), beam, D), np.int32)
        PR = np.zeros((len(dseqs), beam), np.float64)
        for u in range(len(dseqs)):
            p, pr = drm.beam_search(dseqs[u], beam)
            P[u, :len(p)], PR[u, :len(p)] = p, pr
        out[f"dr_paths_b{beam}"], out[f"dr_probs_b{beam}"] = P, PR
    np.savez_compressed(f"{OUT}/oracle_outputs.npz", **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("canonical TDM beam=20 top10:", out["tdm_items_b20"][0], out["tdm_logits_b20"][0])
    print("canonical OTM beam=20 top10:", out["otm_items_b20"][0], out["otm_scores_b20"][0])
    print("canonical DR  beam=20 paths[:3]:", out["dr_paths_b20"][0][:3], out["dr_probs_b20"][0][:3])


if __name__ == "__main__":
    main()
