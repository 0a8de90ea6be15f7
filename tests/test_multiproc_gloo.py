"""N>1 path on CPU: world_size-2 gloo run of the user-sharded retrieval plumbing (the scorer is the
CPU oracle here -- the sharding, gather and max-over-ranks logic is what is under test)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from dismember_b200 import shard
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = dict(np.load(os.path.join(ROOT, "tests", "golden", "jtm_fixture.npz")))
    q = dict(np.load(os.path.join(ROOT, "tests", "golden", "queries.npz")))
    tree = orc.Tree(int(f["max_level"]), f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    model = orc.TdmModel(f["params"], 8191, 16, 10)
    seqs = q["seqs"][:101]                                   # odd count: uneven shards
    lo, hi = shard.shard_range(len(seqs), rank, world)
    items, logits, counts = model.retrieve_batch(tree, seqs[lo:hi], 20, 10)
    all_items = shard.gather_rows(items, len(seqs))
    all_logits = shard.gather_rows(logits, len(seqs))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # bench.py's max-over-ranks timing reduction
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), items=all_items, logits=all_logits, tmax=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from dismember_b200.shard import shard_range
    for n in (0, 1, 7, 101, 1024):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_two_rank_gloo_matches_single_process(tmp_path, golden_out):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npz")
    assert (got["items"] == golden_out["tdm_items_b20"][:101]).all()
    assert (got["logits"].view(np.uint32) == golden_out["tdm_logits_b20"][:101].view(np.uint32)).all()
    assert got["tmax"][0] == 2.0


def _shard_worker(rank, world, port, out_dir):
    """Table-sharded retrieval protocol (csrc/shard.cu) under gloo: requests travel to the owner of each code,
    scores travel back; the scorer is the CPU oracle and REFUSES codes the rank does not own."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dismember_b200 import shard
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = dict(np.load(os.path.join(ROOT, "tests", "golden", "jtm_fixture.npz")))
    q = dict(np.load(os.path.join(ROOT, "tests", "golden", "queries.npz")))
    L, T, beam, topk = int(f["max_level"]), 10, 20, 10
    tree = orc.Tree(L, f["codes"], f["node_ids"], f["is_leaf"], f["leaf_ids"], f["leaf_codes"])
    model = orc.TdmModel(f["params"], 8191, 16, T)
    have = np.zeros(1 << (L + 1), bool)
    have[f["codes"]] = True
    exists = lambda c: have[c]
    leaf_item = np.full(1 << L, -1, np.int64)
    leaf_item[f["leaf_codes"] - ((1 << L) - 1)] = f["leaf_ids"]
    B = 20
    seqs = q["seqs"][rank * B:(rank + 1) * B]                 # users are sharded too, same B on every rank
    hist = np.stack([tree.id_to_code(s)[0] for s in seqs]).astype(np.int32)         # [B, T] codes, -1 = masked
    gathered = [None] * world
    dist.all_gather_object(gathered, hist)                    # mirror of the history-tile all-reduce
    hist_all = np.concatenate(gathered)
    cap = 2 * beam
    s_level = int(np.floor(np.log2(beam)))
    start = (1 << s_level) - 1
    first = np.arange(start, start + (1 << s_level))
    cand = np.zeros((B, cap), np.int64)
    counts = np.zeros(B, np.int64)
    score = np.zeros((B, cap), np.float32)
    beams = [first[exists(first)] for _ in range(B)]
    scored_remote = 0

    def score_owned(requester, slots, codes):
        nonlocal scored_remote
        lvl = shard.code_level(codes)
        assert ((shard.owner_of(codes, world, rank) == rank) | (lvl < shard.shard_bits(world))).all(), "asked for a row this rank does not own"
        assert (shard.local_row(codes, world) < shard.local_rows(world, L)).all()
        gu = requester * B + slots // cap
        sq = hist_all[gu]
        mask = np.nonzero((sq < 0).ravel())[0].astype(np.int32)
        if requester != rank:
            scored_remote += len(codes)
        return model.forward(codes.astype(np.int32), sq.astype(np.int32), mask)

    for level in range(s_level, L):
        for u in range(B):
            nxt = shard.level_select_expand(beams[u], score[u, :len(beams[u])], beam, exists)
            beams[u] = nxt
            counts[u] = len(nxt)
            cand[u, :len(nxt)] = nxt
        score = shard.exchange_scores(cand, counts, score_owned)
    from dismember_b200.jtm import stable_desc_order
    items = np.full((B, topk), -1, np.int32)
    logits = np.zeros((B, topk), np.float32)
    for u in range(B):
        it = leaf_item[beams[u] - ((1 << L) - 1)]
        sc = score[u, :len(beams[u])][it >= 0]
        it = it[it >= 0]
        order = stable_desc_order(sc)[:topk]
        items[u, :len(order)] = it[order]
        logits[u, :len(order)] = sc[order]
    out = [None] * world
    dist.all_gather_object(out, (items, logits, scored_remote))
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), items=np.concatenate([o[0] for o in out]),
                 logits=np.concatenate([o[1] for o in out]), remote=np.array([o[2] for o in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_layout_roundtrip():
    from dismember_b200 import shard
    for world in (1, 2, 4, 8):
        L = 9
        seen = np.zeros((1 << (L + 1)) - 1, np.int64)
        for rank in range(world):
            n = shard.local_rows(world, L)
            g = shard.global_row(np.arange(n), world, rank)
            assert (shard.local_row(g, world) == np.arange(n)).all()
            assert (shard.owner_of(g, world, rank) == rank).all()
            seen[g] += 1
        lvl = shard.code_level(np.arange(len(seen)))
        assert (seen[lvl >= shard.shard_bits(world)] == 1).all() and (seen[lvl < shard.shard_bits(world)] == world).all()


def test_two_rank_table_shard_protocol(tmp_path, golden_out):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    assert (got["items"] == golden_out["tdm_items_b20"][:40]).all()
    assert (got["logits"].view(np.uint32) == golden_out["tdm_logits_b20"][:40].view(np.uint32)).all()
    assert (got["remote"] > 0).all()                          # both ranks really scored rows for the other one


def _dr_shard_worker(rank, world, port, out_dir):
    """Deep Retrieval with item tables sharded by item range (csrc/dr.cu) under gloo: rerank candidates travel to the owner of the
    item, which scores them with the requester's user vector; the scorer is the oracle and refuses items outside its range."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from dismember_b200 import shard
    from dismember_b200.dr import build_path_csr
    from dismember_b200.jtm import stable_desc_order
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = dict(np.load(os.path.join(ROOT, "tests", "golden", "dr_fixture.npz")))
    q = dict(np.load(os.path.join(ROOT, "tests", "golden", "queries.npz")))
    D, K, N = int(f["D"]), int(f["K"]), int(f["num_item"])
    model = orc.DrModel(N, K, D, int(f["T"]), int(f["E"]), f["layer_emb"], [f[f"layer_w{d}"] for d in range(D)],
                        [f[f"layer_b{d}"] for d in range(D)], f["rr_emb"], f["rr_w"], f["rr_b"], f["sm_w"], f["sm_b"])
    off, flat = build_path_csr(f["map_ids"], f["map_paths"], K)
    item_id = {int(a): int(b) for a, b in zip(f["map_items"], f["map_ids"])}
    B, beam, topk = 12, 50, 10
    seqs_all = np.array([[item_id.get(int(x), -1) for x in s] for s in q["seqs"][:world * B]], np.int32)
    seqs = seqs_all[rank * B:(rank + 1) * B]
    # candidates of this rank's users (beam search is local once the history rows are there)
    cand_items, cand_user = [], []
    for u in range(B):
        paths, _ = model.beam_search(seqs[u], beam)
        for pth in paths:
            key = 0
            for c in pth:
                key = key * K + int(c)
            its = flat[off[key]:off[key + 1]]
            cand_items.extend(int(i) for i in its)
            cand_user.extend([u] * len(its))
    cand_items, cand_user = np.array(cand_items, np.int64), np.array(cand_user, np.int64)
    chunk = (N + world - 1) // world
    remote = 0

    def score_owned(requester, gci, items):
        nonlocal remote
        assert ((items >= rank * chunk) & (items < (rank + 1) * chunk)).all(), "asked for an item this rank does not own"
        users = requester_users[requester][gci]
        if requester != rank:
            remote += len(items)
        out = np.zeros(len(items), np.float64)
        for u in np.unique(users):                               # the oracle reranks per user vector
            m = users == u
            out[m] = model.rerank(seqs_all[requester * B + u], items[m].astype(np.int32))
        return out

    gathered = [None] * world
    dist.all_gather_object(gathered, cand_user)                  # mirror of the candidate-offset tables the owners use to find the user
    requester_users = gathered
    scores = shard.exchange_requests(np.arange(len(cand_items)), cand_items, shard.item_owner(cand_items, N, world), score_owned,
                                     dtype=np.float64)
    items = np.full((B, topk), -1, np.int32)
    out_sc = np.zeros((B, topk), np.float64)
    for u in range(B):
        m = np.flatnonzero(cand_user == u)
        order = np.argsort(-scores[m], kind="stable")[:topk]
        items[u, :len(order)] = cand_items[m][order]
        out_sc[u, :len(order)] = scores[m][order]
    want_i = np.full((B, topk), -1, np.int32)
    want_s = np.zeros((B, topk), np.float64)
    for u in range(B):
        oi, os_, _ = model.recommend(seqs[u], topk, beam, off, flat)
        want_i[u, :len(oi)] = oi
        want_s[u, :len(oi)] = os_
    res = [None] * world
    dist.all_gather_object(res, (bool((items == want_i).all()), bool((out_sc.view(np.uint64) == want_s.view(np.uint64)).all()), remote))
    if rank == 0:
        np.save(os.path.join(out_dir, "dr.npy"), np.array([[int(a), int(b), c] for a, b, c in res]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_deep_retrieval_item_shard_protocol(tmp_path):
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_dr_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "dr.npy")
    assert (r[:, 0] == 1).all() and (r[:, 1] == 1).all() and r[:, 2].sum() > 0       # rows really crossed ranks
