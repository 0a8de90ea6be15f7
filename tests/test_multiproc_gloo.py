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
