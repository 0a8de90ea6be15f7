"""User sharding for multi-GPU retrieval: the path shards over users and every BASELINE
table fits one B200, so ranks are replicas -- same table everywhere, users split, no data-path
collective.  The split is the reference's own thread partition (Evaluator.scala:28-37): taskSize =
n / world, the first n % world ranks take one more."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    task, extra = n // world, n % world
    lo = rank * task + min(rank, extra)
    return lo, lo + task + (1 if rank < extra else 0)


def gather_rows(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """all_gather of per-rank result rows back into user order (torch.distributed, any backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((mx,) + local.shape[1:], local.dtype)
    pad[: len(local)] = local
    t = torch.from_numpy(pad)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.concatenate([o.cpu().numpy()[: hi - lo] for o, (lo, hi) in zip(outs, sizes)])


# ---- node table sharded by code range (csrc/shard.cu) ------------------------------------------
# Index arithmetic of the shard layout and a host-side model of the per-level exchange, written with
# torch.distributed so that it runs under gloo on CPU (tests/test_multiproc_gloo.py) and under NCCL.
# The product path is dmg_shard_tdm_retrieve (CUDA kernels + ncclSend/ncclRecv inside the library);
# this mirror pins the protocol: who owns a code, what travels (slot, code -> score), in which order
# scores are put back.  Its scorer is injected by the caller.

def _ilog2(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.int64)
    out = np.zeros(x.shape, np.int64)
    for s in (32, 16, 8, 4, 2, 1):
        m = x >> s > 0
        out[m] += s
        x = np.where(m, x >> s, x)
    return out


def code_level(codes: np.ndarray) -> np.ndarray:
    return _ilog2(np.asarray(codes, np.int64) + 1)


def shard_bits(world: int) -> int:
    if world < 1 or world & (world - 1):
        raise ValueError("world must be a power of two")
    return world.bit_length() - 1


def owner_of(codes: np.ndarray, world: int, self_rank: int) -> np.ndarray:
    """Rank that stores each code; codes on the replicated levels (< log2 world) belong to the asker."""
    g = shard_bits(world)
    c = np.asarray(codes, np.int64)
    lvl = _ilog2(c + 1)
    own = (c - ((1 << lvl) - 1)) >> np.maximum(lvl - g, 0)
    return np.where(lvl < g, self_rank, own).astype(np.int64)


def local_row(codes: np.ndarray, world: int) -> np.ndarray:
    """Row of a code in its owner's local table: replicated levels first, then one block per level."""
    g = shard_bits(world)
    c = np.asarray(codes, np.int64)
    lvl = _ilog2(c + 1)
    sh = np.maximum(lvl - g, 0)
    row = ((1 << g) - 1) + ((1 << sh) - 1) + ((c - ((1 << lvl) - 1)) & ((1 << sh) - 1))
    return np.where(lvl < g, c, row)


def global_row(rows: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Inverse of local_row on `rank`."""
    g = shard_bits(world)
    r = np.asarray(rows, np.int64)
    repl = (1 << g) - 1
    t = np.maximum(r - repl + 1, 1)
    sh = _ilog2(t)
    code = ((1 << (g + sh)) - 1) + (rank << sh) + (t - (1 << sh))
    return np.where(r < repl, r, code)


def local_rows(world: int, max_level: int) -> int:
    g = shard_bits(world)
    return ((1 << g) - 1) + ((1 << (max_level - g + 1)) - 1)


def _send_recv(outgoing, incoming, group=None):
    """The library's ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd round: point-to-point, empty messages skipped,
    the rank's own region handed over in place."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    incoming[rank].copy_(outgoing[rank])
    ops = []
    for p in range(len(outgoing)):
        if p == rank:
            continue
        if outgoing[p].numel():
            ops.append(dist.P2POp(dist.isend, outgoing[p], p, group))
        if incoming[p].numel():
            ops.append(dist.P2POp(dist.irecv, incoming[p], p, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_requests(slots: np.ndarray, codes: np.ndarray, owners: np.ndarray, score_owned, dtype=np.float32, group=None) -> np.ndarray:
    """The request / reply round every sharded path uses (TDM and JTM candidates by node owner, Deep Retrieval rerank candidates
    by item owner): (slot, code) pairs travel to `owners`, score_owned(requester, slots, codes) answers, replies return in
    request order.  -> scores aligned with `slots`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    idx = [np.nonzero(owners == p)[0] for p in range(world)]
    send = [np.stack([slots[i], codes[i]], 1).astype(np.int32) for i in idx]
    n_send = torch.tensor([len(x) for x in send], dtype=torch.int64)
    n_recv = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(n_recv, n_send, group=group)
    recv = [torch.empty((int(n), 2), dtype=torch.int32) for n in n_recv]
    _send_recv([torch.from_numpy(np.ascontiguousarray(x)) for x in send], recv, group)
    ans = []
    for p in range(world):
        r = recv[p].numpy()
        sc = score_owned(p, r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)) if len(r) else np.zeros(0, dtype)
        ans.append(torch.from_numpy(np.ascontiguousarray(sc, dtype)))
    back = [torch.empty(len(x), dtype=tdt) for x in send]
    _send_recv(ans, back, group)
    out = np.zeros(len(slots), dtype)
    for p in range(world):
        out[idx[p]] = back[p].numpy()
    return out


def item_owner(items: np.ndarray, num_item: int, world: int) -> np.ndarray:
    """Deep Retrieval item tables: contiguous ranges of ceil(num_item / world) items (csrc/dr.cu)."""
    chunk = (num_item + world - 1) // world
    return np.asarray(items, np.int64) // chunk


def exchange_scores(cand: np.ndarray, counts: np.ndarray, score_owned, group=None) -> np.ndarray:
    """One level of the sharded search for this rank's users.

    cand [B, cap] candidate codes (first counts[u] valid per user).  Requests (slot = u*cap + pos, code)
    go to the owner of the code (all_to_all), the owner answers with score_owned(requester_rank, slots,
    codes) -> float32 scores (all_to_all back), and the scores are put back at their slots."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, cap = cand.shape
    valid = np.arange(cap)[None, :] < np.asarray(counts)[:, None]
    slots = np.nonzero(valid.ravel())[0].astype(np.int64)
    codes = cand.ravel()[slots].astype(np.int64)
    own = owner_of(codes, world, rank)
    send = [np.stack([slots[own == p], codes[own == p]], 1).astype(np.int32) for p in range(world)]
    n_send = torch.tensor([len(x) for x in send], dtype=torch.int64)
    n_recv = torch.empty(world, dtype=torch.int64)
    dist.all_to_all_single(n_recv, n_send, group=group)
    recv = [torch.empty((int(n), 2), dtype=torch.int32) for n in n_recv]
    _send_recv([torch.from_numpy(np.ascontiguousarray(x)) for x in send], recv, group)
    ans = []
    for p in range(world):
        r = recv[p].numpy()
        s = score_owned(p, r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)) if len(r) else np.zeros(0, np.float32)
        ans.append(torch.from_numpy(np.ascontiguousarray(s, np.float32)))
    back = [torch.empty(len(x), dtype=torch.float32) for x in send]
    _send_recv(ans, back, group)
    out = np.zeros(B * cap, np.float32)
    for p in range(world):
        out[send[p][:, 0]] = back[p].numpy()
    return out.reshape(B, cap)


def make_sharded_engine(device: int, group=None):
    """Engine bound to `device` with the NCCL communicator of dmg_shard_init set up for the ranks of the
    torch.distributed group (the 128-byte NCCL id travels through torch.distributed, any backend)."""
    import torch.distributed as dist
    from ._capi import Engine
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    eng = Engine(device)
    box = [eng.shard_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    eng.shard_init(world, rank, box[0])
    return eng


def level_select_expand(cand: np.ndarray, score: np.ndarray, beam: int, exists) -> np.ndarray:
    """Host mirror of shard_select_expand_kernel for ONE user (Recommender.scala:75-92): keep the best
    `beam` by the stable descending sort when there are more, then the existing children in order."""
    from .jtm import stable_desc_order
    if len(cand) > beam:
        cand = cand[stable_desc_order(score)[:beam]]
    kids = np.stack([2 * cand.astype(np.int64) + 1, 2 * cand.astype(np.int64) + 2], 1).ravel()
    return kids[exists(kids)]
