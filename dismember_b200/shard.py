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
