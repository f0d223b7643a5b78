"""Image sharding across GPUs: one process per GPU, contiguous batch split, weights replicated,
no collective on the data path -- every image's forward is independent (BatchNorm is eval-mode everywhere,
src/decoder.rs:129,139; src/aspp.rs:220,316,330).  `torch.distributed` is used only as plumbing: rank/world
discovery, barriers for timing, and (optionally) collecting the masks on rank 0.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous split of n images over `world` ranks; the first n % world ranks get one extra image."""
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def forward_sharded(forward: Callable[[np.ndarray], np.ndarray], x: np.ndarray, rank: int, world: int,
                    gather: bool = False, group=None) -> Optional[np.ndarray]:
    """Runs `forward` (e.g. BiRefNet.forward_logits) on this rank's contiguous slice of the global batch `x`.

    gather=False: returns the local masks [b_local,1,H,W] (the bench path: nothing crosses ranks).
    gather=True : rank 0 returns the full [B,1,H,W] in input order, other ranks return None; the exchange is a
                  host-side gather of the 1-channel masks (4 MiB per 1024^2 image), not part of the timed hot path.
    """
    lo, hi = shard_bounds(x.shape[0], world)[rank]
    local = forward(x[lo:hi]) if hi > lo else np.zeros((0, 1) + x.shape[2:], np.float32)
    if not gather:
        return local
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate(parts, axis=0)
