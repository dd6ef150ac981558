"""Multi-GPU plumbing: clip sharding and the one collective on the path.

Every stage of the front-end is per clip (medians never cross a clip, top_db and the per-file
StandardScaler are per clip), so the corpus is cut into one contiguous slice of clips per rank,
balanced by samples, with no data-path exchange.  The only collective is the sum all-reduce of
the raw feature moments that get_data_stats (lib/preprocessing.py:461-586) needs:
float64 [n_classes*D sums | D sums of squares | n_classes frame counts | non-finite count],
a few KB -- latency only, NCCL over NVLink/NVSwitch on GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_clips(clip_lengths: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) clip ranges, one per rank, balanced by total samples.

    Cut points are the clip boundaries closest to the ideal k/world_size quantiles of the
    cumulative sample count; every clip lands in exactly one rank, order is preserved."""
    n = len(clip_lengths)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cum = np.concatenate([[0], np.cumsum(np.asarray(clip_lengths, dtype=np.int64))])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        j = int(np.searchsorted(cum, target, side="left"))
        if j > 0 and j <= n and abs(cum[j - 1] - target) <= abs(cum[min(j, n)] - target):
            j -= 1
        j = min(max(j, cuts[-1]), n)
        cuts.append(j)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def moments_size(D: int, n_classes: int) -> int:
    return n_classes * D + D + n_classes + 1


def allreduce_moments(acc, group=None):
    """In-place SUM all-reduce of the moment vector (torch tensor, float64) over the process group.
    No-op when torch.distributed is not initialised (single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def finalize_stats(acc_host: np.ndarray, D: int, n_classes: int):
    """(mean f32[D], stdev f32[D], counts) from the (all-reduced) moment vector -- the closed form
    of the reference's two passes, evaluated by the C library (hpss_stats_finalize)."""
    from . import engine
    mean, std, counts, nonfinite = engine.stats_finalize(acc_host, D, n_classes)
    if nonfinite:
        raise FloatingPointError(
            f"{int(nonfinite)} non-finite feature values: the reference's behaviour is undefined here "
            f"(it drops feature rows and then fails on the shape mismatch, lib/preprocessing.py:507-529)")
    return mean, std, counts
