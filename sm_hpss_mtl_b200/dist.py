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


def stream_shard(n_samples: int, n_fft: int, hop_length: int, l_harm: int, rank: int, world_size: int):
    """Time-axis split of ONE long stream (BASELINE.json configs[3]) with read halos and no data exchange.

    Returns ((t0, t1), (a, b), (s0, s1)): the frames rank ``rank`` owns, the frames it has to compute (its own plus
    l_harm // 2 halo frames per side, so that the harmonic median of every owned frame sees true neighbours; at the
    ends of the stream the kernel's own reflection is the right thing) and the samples those frames cover.
    The frequency-axis median, the masks, the mel projection and the log are per frame; the only stream-wide quantity
    is power_to_db's maximum per stream (top_db), a 2-scalar MAX all-reduce (featuregram_stream_sharded)."""
    T = 1 + (int(n_samples) - int(n_fft)) // int(hop_length)
    if T < world_size:
        raise ValueError("fewer frames than ranks")
    t0, t1 = T * rank // world_size, T * (rank + 1) // world_size
    h = int(l_harm) // 2
    a, b = max(0, t0 - h), min(T, t1 + h)
    return (t0, t1), (a, b), (a * hop_length, (b - 1) * hop_length + n_fft)


def featuregram_stream_sharded(ctx, wave_slice, shard, params, allreduce_max=None):
    """Features of the frames this rank owns of one long stream: ``wave_slice`` = CUDA float32 samples [s0, s1) of
    ``stream_shard``.  Stage entry points (STFT, both medians, masks + mel + log) on the slice, the per-stream maximum
    over the OWNED frames only, ``allreduce_max`` (a callable on a 2-element CUDA float32 tensor; default:
    torch.distributed MAX all-reduce when initialised) and the top_db clip against the stream-wide maximum.
    Returns a (rows, t1 - t0) CUDA float32 tensor identical to the same columns of the unsharded featuregram."""
    import torch
    import torch.distributed as dist
    from . import engine
    (t0, t1), (a, b), _ = shard
    n_fft, hop, M = params.n_fft, params.hop_length, params.n_mels
    F = n_fft // 2 + 1
    batch = engine.Batch(ctx, clip_lengths=[wave_slice.numel()], n_fft=n_fft, hop_length=hop)
    if batch.total_frames != b - a:
        raise ValueError(f"slice has {batch.total_frames} frames, the shard says {b - a}")
    S = engine.stft_mag(batch, wave_slice, n_fft, params.win_length, hop)
    harm = engine.median_time(batch, S, F, params.l_harm)
    perc = engine.median_freq(batch, S, F, params.l_perc)
    out, _ = engine.mask_mel_log(batch, S, harm, perc, F, mel_sr=params.mel_sr, n_mels=M, log_power=1, amin=params.amin)
    own = out.view(2, M, b - a)[:, :, t0 - a:t1 - a].contiguous()
    mx = own.amax(dim=(1, 2))                                   # per stream, owned frames only
    if allreduce_max is not None:
        allreduce_max(mx)
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if params.top_db >= 0:
        own = torch.maximum(own, (mx - params.top_db).view(2, 1, 1))
    batch.close()
    return own.view(2 * M, t1 - t0)


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
