#!/usr/bin/env python3
"""BASELINE.json configs[2]: a MUSAN-scale synthetic corpus (1086 clips, ~103 h of 16 kHz audio by default)
sharded by clip over the ranks, feature extraction + raw moments per rank, ONE all-reduce of the moment vector,
then the reference's get_data_stats closed form (lib/preprocessing.py:461-586) on every rank.

    python tools/run_corpus.py [--scale 1.0]                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_corpus.py

Clip durations follow the shape of cross_validation_info/musan (660 music files, mean 232 s; 426 speech files,
mean 511 s; seeded gamma draws), the audio is synthetic and generated on the device sub-batch by sub-batch
(at most ~2 h of audio resident at a time); n_fft 400, hop 160, k = (21, 11), 120 mels as in the reference.
Prints one JSON line (rank 0): whole-job audio-seconds per second over the max-over-ranks device time.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine  # noqa: E402
from sm_hpss_mtl_b200.dist import allreduce_moments, finalize_stats, shard_clips  # noqa: E402

FS = 16000


def corpus(scale):
    rng = np.random.default_rng(2024)
    music = np.clip(rng.gamma(4.0, 232.0 / 4.0, size=660), 5.0, 1800.0)
    speech = np.clip(rng.gamma(3.0, 511.0 / 3.0, size=426), 5.0, 1800.0)
    dur = np.concatenate([music, speech]) * scale
    cls = np.concatenate([np.zeros(660, np.int32), np.ones(426, np.int32)])
    order = rng.permutation(len(dur))
    return [max(int(d * FS), 400) for d in dur[order]], cls[order]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale every clip duration (1.0 = ~103 h)")
    ap.add_argument("--sub-batch-hours", type=float, default=2.0)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = engine.get_context(local)
    if world > 1:                                   # NCCL communicator set-up outside the timed region
        warm = torch.zeros(8, dtype=torch.float64, device="cuda")
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    lengths, classes = corpus(args.scale)
    a, b = shard_clips(lengths, world)[rank]
    my_len, my_cls = lengths[a:b], classes[a:b]
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=120)
    D = engine.feature_rows(prm)
    acc = torch.zeros(2 * D + D + 2 + 1, dtype=torch.float64, device="cuda")
    # sub-batches of whole clips, at most sub_batch_hours of audio each
    cap = int(args.sub_batch_hours * 3600 * FS)
    subs, cur, cur_n = [], [], 0
    for i, L in enumerate(my_len):
        if cur and cur_n + L > cap:
            subs.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += L
    if cur:
        subs.append(cur)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for sub in subs:
        ls = [my_len[i] for i in sub]
        n = int(sum(ls))
        wave = torch.randn(n, generator=g, device="cuda", dtype=torch.float32).mul_(0.25)      # synthetic audio (untimed)
        wave += 0.5 * torch.sin(torch.arange(n, device="cuda", dtype=torch.float32) * (2 * np.pi * 330.0 / FS))
        batch = engine.Batch(ctx, clip_lengths=ls, n_fft=400, hop_length=160)
        out = torch.empty(D * batch.total_frames, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        ev0.record()
        engine.featuregram_moments(batch, wave, prm, [int(my_cls[i]) for i in sub], 2, out=out, acc=acc)
        ev1.record()
        torch.cuda.synchronize()
        total_ms += ev0.elapsed_time(ev1)
        del wave, out, batch
    if world > 1:
        dist.barrier()                              # time the collective itself, not the wait for the slowest rank
    ev0.record()
    allreduce_moments(acc)
    ev1.record()
    torch.cuda.synchronize()
    ar_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms + ar_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mean, std, counts = finalize_stats(acc.cpu().numpy(), D, 2)
    if rank == 0:
        audio_s = sum(lengths) / FS
        print(json.dumps({"config": "configs[2] MUSAN-scale synthetic corpus", "n_gpus": world, "clips": len(lengths),
                          "audio_hours": round(audio_s / 3600, 2), "ms_max_over_ranks": round(float(t.item()), 2),
                          "allreduce_ms": round(ar_ms, 3), "audio_s_per_s": round(audio_s / (float(t.item()) * 1e-3)),
                          "frames_counted": [int(c) for c in counts], "mean_range": [float(mean.min()), float(mean.max())],
                          "std_range": [float(std.min()), float(std.max())]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
