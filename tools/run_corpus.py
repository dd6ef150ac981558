#!/usr/bin/env python3
"""BASELINE.json configs[2]: a MUSAN-scale synthetic corpus (1086 clips, ~100 h of 16 kHz audio by default) sharded
by clip over the ranks, one corpus pass of get_data_stats per rank, ONE all-reduce of the moment vector, then the
reference's closed form (lib/preprocessing.py:461-586) on every rank.

    python tools/run_corpus.py [--scale 1.0] [--mode host|device]                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_corpus.py

  --mode host    (default) what a real pass does: the decoded files sit in (pinned) host memory as 16-bit PCM; every
                 rank runs hpss_pipeline_run over its slice: upload -> signal preparation (N2) -> STFT / HPSS / mel /
                 log -> raw moments; 8 KB come back per rank.  Timed by wall clock around the call, max over ranks.
  --mode device  the kernels alone: prepared float32 audio generated on the device sub-batch by sub-batch (at most
                 ~8 h resident), hpss_featuregram_moments per sub-batch, CUDA events, max over ranks.

Clip durations follow the shape of cross_validation_info/musan (660 music files, mean 232 s; 426 speech files, mean
511 s; seeded gamma draws); n_fft 400, hop 160, k = (21, 11), 120 mels as in the reference.  Every layout-dependent
table (tile tables, chunk plans) is built by an untimed first pass (`cold_ms` reports that pass, `ms` the second one).
Prints one JSON line (rank 0): whole-job audio-seconds per second over the max-over-ranks time.
"""
import argparse
import json
import os
import sys
import time

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
    return [max(int(d * FS), 1600) for d in dur[order]], cls[order]


def synth_pcm(lengths, seed):
    """16-bit PCM of the rank's clips in pinned host memory: every clip is a window of one long random base signal
    (noise + tones) with its own gain, plus a few near-silent stretches so that the silence excision has work."""
    rng = np.random.default_rng(seed)
    nb = 1 << 24
    t = np.arange(nb, dtype=np.float32)
    base = 0.25 * rng.standard_normal(nb, dtype=np.float32) + 0.5 * np.sin(t * np.float32(2 * np.pi * 330.0 / FS))
    base = np.clip(base * 12000.0, -32768, 32767).astype(np.int16)
    pcm = engine.host_alloc(int(sum(lengths)), np.int16)
    o = 0
    for L in lengths:
        a = int(rng.integers(0, nb))
        pos = 0
        while pos < L:                                   # wrap around the base signal
            n = min(L - pos, nb - a)
            pcm[o + pos:o + pos + n] = base[a:a + n]
            pos += n
            a = 0
        for _ in range(int(rng.integers(0, 4)) + L // (60 * FS)):          # silent stretches of 0.1 .. 1 s
            w = int(rng.integers(1600, 16000))
            if L > 2 * w:
                s = int(rng.integers(0, L - w))
                pcm[o + s:o + s + w] //= 256
        o += L
    return pcm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale every clip duration (1.0 = ~100 h)")
    ap.add_argument("--mode", default="host", choices=["host", "device"])
    ap.add_argument("--sub-batch-hours", type=float, default=8.0,
                    help="device mode: audio resident per sub-batch (~1.5 GB per hour; every sub-batch pays the ~0.2 ms "
                         "of fixed ramp-up / tail latency of five dependent kernels: 1 h 101 ms, 2 h 92 ms, 4 h 86 ms, 8 h 82 ms)")
    ap.add_argument("--chunk-hours", type=float, default=1.0, help="host mode: audio per pipeline chunk")
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
    shards = shard_clips(lengths, world)
    a, b = shards[rank]
    my_len, my_cls = lengths[a:b], [int(c) for c in classes[a:b]]
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=120)
    D = engine.feature_rows(prm)
    n_acc = 2 * D + D + 2 + 1
    share = [sum(lengths[s:e]) for s, e in shards]
    balance = max(share) / (sum(share) / world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.mode == "host":
        need = 2 * sum(lengths)                     # pinned bytes over all ranks of this box
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        except Exception:
            avail = 0
        if avail and need > 0.3 * avail:
            raise SystemExit(f"host mode would pin {need / 1e9:.1f} GB of {avail / 1e9:.1f} GB available: use a smaller --scale")
        pcm = synth_pcm(my_len, 100 + rank)
        hours = sum(my_len) / FS / 3600
        pl = engine.Pipeline(ctx, my_len, prm, pcm_dtype=np.int16, prepare=True, fs=FS,
                             n_chunks=max(1, int(np.ceil(hours / args.chunk_hours))))
        times = []
        for it in range(2):                         # pass 0 builds the per-layout tables (cold), pass 1 is the number
            mom = np.zeros(n_acc)
            barrier()
            t0 = time.perf_counter()
            pl.run(pcm, clip_class=my_cls, n_classes=2, moments=mom, want_features=False)
            acc = torch.from_numpy(mom).cuda()
            allreduce_moments(acc)                  # the one collective of the pass
            torch.cuda.synchronize()
            times.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        cold_ms, ms = times
        extra = {"pipeline_chunks_rank0": pl.n_chunks, "h2d_bytes_rank0": int(pcm.nbytes), "d2h_bytes_rank0": n_acc * 8,
                 "timing": "wall clock around hpss_pipeline_run + all-reduce, max over ranks"}
        pl.close()
    else:
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
        batches = [engine.Batch(ctx, clip_lengths=[my_len[i] for i in sub], n_fft=400, hop_length=160) for sub in subs]
        out = torch.empty(D * max(bt.total_frames for bt in batches), dtype=torch.float32, device="cuda")   # reused
        wave_buf = torch.empty(max(bt.total_samples for bt in batches), dtype=torch.float32, device="cuda")
        ramp = torch.arange(wave_buf.numel(), device="cuda", dtype=torch.float32) * (2 * np.pi * 330.0 / FS)
        g = torch.Generator(device="cuda")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for it in range(2):
            g.manual_seed(100 + rank)
            acc = torch.zeros(n_acc, dtype=torch.float64, device="cuda")
            total_ms = 0.0
            barrier()
            for sub, bt in zip(subs, batches):
                wave = wave_buf[:bt.total_samples]
                wave.normal_(generator=g).mul_(0.25).add_(torch.sin(ramp[:bt.total_samples]), alpha=0.5)   # synthetic audio (untimed)
                torch.cuda.synchronize()
                ev0.record()
                engine.featuregram_moments(bt, wave, prm, [my_cls[i] for i in sub], 2,
                                           out=out[:D * bt.total_frames], acc=acc)
                ev1.record()
                torch.cuda.synchronize()
                total_ms += ev0.elapsed_time(ev1)
            ev0.record()
            allreduce_moments(acc)                  # once per corpus pass
            ev1.record()
            torch.cuda.synchronize()
            times.append(max_over_ranks(total_ms + ev0.elapsed_time(ev1)))
        cold_ms, ms = times
        extra = {"sub_batches_rank0": len(subs), "timing": "CUDA events around hpss_featuregram_moments per sub-batch + "
                 "the all-reduce, max over ranks"}
    mean, std, counts = finalize_stats(acc.cpu().numpy(), D, 2)
    if rank == 0:
        audio_s = sum(lengths) / FS
        print(json.dumps({"config": "configs[2] MUSAN-scale synthetic corpus", "mode": args.mode, "n_gpus": world,
                          "clips": len(lengths), "audio_hours": round(audio_s / 3600, 2), "ms": round(ms, 2),
                          "cold_ms": round(cold_ms, 2), "audio_s_per_s": round(audio_s / (ms * 1e-3)),
                          "shard_imbalance_max_over_mean": round(balance, 4),
                          "frames_counted": [int(c) for c in counts], "mean_range": [float(mean.min()), float(mean.max())],
                          "std_range": [float(std.min()), float(std.max())], **extra}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
