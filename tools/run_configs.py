#!/usr/bin/env python3
"""Device-resident timing of every BASELINE.json configuration (bench.py measures configs[1] only), with the time of
every kernel, and next to it the reference's algorithm on the host cores.

    python tools/run_configs.py [--out gpurun_out/configs.json] [--cpu]                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_configs.py --sweep-only

Under torchrun every rank runs the same per-GPU workload on its own synthetic data (weak scaling, no collective in
the data path); times are the max over ranks, audio-s/s is the whole job.  Synthetic audio is generated on the device
(noise + sinusoids, peak normalised per clip: the content does not change the work).  Each row reports the median of
the runs after 2 warm-ups (CUDA events on the launching stream), the per-kernel split from a second pass through the
stage entry points, and a sanity property of the result (finite, per-clip top_db floor).

--cpu adds `cpu`: the oracle (librosa's algorithm on scipy / numpy) on ONE core over a bounded sample of the same
configuration (<= 60 s of audio), and that rate times the usable cores as the whole-box figure -- an EXTRAPOLATION,
labelled as such: the large configurations would take hours on the CPU.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FS = 16000


def configs(skip_corpus, sweep_only):
    rows = []
    if not sweep_only:
        rows.append(("configs[0] one 10 s clip, k=(21,11)", [160000], 400, 400, 160, 21, 11, 5))
        rows.append(("configs[0] one 10 s clip, k=(31,31)", [160000], 400, 400, 160, 31, 31, 5))
        rows.append(("configs[1] 4096 x 1 s, k=31", [16000] * 4096, 400, 400, 160, 31, 31, 5))
        rows.append(("configs[1] 4096 x 1 s, k=(21,11) (the reference's default kernels)", [16000] * 4096, 400, 400, 160, 21, 11, 5))
        if not skip_corpus:
            # configs[2]: one GPU's share (1/8) of a MUSAN-shaped corpus: 136 clips, mean 340 s (music 232 s / speech
            # 511 s in cross_validation_info/musan), ~12 h of audio, k = (21, 11) as in the reference
            rng = np.random.default_rng(2024)
            d = np.concatenate([rng.gamma(4.0, 232.0 / 4.0, size=83), rng.gamma(3.0, 511.0 / 3.0, size=53)])
            lengths = [int(x * FS) for x in np.clip(d, 5.0, 1800.0)]
            rows.append(("configs[2] 1/8 of a MUSAN-scale corpus (136 clips), k=(21,11)", lengths, 400, 400, 160, 21, 11, 3))
        rows.append(("configs[3] one 1-hour stream, n_fft=2048 hop=512, k=31", [57600000], 2048, 2048, 512, 31, 31, 3))
    for n_fft in (512, 1024, 2048):
        for k in (17, 31, 63):
            rows.append((f"configs[4] sweep: 64 x 60 s per GPU, n_fft={n_fft}, k={k}", [60 * FS] * 64, n_fft, n_fft,
                         n_fft // 4, k, k, 3))
    return rows


# ------------------------------------------------------------------------------------------------ CPU column
def _cpu_one(job):
    """Oracle on one core over a bounded sample of the configuration (at most 60 s of audio)."""
    name, lengths, n_fft, win, hop, kh, kp = job
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import preprocessing_oracle as po
    from sm_hpss_mtl_b200 import synth
    L = int(min(max(lengths), 60 * FS))
    n = max(1, min(len(lengths), (60 * FS) // L))
    ys = [synth.synth_clip(i, L) for i in range(n)]
    Tw, Ts = 1000.0 * win / FS, 1000.0 * hop / FS
    po.featuregram(ys[0][:max(n_fft * 4, 8000)], FS, Tw, Ts, kh, kp, n_fft, 120, "LogMelHarmPercSpec")
    t0 = time.perf_counter()
    for y in ys:
        po.featuregram(y, FS, Tw, Ts, kh, kp, n_fft, 120, "LogMelHarmPercSpec")
    dt = time.perf_counter() - t0
    return name, {"sample_audio_s": round(n * L / FS, 1), "seconds_1core": round(dt, 2),
                  "audio_s_per_s_1core": round(n * L / FS / dt, 1)}


def cpu_column(cfgs):
    import multiprocessing as mp
    from oracle.cpu_baseline import usable_cores
    cores = usable_cores()
    jobs = [(c[0], c[1], c[2], c[3], c[4], c[5], c[6]) for c in cfgs]
    with mp.get_context("spawn").Pool(min(cores, len(jobs))) as pool:
        res = dict(pool.map(_cpu_one, jobs, chunksize=1))
    for v in res.values():
        v["cores"] = cores
        v["audio_s_per_s_all_cores_extrapolated"] = round(v["audio_s_per_s_1core"] * cores, 1)
        v["note"] = ("oracle = librosa's algorithm on scipy.ndimage / numpy.fft / np.dot, one single-threaded process on a "
                     "bounded sample; the all-core figure is that rate x cores (extrapolated, not run)")
    return res


# ------------------------------------------------------------------------------------------------ GPU
def make_wave(torch, lengths, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    total = int(sum(lengths))
    x = torch.randn(total, generator=g, device="cuda", dtype=torch.float32) * 0.5
    t = torch.arange(total, device="cuda", dtype=torch.float32)
    for f0 in (220.0, 1330.0, 3100.0):
        x += 0.6 * torch.sin(t * (2 * np.pi * f0 / FS))
    del t
    if len(lengths) <= 256:
        off = 0
        for L in lengths:                     # the reference's normalisation (mean, then peak) per clip
            seg = x[off:off + L]
            seg -= seg.mean()
            seg /= seg.abs().max()
            off += L
    else:                                     # equal clips: vectorised
        v = x.view(len(lengths), -1)
        v -= v.mean(dim=1, keepdim=True)
        v /= v.abs().amax(dim=1, keepdim=True)
    return x


def run(torch, dist, engine, ctx, world, rank, name, lengths, n_fft, win, hop, kh, kp, reps, n_mels=120):
    batch = engine.Batch(ctx, clip_lengths=lengths, n_fft=n_fft, hop_length=hop)
    prm = engine.make_params(n_fft=n_fft, win_length=win, hop_length=hop, l_harm=kh, l_perc=kp, n_mels=n_mels)
    wave = make_wave(torch, lengths, 99 + rank)
    D = engine.feature_rows(prm)
    F = n_fft // 2 + 1
    out = torch.empty(D * batch.total_frames, dtype=torch.float32, device="cuda")
    cls = (np.arange(len(lengths)) % 3).astype(np.int32)
    acc = torch.zeros(3 * D + D + 4, dtype=torch.float64, device="cuda")

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        if world == 1:
            return x
        t = torch.tensor(x, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    times = []
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc.zero_()
        sync()
        e0.record()
        engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(max_ranks([float(np.median(times))])[0]) if world > 1 else float(np.median(times))
    # sanity: finite, and per clip / stream min >= max - 80 (power_to_db top_db)
    ok = bool(torch.isfinite(out).all())
    c0 = batch.clip(out, D, 0).view(2, n_mels, -1)
    mx, mn = c0.amax(dim=(1, 2)), c0.amin(dim=(1, 2))
    ok = ok and bool((mn >= mx - 80.0 - 1e-3).all())
    del out
    # per-kernel split through the stage entry points
    names = ["K1 stft_mag", "K2h median_time", "K2p median_freq", "K3 mask_mel_log", "K3b+K5 topdb_moments"]
    bpf = [4 * hop + 4 * F, 8 * F, 8 * F, 12 * F + 8 * n_mels, 8 * n_mels]
    tot = [0.0] * 5
    for it in range(reps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        acc.zero_()
        ev[0].record()
        S = engine.stft_mag(batch, wave, n_fft, win, hop); ev[1].record()
        harm = engine.median_time(batch, S, F, kh); ev[2].record()
        perc = engine.median_freq(batch, S, F, kp); ev[3].record()
        o, cmax = engine.mask_mel_log(batch, S, harm, perc, F, mel_sr=22050, n_mels=n_mels, log_power=1); ev[4].record()
        engine.topdb_moments(batch, o, n_mels, 2, cmax, 80.0, cls, 3, acc=acc); ev[5].record()
        torch.cuda.synchronize()
        if it >= 1:
            for i in range(5):
                tot[i] += ev[i].elapsed_time(ev[i + 1]) / reps
        del S, harm, perc, o, cmax
    if world > 1:
        tot = max_ranks(tot)
    audio_s = world * sum(lengths) / FS
    frames = int(batch.total_frames)
    stages = [{"kernel": nme, "ms": round(t, 4), "algorithmic_GBps": round(b * frames / (t * 1e-3) / 1e9, 1)}
              for nme, t, b in zip(names, tot, bpf)]
    alg = sum(bpf) * frames
    row = {"config": name, "n_gpus": world, "clips_per_gpu": len(lengths), "audio_s": round(audio_s, 1), "n_fft": n_fft,
           "hop": hop, "k": [kh, kp], "frames_per_gpu": frames, "ms": round(ms, 3),
           "audio_s_per_s": round(audio_s / (ms * 1e-3), 0),
           "algorithmic_GBps_all_stages_per_gpu": round(alg / (ms * 1e-3) / 1e9, 1), "stages": stages, "sane": ok}
    del wave, batch
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--skip-corpus", action="store_true")
    ap.add_argument("--sweep-only", action="store_true", help="only the configs[4] sweep (the multi-GPU runs)")
    ap.add_argument("--cpu", action="store_true", help="add the CPU column (oracle on the host cores, bounded sample)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    cfgs = configs(args.skip_corpus, args.sweep_only)
    cpu = cpu_column(cfgs) if (args.cpu and rank == 0) else {}        # before CUDA is initialised in this process
    import torch
    import torch.distributed as dist
    from sm_hpss_mtl_b200 import engine
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = engine.get_context(local)
    rows = []
    for (name, lengths, n_fft, win, hop, kh, kp, reps) in cfgs:
        row = run(torch, dist, engine, ctx, world, rank, name, lengths, n_fft, win, hop, kh, kp, reps)
        if name in cpu:
            row["cpu"] = cpu[name]
            row["gpu_over_cpu_all_cores_extrapolated"] = round(row["audio_s_per_s"] / cpu[name]["audio_s_per_s_all_cores_extrapolated"], 1)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
