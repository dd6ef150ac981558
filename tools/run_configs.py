#!/usr/bin/env python3
"""Device-resident timing of every BASELINE.json configuration on ONE B200 (bench.py measures configs[1] only).

    python tools/run_configs.py [--out gpurun_out/configs.json]

Synthetic audio is generated on the device (normal noise + sinusoids, peak normalised per clip: the content
does not change the work).  Each line reports the median of 5 runs after 2 warm-ups, CUDA events on the
launching stream, and a sanity property of the result (finite, per-clip top_db floor).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine  # noqa: E402

FS = 16000


def make_wave(lengths, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    total = int(sum(lengths))
    x = torch.randn(total, generator=g, device="cuda", dtype=torch.float32) * 0.5
    t = torch.arange(total, device="cuda", dtype=torch.float32)
    for f0 in (220.0, 1330.0, 3100.0):
        x += 0.6 * torch.sin(t * (2 * np.pi * f0 / FS))
    off = 0
    for L in lengths:                     # the reference's normalisation (mean, then peak) per clip
        seg = x[off:off + L]
        seg -= seg.mean()
        seg /= seg.abs().max()
        off += L
    return x


def run(ctx, name, lengths, n_fft, win, hop, kh, kp, n_mels=120, reps=5):
    batch = engine.Batch(ctx, clip_lengths=lengths, n_fft=n_fft, hop_length=hop)
    prm = engine.make_params(n_fft=n_fft, win_length=win, hop_length=hop, l_harm=kh, l_perc=kp, n_mels=n_mels)
    wave = make_wave(lengths, 99)
    D = engine.feature_rows(prm)
    out = torch.empty(D * batch.total_frames, dtype=torch.float32, device="cuda")
    cls = (np.arange(len(lengths)) % 3).astype(np.int32)
    acc = torch.zeros(3 * D + D + 4, dtype=torch.float64, device="cuda")
    times = []
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc.zero_()
        e0.record()
        engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    audio_s = sum(lengths) / FS
    # sanity: finite, and per clip / stream min >= max - 80 (power_to_db top_db)
    ok = bool(torch.isfinite(out).all())
    c0 = batch.clip(out, D, 0).view(2, n_mels, -1)
    mx, mn = c0.amax(dim=(1, 2)), c0.amin(dim=(1, 2))
    ok = ok and bool((mn >= mx - 80.0 - 1e-3).all())
    F = n_fft // 2 + 1
    alg = ((4 * hop + 4 * F) + 8 * F + 8 * F + (12 * F + 8 * n_mels) + 16 * n_mels) * batch.total_frames
    row = {"config": name, "clips": len(lengths), "audio_s": round(audio_s, 1), "n_fft": n_fft, "hop": hop,
           "k": [kh, kp], "frames": int(batch.total_frames), "ms": round(ms, 3),
           "audio_s_per_s": round(audio_s / (ms * 1e-3), 0),
           "algorithmic_GBps_all_stages": round(alg / (ms * 1e-3) / 1e9, 1), "sane": ok}
    print(json.dumps(row), flush=True)
    del out, wave, batch
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--skip-corpus", action="store_true")
    args = ap.parse_args()
    ctx = engine.get_context(0)
    rows = []
    rows.append(run(ctx, "configs[0] one 10 s clip, k=(21,11)", [160000], 400, 400, 160, 21, 11))
    rows.append(run(ctx, "configs[0] one 10 s clip, k=(31,31)", [160000], 400, 400, 160, 31, 31))
    rows.append(run(ctx, "configs[1] 4096 x 1 s, k=31", [16000] * 4096, 400, 400, 160, 31, 31))
    if not args.skip_corpus:
        # configs[2]: one GPU's share (1/8) of a MUSAN-shaped corpus: 136 clips, mean 340 s (music 232 s / speech
        # 511 s in cross_validation_info/musan), 12.9 h of audio, k = (21, 11) as in the reference
        rng = np.random.default_rng(2024)
        d = np.concatenate([rng.gamma(4.0, 232.0 / 4.0, size=83), rng.gamma(3.0, 511.0 / 3.0, size=53)])
        lengths = [int(x * FS) for x in np.clip(d, 5.0, 1800.0)]
        rows.append(run(ctx, "configs[2] 1/8 of a MUSAN-scale corpus (136 clips), k=(21,11)", lengths, 400, 400, 160, 21, 11, reps=3))
    rows.append(run(ctx, "configs[3] one 1-hour stream, n_fft=2048 hop=512, k=31", [57600000], 2048, 2048, 512, 31, 31, reps=3))
    for n_fft in (512, 1024, 2048):
        for k in (17, 31, 63):
            rows.append(run(ctx, f"configs[4] sweep: 64 x 60 s, n_fft={n_fft}, k={k}", [60 * FS] * 64, n_fft, n_fft,
                            n_fft // 4, k, k, reps=3))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
