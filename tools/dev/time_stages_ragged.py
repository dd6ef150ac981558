#!/usr/bin/env python3
"""Per-kernel timing on a ragged MUSAN-shaped batch (development helper; k = 21/11 as in the reference)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine  # noqa: E402

FS = 16000
rng = np.random.default_rng(2024)
d = np.concatenate([rng.gamma(4.0, 232.0 / 4.0, size=83), rng.gamma(3.0, 511.0 / 3.0, size=53)])
lengths = [int(x * FS) for x in np.clip(d, 5.0, 1800.0)]
kh, kp = int(os.environ.get("KH", 21)), int(os.environ.get("KP", 11))
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=lengths, n_fft=400, hop_length=160)
wave = torch.randn(int(sum(lengths)), device="cuda") * 0.3
cls = (np.arange(len(lengths)) % 3).astype(np.int32)
acc = torch.zeros(3 * 240 + 240 + 4, dtype=torch.float64, device="cuda")
names = ["K1", "K2h", "K2p", "K3", "K3b+K5"]
tot = [0.0] * 5
reps = 3
for it in range(reps + 1):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    S = engine.stft_mag(batch, wave, 400, 400, 160); ev[1].record()
    harm = engine.median_time(batch, S, 201, kh); ev[2].record()
    perc = engine.median_freq(batch, S, 201, kp); ev[3].record()
    o, cm = engine.mask_mel_log(batch, S, harm, perc, 201, mel_sr=22050, n_mels=120, log_power=1); ev[4].record()
    engine.topdb_moments(batch, o, 120, 2, cm, 80.0, cls, 3, acc=acc); ev[5].record()
    torch.cuda.synchronize()
    if it >= 1:
        for i in range(5):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
    del S, harm, perc, o, cm
fr = batch.total_frames
print(f"frames {fr}", " ".join(f"{nm}={t / reps:.3f}ms({t / reps * 1e6 / fr:.2f}ns/fr)" for nm, t in zip(names, tot)), f"sum={sum(tot) / reps:.3f} ms")
