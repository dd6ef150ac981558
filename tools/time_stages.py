#!/usr/bin/env python3
"""Quick per-kernel timing of one bench step (development helper; BASELINE.json configs[1])."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

n, L, k = int(os.environ.get("N", 4096)), 16000, int(os.environ.get("K", 31))
KH, KP = int(os.environ.get("KH", k)), int(os.environ.get("KP", k))
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
cls = (np.arange(n) % 3).astype(np.int32)
acc = torch.zeros(3 * 240 + 240 + 4, dtype=torch.float64, device="cuda")
names = ["K1", "K2h", "K2p", "K3", "K3b+K5", "DCT20"]
tot = [0.0] * 6
mel = torch.from_numpy(engine.mel_filterbank(22050, 400, 120)).cuda()
reps = 10
for it in range(reps + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record()
    S = engine.stft_mag(batch, wave, 400, 400, 160); ev[1].record()
    harm = engine.median_time(batch, S, 201, KH); ev[2].record()
    perc = engine.median_freq(batch, S, 201, KP); ev[3].record()
    o2, cm2 = engine.mask_mel_log(batch, S, harm, perc, 201, mel_sr=22050, n_mels=120, log_power=1); ev[4].record()
    engine.topdb_moments(batch, o2, 120, 2, cm2, 80.0, cls, 3, acc=acc); ev[5].record()
    mf = engine.dct_mfcc(batch, o2, 120, 2, 20); ev[6].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i in range(6):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
print(os.environ.get("TAG", ""), " ".join(f"{nm}={t / reps:.3f}" for nm, t in zip(names, tot)), f"sum={sum(tot) / reps:.3f} ms")
