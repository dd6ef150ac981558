#!/usr/bin/env python3
"""Development helper: the bench step (hpss_featuregram_moments) for different batch sizes, same buffers reused back to
back -- does a batch whose intermediates fit the 126 MB L2 run faster per clip?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

ctx = engine.get_context(0)
prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=31, l_perc=31, n_mels=120)
D = engine.feature_rows(prm)
L = 16000
for n in (128, 256, 512, 1024, 2048, 4096):
    batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
    wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
    out = torch.empty(D * batch.total_frames, device="cuda")
    cls = (np.arange(n) % 3).astype(np.int32)
    acc = torch.zeros(3 * D + D + 4, dtype=torch.float64, device="cuda")
    reps = max(8, 4096 // n * 8)
    for _ in range(3):
        engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"clips {n:5d}: {ms:.4f} ms per step, {ms * 1e3 / n:.4f} us per clip, {n / ms / 1e3:.3f} M audio-s/s", flush=True)
