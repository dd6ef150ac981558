#!/usr/bin/env python3
"""Development helper for ncu: time-axis median at KH and frequency-axis median at KP on the 4096 x 1 s batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

kh, kp = int(os.environ.get("KH", 21)), int(os.environ.get("KP", 11))
n, L = 4096, 16000
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
S = engine.stft_mag(batch, wave, 400, 400, 160)
for _ in range(2):
    engine.median_time(batch, S, 201, kh)
    engine.median_freq(batch, S, 201, kp)
torch.cuda.synchronize()
