#!/usr/bin/env python3
"""Three bench steps (BASELINE.json configs[1], hpss_featuregram_moments = 5 kernel launches each) for ncu.

    ncu --set full --clock-control none --import-source on -s 10 -c 5 -o gpurun_out/prof python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

n_clips = int(os.environ.get("HPSS_PROFILE_CLIPS", 4096))
k = int(os.environ.get("HPSS_PROFILE_K", 31))
n_fft = int(os.environ.get("HPSS_PROFILE_NFFT", 400))
hop = int(os.environ.get("HPSS_PROFILE_HOP", 160))
L = int(os.environ.get("HPSS_PROFILE_SAMPLES", 16000))
ctx = engine.get_context(0)
prm = engine.make_params(n_fft=n_fft, win_length=min(400, n_fft) if n_fft <= 512 else n_fft, hop_length=hop,
                         l_harm=k, l_perc=k, n_mels=120)
batch = engine.Batch(ctx, clip_lengths=[L] * n_clips, n_fft=n_fft, hop_length=hop)
wave = torch.from_numpy(synth.synth_batch_fast(n_clips, L).ravel()).cuda()
D = engine.feature_rows(prm)
out = torch.empty(D * batch.total_frames, device="cuda")
cls = (np.arange(n_clips) % 3).astype(np.int32)
acc = torch.zeros(3 * D + D + 4, dtype=torch.float64, device="cuda")
for _ in range(3):
    acc.zero_()
    engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
torch.cuda.synchronize()
print("launches", engine.launch_count())
