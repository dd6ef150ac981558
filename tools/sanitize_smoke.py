#!/usr/bin/env python3
"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

ctx = engine.get_context(0)
Ls = [16000, 4000, 23457, 1601]          # ragged, one odd length (misaligned neighbours), one very short clip
for (n_fft, win, hop, kh, kp) in [(400, 400, 160, 31, 31), (400, 400, 160, 21, 11), (512, 400, 160, 15, 17), (1024, 1024, 256, 63, 4)]:
    lens = [max(L, n_fft) for L in Ls]
    waves = np.concatenate([synth.synth_clip(i, L) for i, L in enumerate(lens)])
    batch = engine.Batch(ctx, clip_lengths=lens, n_fft=n_fft, hop_length=hop)
    prm = engine.make_params(n_fft=n_fft, win_length=win, hop_length=hop, l_harm=kh, l_perc=kp, n_mels=40)
    wave = torch.from_numpy(waves).cuda()
    D = engine.feature_rows(prm)
    cls = [i % 3 for i in range(len(lens))]
    out, acc = engine.featuregram_moments(batch, wave, prm, cls, 3)
    F = n_fft // 2 + 1
    S = engine.stft_mag(batch, wave, n_fft, win, hop)
    harm = engine.median_time(batch, S, F, kh)
    perc = engine.median_freq(batch, S, F, kp)
    o1, cm = engine.mask_mel_log(batch, S, harm, perc, F, mel_sr=22050, n_mels=40, log_power=True)
    engine.row_standardize(batch, out.clone(), D)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print("ok", n_fft, hop, kh, kp, engine.launch_count())
