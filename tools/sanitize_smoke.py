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

# equal clips (the dense bulk-copy median tiles, uniform STFT tiles, uniform moments) and a ragged batch with several
# tiles per line (TileWalk: contiguous tile ranges, clips without a frame are not possible here, see tests)
for lens, kh, kp in [([16000] * 9, 31, 31), ([16000] * 9, 21, 11), ([16000] * 3, 7, 63), ([60000, 3000, 45000, 400], 21, 11),
                     ([60000, 3000, 45000, 400], 31, 31)]:
    waves = np.concatenate([synth.synth_clip(i, L) for i, L in enumerate(lens)])
    batch = engine.Batch(ctx, clip_lengths=lens, n_fft=400, hop_length=160)
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=kh, l_perc=kp, n_mels=120)
    out, acc = engine.featuregram_moments(batch, torch.from_numpy(waves).cuda(), prm, [i % 2 for i in range(len(lens))], 2)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print("ok", lens[:2], kh, kp, engine.launch_count())

# signal preparation (N2) and the host pipeline on 16-bit PCM
rng = np.random.default_rng(0)
lens = [16000, 30000, 1700]
pcm = engine.host_alloc(sum(lens), np.int16)
pcm[:] = (rng.standard_normal(sum(lens)) * 3000).astype(np.int16)
pcm[4000:9000] //= 64                                         # a silent stretch
res = engine.prep_signals(ctx, torch.from_numpy(np.asarray(pcm)).cuda(), lens, fs=16000, win_length=400, hop_length=160,
                          alpha=0.025, beta=0.075, markers=True)
torch.cuda.synchronize()
prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=120)
pl = engine.Pipeline(ctx, lens, prm, pcm_dtype=np.int16, prepare=True, fs=16000, n_chunks=2)
mom = np.zeros(2 * 240 + 240 + 2 + 1)
pl.run(pcm, clip_class=[0, 1, 0], n_classes=2, moments=mom, want_features=False)
pl.close()
print("ok prep + pipeline", engine.launch_count())
