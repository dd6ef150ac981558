#!/usr/bin/env python3
"""Development helper: end-to-end time of hpss_featuregram_host on the bench batch (pinned host buffers)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth
n, L = 4096, 16000
ctx = engine.get_context(0)
prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=31, l_perc=31, n_mels=120)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
wave_host = engine.host_alloc(n * L)
wave_host[:] = synth.synth_batch_fast(n, L).ravel()
out_host = engine.host_alloc(240 * batch.total_frames)
for _ in range(3):
    engine.featuregram_host(batch, wave_host, prm, out_host)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    engine.featuregram_host(batch, wave_host, prm, out_host)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(os.environ.get("HPSS_HOST_CHUNKS", "16"), f"chunks: {dt * 1e3:.2f} ms per call, {n / dt / 1e3:.0f} k audio-s/s")
