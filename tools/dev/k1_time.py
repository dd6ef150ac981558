#!/usr/bin/env python3
"""Time K1 (stft_mag, magnitudes only) alone on configs[1] with the L2 flushed between launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

n, L = int(os.environ.get("N", 4096)), int(os.environ.get("L", 16000))
NFFT, HOP = int(os.environ.get("NFFT", 400)), int(os.environ.get("HOP", 160))
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=NFFT, hop_length=HOP)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
reps = int(os.environ.get("REPS", 20))
ts = []
for it in range(reps + 3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S = engine.stft_mag(batch, wave, NFFT, min(NFFT, int(os.environ.get("WIN", NFFT))), HOP)
    e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"{os.environ.get('TAG', '')} K1 median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}  max {ts[-1]:.4f}  checksum {float(S.double().sum()):.6e}")
