#!/usr/bin/env python3
"""Development helper: time the two median kernels alone for a list of kernel sizes (4096 x 1 s batch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

ks = [int(k) for k in (sys.argv[1] if len(sys.argv) > 1 else "11,15,19,21,23,27,31,35,39,47,51,63").split(",")]
n, L = 4096, 16000
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
S = engine.stft_mag(batch, wave, 400, 400, 160)
out = []
for k in ks:
    res = []
    for fn in (engine.median_time, engine.median_freq):
        for _ in range(2):
            fn(batch, S, 201, k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(batch, S, 201, k)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    out.append(f"k={k}: time {res[0]:.3f} freq {res[1]:.3f}")
print(os.environ.get("TAG", ""), " | ".join(out))
