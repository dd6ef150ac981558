#!/usr/bin/env python3
"""Development helper: K2p (frequency-axis median, register walk) per frame against the number of resident-warp waves."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine  # noqa: E402

k = int(os.environ.get("K", 31))
ctx = engine.get_context(0)
slots = 148 * 16                       # resident warps at 128 registers
for waves in (1.0, 2.0, 3.0, 4.0, 5.0, 5.3, 6.0, 8.0, 16.0):
    frames = int(slots * 32 * waves)
    batch = engine.Batch(ctx, clip_frames=[frames])
    S = torch.rand(201 * frames, device="cuda")
    for _ in range(2):
        engine.median_freq(batch, S, 201, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        engine.median_freq(batch, S, 201, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"k={k} waves {waves:5.2f} frames {frames:8d}  {ms:7.4f} ms  {ms * 1e6 / frames:.4f} ns/frame", flush=True)
