#!/usr/bin/env python3
"""Time / profile the MFCC extension kernel alone on the configs[1] feature shape (4096 clips x 98 frames, 2 x 120 mels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine  # noqa: E402

n, T = 4096, 98
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_frames=[T] * n)
feat = torch.randn(240 * n * T, device="cuda") * 20 - 40
out = torch.empty(40 * n * T, device="cuda")
reps = int(os.environ.get("REPS", 20))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for it in range(reps + 3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.dct_mfcc(batch, feat, 120, 2, 20, out=out)
    e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        tot += e0.elapsed_time(e1)
ms = tot / reps
print(f"mult={os.environ.get('HPSS_DCT_GRID_MULT', '0')} dct20: {ms:.4f} ms, {(240 + 40) * 4 * n * T / ms / 1e6:.0f} GB/s algorithmic")
