#!/usr/bin/env python3
"""Development helper: K2h (time-axis median) on batches of the same size and different shapes (ns per frame)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine  # noqa: E402

k = int(os.environ.get("K", 21))
ROWS = int(os.environ.get("ROWS", 201))
ctx = engine.get_context(0)
rng = np.random.default_rng(2024)
d = np.concatenate([rng.gamma(4.0, 232.0 / 4.0, size=83), rng.gamma(3.0, 511.0 / 3.0, size=53)])
ragged = [int(x * 100) for x in np.clip(d, 5.0, 1800.0)]
mean = int(np.mean(ragged))
shapes = {"ragged 136 (MUSAN-shaped)": ragged, "136 equal clips": [mean] * 136, "136 clips, two lengths": [mean - 7000, mean + 7000] * 68,
          "sorted ragged": sorted(ragged), "16 x 300 s": [30001] * 16, "4096 x 1 s": [98] * 4096}
only = os.environ.get("ONLY")
for name, frames in shapes.items():
    if only and only not in name:
        continue
    batch = engine.Batch(ctx, clip_frames=frames)
    S = torch.rand(ROWS * batch.total_frames, device="cuda")
    for _ in range(2):
        engine.median_time(batch, S, ROWS, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        engine.median_time(batch, S, ROWS, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"k={k} {name:28s} frames {batch.total_frames:8d}  {ms:7.3f} ms  {ms * 1e6 / batch.total_frames:.3f} ns/frame", flush=True)
    del S, batch
