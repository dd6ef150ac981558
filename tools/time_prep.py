#!/usr/bin/env python3
"""Timing of the signal-preparation stage (N2) on the device: the bench batch (4096 x 1 s) and a ragged MUSAN-like
batch, float32 and int16 PCM input.  Prints ms per call and algorithmic GB/s (20 B / 12 B per sample)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

ctx = engine.get_context(0)
rng = np.random.default_rng(0)
for name, lens in [("4096 x 1 s", [16000] * 4096), ("64 x 30..300 s", [int(x) for x in rng.integers(480000, 4800000, 64)])]:
    n = sum(lens)
    x = torch.randn(n, device="cuda") * 0.1
    # a few silent stretches per clip so that the excision path runs
    off = np.concatenate([[0], np.cumsum(lens)])
    for c in range(0, len(lens), 2):
        a = int(off[c] + lens[c] // 3)
        x[a:a + 3000] *= 1e-4
        b = int(off[c] + 2 * lens[c] // 3)
        x[b:b + 3000] *= 1e-4
    for dt, bps in ((torch.float32, 20), (torch.int16, 12)):
        pcm = x if dt == torch.float32 else (x * 20000).clamp(-32768, 32767).to(torch.int16)
        for _ in range(3):
            out, _ = engine.prep_signals(ctx, pcm, lens)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        ev0.record()
        for _ in range(reps):
            out, _ = engine.prep_signals(ctx, pcm, lens)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        print(f"{name:16s} {str(dt):14s} {ms:.3f} ms  {n / 16000 / ms * 1e3 / 1e6:.2f} M audio-s/s  "
              f"{bps * n / ms / 1e6:.0f} GB/s algorithmic ({bps} B/sample)")
