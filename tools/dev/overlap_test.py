#!/usr/bin/env python3
"""Experiment: does running the HBM-bound stages of one half of the batch next to the ALU-bound medians of the
other half (two CUDA streams, staged API, no shared workspace) shorten a step?  BASELINE.json configs[1]."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

n, L, k = 4096, 16000, 31
ctx = engine.get_context(0)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()


def stages(batch, w, cls, acc):
    S = engine.stft_mag(batch, w, 400, 400, 160)
    harm = engine.median_time(batch, S, 201, k)
    perc = engine.median_freq(batch, S, 201, k)
    o, cm = engine.mask_mel_log(batch, S, harm, perc, 201, mel_sr=22050, n_mels=120, log_power=1)
    engine.topdb_moments(batch, o, 120, 2, cm, 80.0, cls, 3, acc=acc)
    return o


def run(parts, reps=10, offset=False):
    m = n // parts
    batches = [engine.Batch(ctx, clip_lengths=[L] * m, n_fft=400, hop_length=160) for _ in range(parts)]
    waves = [wave[i * m * L:(i + 1) * m * L] for i in range(parts)]
    cls = (np.arange(m) % 3).astype(np.int32)
    accs = [torch.zeros(3 * 240 + 240 + 4, dtype=torch.float64, device="cuda") for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    times = []
    for it in range(reps + 3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        start = torch.cuda.Event()
        start.record()
        keep = []
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                s.wait_event(start)
                keep.append(stages(batches[i], waves[i], cls, accs[i]))
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
        del keep
    return float(np.median(times))


for parts in (1, 2, 4, 8):
    print(f"parts={parts}: {run(parts):.3f} ms", flush=True)
