#!/usr/bin/env python3
"""Host pipeline (hpss_pipeline_run) against the number of clip chunks, on the bench batch: features out (the e2e
leg) and moments only from 16-bit PCM with the signal preparation (the e2e_stats leg)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm_hpss_mtl_b200 import engine, synth  # noqa: E402

ctx = engine.get_context(0)
n, L = 4096, 16000
prm = engine.make_params(l_harm=31, l_perc=31, n_mels=120)
wave = engine.host_alloc(n * L)
wave[:] = synth.synth_batch_fast(n, L).ravel()
pcm = engine.host_alloc(n * L, np.int16)
pcm[:] = np.round(wave * 30000).astype(np.int16)
cls = (np.arange(n) % 3).astype(np.int32)
for nc in [int(x) for x in (sys.argv[1:] or [2, 4, 6, 8, 12, 16, 24, 32])]:
    pf = engine.Pipeline(ctx, [L] * n, prm, n_chunks=nc)
    ps = engine.Pipeline(ctx, [L] * n, prm, pcm_dtype=np.int16, prepare=True, n_chunks=nc)
    out = engine.host_alloc(pf.rows * pf.total_frames)
    res = []
    for fn in (lambda: pf.run(wave, feat_host=out), lambda: ps.run(pcm, clip_class=cls, n_classes=3, want_features=False)):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        res.append((time.perf_counter() - t0) / 10 * 1e3)
    print(f"chunks {pf.n_chunks:3d}: features out {res[0]:.2f} ms   pcm->moments {res[1]:.2f} ms", flush=True)
    pf.close(); ps.close()
