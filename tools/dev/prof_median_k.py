import os, sys, torch
sys.path.insert(0, os.getcwd())
from sm_hpss_mtl_b200 import engine, synth
k = int(os.environ.get("K", 7))
n, L = 4096, 16000
ctx = engine.get_context(0)
batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
wave = torch.from_numpy(synth.synth_batch_fast(n, L).ravel()).cuda()
S = engine.stft_mag(batch, wave, 400, 400, 160)
for _ in range(3):
    engine.median_time(batch, S, 201, k)
torch.cuda.synchronize()
