"""GPU parity tests: every stage of the CUDA path against the CPU oracle, through the C ABI.

Bars (BASELINE.json north_star): medians and frame indexing bit-exact; soft-masked spectrograms
bit-exact given identical (S, harm, perc); STFT / features within 1e-4 relative L2 (max-abs
reported in the assertion message).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from oracle import preprocessing_oracle as po
from sm_hpss_mtl_b200._lib import check
from sm_hpss_mtl_b200 import engine, synth

pytestmark = pytest.mark.gpu

REL_L2_TOL = 1e-4     # north_star: "librosa-matching features within 1e-4 relative L2"


def rel_l2(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def to_dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def flat_batch(mats):
    """list of (rows, T_c) arrays -> flat float32 batch array."""
    return np.concatenate([np.ascontiguousarray(m, dtype=np.float32).ravel() for m in mats]) if mats else np.zeros(0, np.float32)


# ------------------------------------------------------------------------------ medians
MEDIAN_CASES = [
    # (rows, [T_c...], k)
    (201, [98], 31), (201, [98, 98, 98], 31), (201, [998], 21), (201, [998], 11),
    (201, [8], 21), (5, [8], 31), (40, [16], 16),           # multi-reflection, even k (generic path)
    (257, [300, 7, 131, 64], 17), (65, [33, 1, 2, 250], 63),
    (1025, [120], 31), (33, [70], 3), (33, [70], 5), (33, [70], 64), (17, [50], 101),
    # ragged batches with several tiles per line: more tiles than CTAs, clips without frames, whole empty line blocks
    (201, [3000, 17, 1500], 31), (40, [0, 0, 500, 0, 0, 0, 300], 21), (33, [0, 400, 0, 0, 5, 700, 0], 11),
]


@pytest.mark.parametrize("rows,Ts,k", MEDIAN_CASES)
def test_median_bit_exact(ctx, rows, Ts, k):
    rng = np.random.default_rng(rows * 1000 + k)
    mats = [np.abs(rng.standard_normal((rows, T))).astype(np.float32) for T in Ts]
    # ties: quantise a part of the data
    mats = [np.where(rng.random(m.shape) < 0.3, np.round(m * 4) / 4, m).astype(np.float32) for m in mats]
    batch = engine.Batch(ctx, clip_frames=Ts)
    S = to_dev(flat_batch(mats))
    harm = engine.median_time(batch, S, rows, k)
    perc = engine.median_freq(batch, S, rows, k)
    torch.cuda.synchronize()
    for c, (h, p) in enumerate(zip(batch.split(harm, rows), batch.split(perc, rows))):
        assert np.array_equal(h.cpu().numpy(), lr.median_filter_1d(mats[c], k, axis=1)), f"time axis, clip {c}, k={k}"
        assert np.array_equal(p.cpu().numpy(), lr.median_filter_1d(mats[c], k, axis=0)), f"freq axis, clip {c}, k={k}"
        # scipy itself, wherever it is well defined (see test_oracle.py::test_scipy_reflect_overshoot_bug)
        if lr.scipy_median_well_defined(Ts[c], k):
            assert np.array_equal(h.cpu().numpy(), lr.median_filter_scipy(mats[c], k, axis=1))
        if lr.scipy_median_well_defined(rows, k):
            assert np.array_equal(p.cpu().numpy(), lr.median_filter_scipy(mats[c], k, axis=0))


@pytest.mark.parametrize("k", [25, 31, 41, 63])
@pytest.mark.parametrize("n_clips,rows,T", [(1, 201, 98), (3, 201, 98), (5, 7, 33), (2, 64, 5), (4, 33, 150), (1, 1, 2), (9, 11, 27)])
def test_median_time_dense_tiles(ctx, k, n_clips, rows, T):
    """Batches of equal clips take the bulk-copy tile path for k >= 25 (median_dense_kernel): whole tiles and a
    partial last tile (lines not a multiple of 32, byte count not a multiple of 16), lines shorter than the halo
    (several reflections), canary regions around the output, and an input that is NOT 16-byte aligned (falls back
    to the cp.async ring): bit-exact against the oracle each time."""
    rng = np.random.default_rng(k * 131 + rows * 7 + T)
    mats = [np.abs(rng.standard_normal((rows, T))).astype(np.float32) for _ in range(n_clips)]
    mats = [np.where(rng.random(m.shape) < 0.3, np.round(m * 4) / 4, m).astype(np.float32) for m in mats]
    batch = engine.Batch(ctx, clip_frames=[T] * n_clips)
    flat = flat_batch(mats)
    for shift in (0, 1):                                   # 1: the input starts 4 bytes after a 16-byte boundary
        buf = torch.zeros(flat.size + 8, dtype=torch.float32, device="cuda")
        S = buf[shift:shift + flat.size]
        S.copy_(torch.from_numpy(flat))
        pad = 512
        big = torch.full((flat.size + 2 * pad,), -7.0, dtype=torch.float32, device="cuda")
        out = big[pad:pad + flat.size]
        check(batch.lib.hpss_median_time(batch.ctx.handle, batch.handle, C.c_void_p(S.data_ptr()), rows, k,
                                         C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert bool((big[:pad] == -7.0).all()) and bool((big[pad + flat.size:] == -7.0).all())
        for c, m in enumerate(mats):
            got = batch.clip(out, rows, c).cpu().numpy()
            assert np.array_equal(got, lr.median_filter_1d(m, k, axis=1)), (shift, c)


def test_median_all_fast_kernel_sizes(ctx):
    rows, T = 70, 75
    rng = np.random.default_rng(5)
    m = np.abs(rng.standard_normal((rows, T))).astype(np.float32)
    batch = engine.Batch(ctx, clip_frames=[T])
    S = to_dev(m.ravel())
    for k in range(1, 66):
        h = engine.median_time(batch, S, rows, k).cpu().numpy().reshape(rows, T)
        p = engine.median_freq(batch, S, rows, k).cpu().numpy().reshape(rows, T)
        assert np.array_equal(h, lr.median_filter_scipy(m, k, axis=1)), k
        assert np.array_equal(p, lr.median_filter_scipy(m, k, axis=0)), k


# ------------------------------------------------------------------------------ STFT
STFT_CASES = [
    # (n_fft, win, hop, [L...])
    (400, 400, 160, [16000]), (400, 400, 160, [16000, 400, 559, 560, 3000]), (400, 320, 160, [4001, 1234, 877]),
    (512, 400, 160, [16000, 8000]), (512, 512, 128, [9000]), (1024, 1024, 256, [20000]),
    (2048, 2048, 512, [40000]), (240, 200, 80, [5000]), (600, 600, 150, [7000]),
    # batches of equal clips whose length is a multiple of the hop: tiles of 16 virtual frames across clip borders
    (400, 400, 160, [16000] * 5), (512, 400, 160, [16000] * 3), (512, 512, 128, [12800] * 4),
    (1024, 1024, 256, [25600] * 3), (2048, 2048, 512, [51200] * 2), (400, 400, 160, [480] * 7),
]


@pytest.mark.parametrize("n_fft,win,hop,Ls", STFT_CASES)
def test_stft_matches_oracle(ctx, n_fft, win, hop, Ls):
    waves = [synth.synth_clip(i, L) for i, L in enumerate(Ls)]
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=n_fft, hop_length=hop)
    S, cplx = engine.stft_mag(batch, to_dev(np.concatenate(waves)), n_fft, win, hop, return_complex=True)
    torch.cuda.synchronize()
    F = n_fft // 2 + 1
    for c, y in enumerate(waves):
        want = lr.stft(y, n_fft=n_fft, hop_length=hop, win_length=win)
        assert batch.frames(c) == want.shape[1] == 1 + (len(y) - n_fft) // hop      # frame indexing exact
        got_c = batch.clip(cplx, F, c).cpu().numpy()
        got_m = batch.clip(S, F, c).cpu().numpy()
        e_c, e_m = rel_l2(got_c, want), rel_l2(got_m, np.abs(want))
        maxabs = float(np.abs(got_m - np.abs(want)).max())
        assert e_c < 1e-5 and e_m < 1e-5, f"rel-L2 complex {e_c:.2e} mag {e_m:.2e} max-abs {maxabs:.2e}"
    # the magnitudes-only call (the feature path; for 400/160 a different kernel: real-input 20 x 20 split)
    S0 = engine.stft_mag(batch, to_dev(np.concatenate(waves)), n_fft, win, hop)
    torch.cuda.synchronize()
    for c, y in enumerate(waves):
        want = np.abs(lr.stft(y, n_fft=n_fft, hop_length=hop, win_length=win))
        got = batch.clip(S0, F, c).cpu().numpy()
        assert got.shape == want.shape
        e = rel_l2(got, want)
        assert e < 1e-5, f"magnitude-only path, clip {c}: rel-L2 {e:.2e} max-abs {float(np.abs(got - want).max()):.2e}"


def test_stft_short_signal_raises(ctx):
    from sm_hpss_mtl_b200._lib import ParameterError
    with pytest.raises(ParameterError):
        engine.Batch(ctx, clip_lengths=[399], n_fft=400, hop_length=160)      # librosa: n_fft > len(y)
    with pytest.raises(lr.ParameterError):
        lr.stft(np.zeros(399, np.float32), n_fft=400, hop_length=160, win_length=400)


def test_stft_unsupported_nfft(ctx):
    from sm_hpss_mtl_b200._lib import HpssError
    batch = engine.Batch(ctx, clip_lengths=[4000], n_fft=14 * 2, hop_length=7)
    with pytest.raises(HpssError):
        engine.stft_mag(batch, torch.zeros(4000, device="cuda"), 28, 28, 7)     # 14 = 2 * 7


# ------------------------------------------------------------------------------ masks
@pytest.mark.parametrize("rows,Ts,kh,kp", [(201, [98, 60], 31, 31), (201, [300], 21, 11), (257, [120], 17, 17)])
def test_softmask_bit_exact(ctx, rows, Ts, kh, kp):
    n_fft = 2 * (rows - 1)
    mats = []
    for i, T in enumerate(Ts):
        y = synth.synth_clip(10 + i, n_fft + 160 * (T - 1))
        mats.append(np.abs(lr.stft(y, n_fft=n_fft, hop_length=160, win_length=min(400, n_fft))))
    mats[0][3:6, 4:9] = 0.0             # exact zeros -> split_zeros branch (mask 0.5)
    batch = engine.Batch(ctx, clip_frames=Ts)
    S = to_dev(flat_batch(mats))
    harm = engine.median_time(batch, S, rows, kh)
    perc = engine.median_freq(batch, S, rows, kp)
    out, _ = engine.mask_mel_log(batch, S, harm, perc, rows)
    torch.cuda.synchronize()
    for c, o in enumerate(batch.split(out, 2 * rows)):
        H, P = lr.hpss(mats[c], kernel_size=(kh, kp))
        got = o.cpu().numpy()
        assert np.array_equal(got[:rows], H), "harmonic masked spectrogram not bit-exact"
        assert np.array_equal(got[rows:], P), "percussive masked spectrogram not bit-exact"


def test_all_zero_input_gives_half_mask(ctx):
    rows, T = 201, 40
    batch = engine.Batch(ctx, clip_frames=[T])
    S = torch.zeros(rows * T, device="cuda")
    one = torch.ones(rows * T, device="cuda")
    out, _ = engine.mask_mel_log(batch, one, S, S, rows)     # harm = perc = 0 -> mask 0.5 -> 0.5 * S
    assert torch.equal(out, torch.full_like(out, 0.5))


# ------------------------------------------------------------------------------ mel table
@pytest.mark.parametrize("sr,n_fft,n_mels", [(22050, 400, 120), (16000, 400, 120), (22050, 512, 21), (22050, 2048, 128)])
def test_mel_table_matches_oracle(sr, n_fft, n_mels):
    got = engine.mel_filterbank(sr, n_fft, n_mels)
    want = lr.mel(sr, n_fft, n_mels)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= 1e-7
    assert np.array_equal(got > 0, want > 0)


# ------------------------------------------------------------------------------ featuregrams
FEATS = [("Spec", "SPEC"), ("LogSpec", "LOGSPEC"), ("MelSpec", "MELSPEC"), ("LogMelSpec", "LOGMELSPEC"),
         ("HarmPercSpec", "HARMPERC"), ("LogHarmPercSpec", "LOG_HARMPERC"), ("MelHarmPercSpec", "MEL_HARMPERC"),
         ("LogMelHarmPercSpec", "LOGMEL_HARMPERC")]


@pytest.mark.parametrize("featName,fid", FEATS)
@pytest.mark.parametrize("n_fft,lh,lp,n_mels", [(400, 21, 11, 120), (512, 31, 31, 21)])
def test_featuregram_matches_oracle(ctx, featName, fid, n_fft, lh, lp, n_mels):
    fs, Tw, Ts = 16000, 25, 10
    Ls = [16000, 48000, 1600, 7777]
    waves = [synth.synth_clip(100 + i, L) for i, L in enumerate(Ls)]
    mel_sr = fs if fid in ("MELSPEC", "LOGMELSPEC") else 22050
    prm = engine.make_params(n_fft=n_fft, win_length=400, hop_length=160, l_harm=lh, l_perc=lp, n_mels=n_mels,
                             mel_sr=mel_sr, feature=fid)
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=n_fft, hop_length=160)
    out = engine.featuregram(batch, to_dev(np.concatenate(waves)), prm)
    torch.cuda.synchronize()
    rows = engine.feature_rows(prm)
    for c, y in enumerate(waves):
        want = po.featuregram(y, fs, Tw, Ts, lh, lp, n_fft, n_mels, featName)
        got = batch.clip(out, rows, c).cpu().numpy()
        assert got.shape == want.shape and got.dtype == np.float32
        err, maxabs = rel_l2(got, want), float(np.abs(got - want).max())
        assert err < REL_L2_TOL, f"{featName} clip {c}: rel-L2 {err:.3e} (max-abs {maxabs:.3e})"


def test_featuregram_host_equals_device(ctx):
    Ls = [16000] * 37 + [5000, 123456]
    waves = np.concatenate([synth.synth_clip(i, L) for i, L in enumerate(Ls)])
    prm = engine.make_params(l_harm=31, l_perc=31)
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=400, hop_length=160)
    dev = engine.featuregram(batch, to_dev(waves), prm).cpu().numpy()
    pinned = engine.host_alloc(waves.size)
    pinned[:] = waves
    host = engine.featuregram_host(batch, pinned, prm)
    assert np.array_equal(dev, host)
    host2 = engine.featuregram_host(batch, waves, prm)        # pageable memory also works
    assert np.array_equal(dev, host2)


def test_from_spec_dafx_variant(ctx):
    rows, T = 201, 500
    y = synth.synth_clip(3, 400 + 160 * (T - 1))
    Spec = np.abs(lr.stft(y, n_fft=400, hop_length=160, win_length=400))
    batch = engine.Batch(ctx, clip_frames=[T])
    prm = engine.make_params(l_harm=21, l_perc=11, n_mels=120, feature="LOGMEL_HARMPERC")
    got = engine.featuregram_from_spec(batch, to_dev(Spec.ravel()), rows, prm).cpu().numpy().reshape(240, T)
    want = po.featuregram_from_spec(Spec, 21, 11, 120, "LogMelHarmPercSpec")
    assert rel_l2(got, want) < REL_L2_TOL


# ------------------------------------------------------------------------------ stats / patches
def test_moments_and_stats_match_oracle(ctx):
    D = 240
    rng = np.random.default_rng(11)
    Ts = [98, 40, 301, 77, 150, 64]
    cls = [0, 1, 2, 0, 1, 2]
    fvs = [(rng.standard_normal((D, T)) * 7 - 40).astype(np.float32) for T in Ts]
    batch = engine.Batch(ctx, clip_frames=Ts)
    acc = engine.moments(batch, to_dev(flat_batch(fvs)), D, cls, 3)
    mean, std, counts, bad = engine.stats_finalize(acc.cpu().numpy(), D, 3)
    names = ["music", "speech", "speech_music"]
    groups = {n: [fvs[i] for i in range(len(Ts)) if cls[i] == k] for k, n in enumerate(names)}
    want_mean, want_std, n0, n1, n2 = po.get_data_stats(groups, names)
    assert bad == 0 and list(counts) == [n0, n1, n2]
    assert np.allclose(mean, want_mean, rtol=1e-5, atol=1e-5)
    assert np.allclose(std, want_std, rtol=1e-5, atol=1e-5)
    # apply: Cython scale_data semantics (float64, eps 1e-10)
    got = engine.scale_data(batch, to_dev(flat_batch(fvs)), D, to_dev(mean), to_dev(std)).cpu().numpy()
    off = 0
    for fv in fvs:
        want = po.cscale_data(fv, mean, std)
        assert np.allclose(got[off:off + fv.size].reshape(fv.shape), want, rtol=1e-12, atol=1e-12)
        off += fv.size


def test_row_standardize_and_patches(ctx):
    D = 240
    rng = np.random.default_rng(12)
    Ts = [998, 300]
    fvs = [(rng.standard_normal((D, T)) * 5 - 30).astype(np.float32) for T in Ts]
    fvs[0][7] = 3.25                                         # constant row -> scale 1
    batch = engine.Batch(ctx, clip_frames=Ts)
    feat = to_dev(flat_batch(fvs))
    engine.row_standardize(batch, feat, D)
    for c, fv in enumerate(fvs):
        want = po._standard_scale_rows(fv)
        got = batch.clip(feat, D, c)
        assert np.allclose(got.cpu().numpy(), want, rtol=0, atol=2e-6)
        for (W, shift) in [(249, 24), (68, 68), (99, 1)]:
            p = engine.extract_patches(ctx, got.contiguous(), W, shift).cpu().numpy()
            wantp = po.extract_patches(got.cpu().numpy(), W, shift)
            assert p.dtype == np.float64 and p.shape == wantp.shape
            assert np.array_equal(p, wantp)


def test_featuregram_moments_fused_equals_separate(ctx):
    Ls = [16000] * 9 + [4000, 30000]
    cls = [i % 3 for i in range(len(Ls))]
    wave = to_dev(np.concatenate([synth.synth_clip(50 + i, L) for i, L in enumerate(Ls)]))
    for feat in ("LOGMEL_HARMPERC", "MEL_HARMPERC", "LOGSPEC", "SPEC"):
        prm = engine.make_params(l_harm=31, l_perc=31, feature=feat)
        batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=400, hop_length=160)
        D = engine.feature_rows(prm)
        ref = engine.featuregram(batch, wave, prm)
        acc_ref = engine.moments(batch, ref, D, cls, 3)
        out, acc = engine.featuregram_moments(batch, wave, prm, cls, 3)
        assert torch.equal(out, ref), feat
        a, b = acc.cpu().numpy(), acc_ref.cpu().numpy()
        assert np.allclose(a, b, rtol=1e-12, atol=1e-9), feat


@pytest.mark.parametrize("feat,lh,lp,n_fft", [("LOGMEL_HARMPERC", 31, 31, 400), ("LOGMEL_HARMPERC", 21, 11, 400),
                                              ("MEL_HARMPERC", 17, 63, 512), ("HARMPERC", 21, 11, 400),
                                              ("LOG_HARMPERC", 31, 5, 2048), ("LOGMEL_HARMPERC", 31, 16, 400)])
def test_fused_freq_median_equals_staged_pipeline(ctx, feat, lh, lp, n_fft):
    """hpss_featuregram runs the frequency median fused with masks + mel + log; the stage entry points run
    them as separate kernels.  Same arithmetic in the same order -> bit-identical (l_perc = 16 has no
    generated network and exercises the un-fused fallback inside hpss_featuregram)."""
    hop = 160 if n_fft <= 512 else 512
    win = 400 if n_fft <= 512 else n_fft
    Ls = [16000, 5 * n_fft + 3, 40000, n_fft]
    wave = to_dev(np.concatenate([synth.synth_clip(70 + i, L) for i, L in enumerate(Ls)]))
    prm = engine.make_params(n_fft=n_fft, win_length=win, hop_length=hop, l_harm=lh, l_perc=lp, n_mels=120, feature=feat)
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=n_fft, hop_length=hop)
    fused = engine.featuregram(batch, wave, prm)
    F = n_fft // 2 + 1
    S = engine.stft_mag(batch, wave, n_fft, win, hop)
    harm = engine.median_time(batch, S, F, lh)
    perc = engine.median_freq(batch, S, F, lp)
    is_mel, is_log = "MEL" in feat, "LOG" in feat
    mel = to_dev(engine.mel_filterbank(22050, n_fft, 120)) if is_mel else None
    staged, cmax = engine.mask_mel_log(batch, S, harm, perc, F, mel=mel, log_power=is_log)
    if is_log:
        engine.topdb_clip(batch, staged, 120 if is_mel else F, 2, cmax, 80.0)
    assert torch.equal(fused, staged)


def test_softmask_bit_exact_adversarial(ctx):
    """The kernel shares one refined reciprocal between the two mask divisions; check bit-exactness against
    numpy over a wide dynamic range: ratios near 1, tiny ratios (q*q subnormal / zero), subnormal inputs,
    exact zeros, huge values."""
    rng = np.random.default_rng(99)
    rows, T = 256, 4096
    n = rows * T
    expo = rng.uniform(-44, 12, size=n)
    h = (10.0 ** expo * rng.uniform(1, 10, size=n)).astype(np.float32)
    ratio_kind = rng.integers(0, 6, size=n)
    ratio = np.where(ratio_kind == 0, 1.0 + rng.uniform(-1e-6, 1e-6, size=n),
             np.where(ratio_kind == 1, 10.0 ** rng.uniform(-25, 0, size=n),
             np.where(ratio_kind == 2, 10.0 ** rng.uniform(0, 25, size=n),
             np.where(ratio_kind == 3, rng.uniform(0.5, 2.0, size=n), 10.0 ** rng.uniform(-3, 3, size=n)))))
    p = (h.astype(np.float64) * ratio).astype(np.float32)
    p[~np.isfinite(p)] = 1.0
    zero = rng.random(n) < 0.01
    h[zero] = 0.0
    p[rng.random(n) < 0.01] = 0.0
    s = np.abs(rng.standard_normal(n)).astype(np.float32) * 10
    h, p, s = h.reshape(rows, T), p.reshape(rows, T), s.reshape(rows, T)
    batch = engine.Batch(ctx, clip_frames=[T])
    out, _ = engine.mask_mel_log(batch, to_dev(s.ravel()), to_dev(h.ravel()), to_dev(p.ravel()), rows)
    got = out.cpu().numpy().reshape(2 * rows, T)
    with np.errstate(all="ignore"):
        mh = lr.softmask(h, p, power=2.0, split_zeros=True)
        mp = lr.softmask(p, h, power=2.0, split_zeros=True)
        want_h, want_p = s * mh, s * mp
    assert np.array_equal(got[:rows], want_h), int((got[:rows] != want_h).sum())
    assert np.array_equal(got[rows:], want_p), int((got[rows:] != want_p).sum())


@pytest.mark.parametrize("n_fft,sr,n_mels,Ts", [(400, 22050, 120, [98, 7, 300, 33]), (400, 16000, 21, [64]),
                                                (512, 22050, 120, [50, 50]), (2048, 22050, 128, [40, 9])])
def test_mask_mel_sweep_equals_dense_basis(ctx, n_fft, sr, n_mels, Ts):
    """hpss_mask_mel_log_sr (single-sweep kernel on the cached Slaney basis) against hpss_mask_mel_log with the
    same basis passed as a dense matrix: identical bits while the dense kernel stages whole columns (same
    f-ascending fmaf order), 1e-6 relative L2 where it accumulates in 64-row chunks (n_fft = 2048)."""
    rng = np.random.default_rng(n_fft + n_mels)
    F = n_fft // 2 + 1
    mats = [np.abs(rng.standard_normal((F, T))).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 3, (F, 1)).astype(np.float32)
            for T in Ts]
    batch = engine.Batch(ctx, clip_frames=Ts)
    S = to_dev(flat_batch(mats))
    harm = engine.median_time(batch, S, F, 17)
    perc = engine.median_freq(batch, S, F, 17)
    mel = to_dev(engine.mel_filterbank(sr, n_fft, n_mels))
    for log_power in (False, True):
        dense, cm_d = engine.mask_mel_log(batch, S, harm, perc, F, mel=mel, log_power=log_power)
        sweep, cm_s = engine.mask_mel_log(batch, S, harm, perc, F, mel_sr=sr, n_mels=n_mels, log_power=log_power)
        if F <= 264:
            assert torch.equal(dense, sweep)
            if log_power:
                assert torch.equal(cm_d, cm_s)
        else:
            assert rel_l2(sweep.cpu().numpy(), dense.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("k", [31, 15, 21, 11])
@pytest.mark.parametrize("n_fft,Ts", [(400, [98, 5, 130, 33]), (512, [40, 77]), (2048, [24])])
def test_median_freq_walk_dynamic_range(ctx, k, n_fft, Ts):
    """Frequency-axis register walk (stateful steps for K = 4G - 1, stateless groups otherwise) on spectrogram rows
    spread over eight decades: bit-exact against scipy."""
    rng = np.random.default_rng(k * 1000 + n_fft)
    F = n_fft // 2 + 1
    mats = [np.abs(rng.standard_normal((F, T))).astype(np.float32) * np.float32(10.0) ** rng.integers(-5, 3, (F, 1)).astype(np.float32)
            for T in Ts]
    batch = engine.Batch(ctx, clip_frames=Ts)
    S = to_dev(flat_batch(mats))
    perc = engine.median_freq(batch, S, F, k)
    for c, m in enumerate(mats):
        assert np.array_equal(batch.clip(perc, F, c).cpu().numpy(), lr.median_filter_scipy(m, k, axis=0))


def test_no_out_of_bounds_writes_canary(ctx):
    """compute-sanitizer is closed on the GPU pool, so out-of-bounds writes are looked for by hand: the feature
    and moment buffers sit between canary regions that must stay untouched (ragged batch, odd clip lengths,
    both median kernel families, clip + moments fused)."""
    Ls = [16000, 4001, 23457, 1601, 400]
    cls = [0, 1, 2, 0, 1]
    wave = to_dev(np.concatenate([synth.synth_clip(300 + i, L) for i, L in enumerate(Ls)]))
    for (lh, lp) in [(31, 31), (21, 11), (15, 64)]:
        prm = engine.make_params(l_harm=lh, l_perc=lp, n_mels=40)
        batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=400, hop_length=160)
        D = engine.feature_rows(prm)
        n = D * batch.total_frames
        pad = 4096
        big = torch.full((n + 2 * pad,), float("nan"), dtype=torch.float32, device="cuda")
        n_acc = 3 * D + D + 3 + 1
        big_acc = torch.full((n_acc + 64,), -7.0, dtype=torch.float64, device="cuda")
        acc = big_acc[32:32 + n_acc]
        acc.zero_()
        out = big[pad:pad + n]
        engine.featuregram_moments(batch, wave, prm, cls, 3, out=out, acc=acc)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        assert torch.isnan(big[:pad]).all() and torch.isnan(big[pad + n:]).all()
        assert (big_acc[:32] == -7.0).all() and (big_acc[32 + n_acc:] == -7.0).all()
        ref = engine.featuregram(batch, wave, prm)
        assert torch.equal(ref, out)


@pytest.mark.parametrize("frames,rows,k", [([98] * 37, 201, 31), ([98] * 37, 201, 21), ([98] * 5, 201, 7), ([101] * 33, 64, 11),
                                            ([3000, 17, 1500], 201, 21), ([700, 0, 5, 333], 33, 31), ([98, 60, 131], 201, 63),
                                            ([50], 17, 101)])
def test_median_outputs_stay_inside_their_buffers(ctx, frames, rows, k):
    """Canaries around the outputs of both median axes for every kernel family: dense bulk-copy tiles (equal clips),
    the cp.async ring with the prefix walk (ragged, several tiles per line), the register walk and the generic
    rank-counting fallback.  The input sits between canaries too (a bulk copy that overran would read them: +inf
    would show up in the medians)."""
    rng = np.random.default_rng(len(frames) * 131 + k)
    n = rows * int(sum(frames))
    pad = 8192
    big_in = torch.full((n + 2 * pad,), float("inf"), dtype=torch.float32, device="cuda")
    x = np.abs(rng.standard_normal(n)).astype(np.float32)
    big_in[pad:pad + n] = to_dev(x)
    S = big_in[pad:pad + n]
    batch = engine.Batch(ctx, clip_frames=frames)
    for fn, axis in ((engine.median_time, 1), (engine.median_freq, 0)):
        big = torch.full((n + 2 * pad,), float("nan"), dtype=torch.float32, device="cuda")
        out = big[pad:pad + n]
        fn(batch, S, rows, k, out=out)
        torch.cuda.synchronize()
        assert torch.isnan(big[:pad]).all() and torch.isnan(big[pad + n:]).all(), f"axis {axis}: canary overwritten"
        assert torch.isfinite(out).all(), f"axis {axis}: read outside the input"
        o = 0
        got = out.cpu().numpy()
        for T in frames:
            m = x[o:o + rows * T].reshape(rows, T)
            assert np.array_equal(got[o:o + rows * T].reshape(rows, T), lr.median_filter_1d(m, k, axis=axis))
            o += rows * T


def test_moments_uniform_more_than_four_classes(ctx):
    """moments_uniform_kernel with 6 classes (the 8-class instantiation)."""
    rng = np.random.default_rng(77)
    T, D, n, nc = 98, 64, 50, 6
    fvs = [(rng.standard_normal((D, T)) * 4 - 20).astype(np.float32) for _ in range(n)]
    cls = [int(c) for c in rng.integers(0, nc, size=n)]
    batch = engine.Batch(ctx, clip_frames=[T] * n)
    acc = engine.moments(batch, to_dev(flat_batch(fvs)), D, cls, nc).cpu().numpy()
    X = np.stack(fvs).astype(np.float64)
    for k in range(nc):
        idx = [i for i in range(n) if cls[i] == k]
        want = X[idx].sum(axis=(0, 2)) if idx else np.zeros(D)
        assert np.allclose(acc[k * D:(k + 1) * D], want, rtol=1e-6, atol=1e-3)
    assert np.allclose(acc[nc * D:(nc + 1) * D], (X ** 2).sum(axis=(0, 2)), rtol=1e-6)


@pytest.mark.parametrize("T,D,n", [(98, 240, 37), (128, 80, 9), (5, 402, 64)])
def test_moments_uniform_short_clips(ctx, T, D, n):
    """Batches of equal short clips take moments_uniform_kernel (constant strides, no table lookups): raw
    moments against float64 numpy, and the fused top_db clip against hpss_topdb_clip."""
    rng = np.random.default_rng(T * 1000 + D)
    fvs = [(rng.standard_normal((D, T)) * 9 - 35).astype(np.float32) for _ in range(n)]
    cls = [int(c) for c in rng.integers(0, 3, size=n)]
    batch = engine.Batch(ctx, clip_frames=[T] * n)
    feat = to_dev(flat_batch(fvs))
    acc = engine.moments(batch, feat, D, cls, 3).cpu().numpy()
    X = np.stack(fvs).astype(np.float64)                       # (n, D, T)
    want_sum = np.stack([X[[i for i in range(n) if cls[i] == k]].sum(axis=(0, 2)) if any(c == k for c in cls)
                         else np.zeros(D) for k in range(3)])
    assert np.allclose(acc[:3 * D].reshape(3, D), want_sum, rtol=1e-6, atol=1e-3)
    assert np.allclose(acc[3 * D:4 * D], (X ** 2).sum(axis=(0, 2)), rtol=1e-6)
    assert list(acc[4 * D:4 * D + 3]) == [T * sum(1 for c in cls if c == k) for k in range(3)]
    assert acc[4 * D + 3] == 0
    # fused clip: two streams of D/2 rows, per-clip / per-stream maxima
    rps = D // 2
    mx = X.reshape(n, 2, rps * T).max(axis=2).astype(np.float32)
    key = np.where(mx.view(np.uint32) & 0x80000000, ~mx.view(np.uint32), mx.view(np.uint32) | 0x80000000).astype(np.uint32)
    cmax = torch.from_numpy(key.view(np.int32).reshape(-1)).cuda()
    a = feat.clone()
    b_ = feat.clone()
    acc2 = engine.topdb_moments(batch, a, rps, 2, cmax, 30.0, cls, 3).cpu().numpy()
    engine.topdb_clip(batch, b_, rps, 2, cmax, 30.0)
    assert torch.equal(a, b_)
    Xc = b_.cpu().numpy().reshape(n, D, T).astype(np.float64)
    assert np.allclose(acc2[3 * D:4 * D], (Xc ** 2).sum(axis=(0, 2)), rtol=1e-6)


# ------------------------------------------------------------------------------ MFCC extension (not in the reference)
@pytest.mark.parametrize("M,streams,n_mfcc,Ts", [(120, 2, 20, [98, 7, 300, 33]), (21, 1, 13, [64]), (128, 2, 40, [31, 1, 2]),
                                                 (120, 2, 64, [50]), (16, 3, 16, [40, 90])])
def test_dct_mfcc_matches_scipy(ctx, M, streams, n_mfcc, Ts):
    import scipy.fft
    rng = np.random.default_rng(M + n_mfcc)
    mats = [(rng.standard_normal((streams * M, T)) * 25 - 40).astype(np.float32) for T in Ts]
    batch = engine.Batch(ctx, clip_frames=Ts)
    out = engine.dct_mfcc(batch, to_dev(flat_batch(mats)), M, streams, n_mfcc)
    torch.cuda.synchronize()
    for c, got in enumerate(batch.split(out, streams * n_mfcc)):
        x = mats[c].astype(np.float64).reshape(streams, M, -1)
        want = scipy.fft.dct(x, axis=1, type=2, norm="ortho")[:, :n_mfcc].reshape(streams * n_mfcc, -1)
        assert np.array_equal(lr.dct_ortho(x[0], n_mfcc), want[:n_mfcc]) or rel_l2(lr.dct_ortho(x[0], n_mfcc), want[:n_mfcc]) < 1e-12
        r = rel_l2(got.cpu().numpy(), want)
        assert r < 1e-6, f"clip {c}: rel-L2 {r:.3e}, max-abs {np.abs(got.cpu().numpy() - want).max():.3e}"


def test_dct_mfcc_argument_errors(ctx):
    batch = engine.Batch(ctx, clip_frames=[10])
    x = torch.zeros(40 * 10, device="cuda")
    with pytest.raises(Exception):
        engine.dct_mfcc(batch, x, 40, 1, 41)
    with pytest.raises(Exception):
        engine.dct_mfcc(batch, x, 40, 1, 20, out=x)          # aliasing


def test_mfcc_of_hpss_featuregram_end_to_end(ctx):
    """waveform -> HPSS log-mel featuregram (reference path) -> MFCC per stream, against the oracle + scipy."""
    import scipy.fft
    y = synth.synth_clip(7, 16000)
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=120)
    batch = engine.Batch(ctx, clip_lengths=[len(y)], n_fft=400, hop_length=160)
    feat = engine.featuregram(batch, to_dev(y.astype(np.float32)), prm)
    got = engine.dct_mfcc(batch, feat, 120, 2, 20).cpu().numpy().reshape(40, -1)
    fv = po.featuregram(y, 16000, 25, 10, 21, 11, 400, 120, "LogMelHarmPercSpec")
    want = scipy.fft.dct(fv.astype(np.float64).reshape(2, 120, -1), axis=1, type=2, norm="ortho")[:, :20].reshape(40, -1)
    assert rel_l2(got, want) < REL_L2_TOL
