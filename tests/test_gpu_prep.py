"""GPU tests of the rows around the hot path: signal preparation on the device (N2: normalise, RMS gate, silence
excision, doubling, SMR mixing, int16 PCM upload), the host pipeline that starts from decoded files, the
device-resident patch tensor (N1), per-patch statistics (N4), non-finite / negative input reporting and the
statistics pickle (N3).  Everything is compared with the golden vectors the reference's own code produced
(tests/golden/make_golden.py) and with the oracle on seeded random inputs."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import preprocessing_oracle as po
from sm_hpss_mtl_b200 import _lib, engine, synth
from sm_hpss_mtl_b200 import preprocessing as pp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_glue.npz")
MODEL = "Lemaire_et_al_MTL"
PARAMS = {"Tw": 25, "Ts": 10, "Model": MODEL, "l_harm": {MODEL: 21}, "l_perc": {MODEL: 11},
          "frame_level_scaling": False}
SIG_TOL = 2e-7      # prepared signals: float32 results of float64 reductions vs numpy's float32 pairwise sums


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def audio(golden):
    return {k[len("audio:"):]: golden[k] for k in golden.files if k.startswith("audio:")}


def gappy_clip(seed, n, n_gaps):
    """Synthetic clip with `n_gaps` near-silent stretches of 80 .. 400 ms."""
    rng = np.random.default_rng(seed)
    x = synth.synth_clip(seed, n).astype(np.float32)
    for _ in range(n_gaps):
        w = int(rng.integers(1300, 6400))
        if n <= w + 800:
            continue
        a = int(rng.integers(400, n - w - 400))
        x[a:a + w] *= np.float32(1e-4)
    return x * np.float32(rng.uniform(0.05, 0.9)) + np.float32(rng.uniform(-0.01, 0.01))


# ------------------------------------------------------------------------------------------------ N2: preparation
def test_prep_matches_reference_golden(ctx, golden, audio):
    """All six golden files in ONE batched call: prepared signals within 2e-7 of what the reference's
    load_and_preprocess_signal returned, frame and sample markers bit-exact against its Cython removeSilence."""
    paths = list(audio)
    out, lens, fm, sm, ns = pp.prepare_signals_device(ctx, [audio[p] for p in paths], 25, 10, markers=True)
    engine.ctx_check(ctx)
    out, fm, sm, ns = out.cpu().numpy(), fm.cpu().numpy(), sm.cpu().numpy(), ns.cpu().numpy()
    o = f = s = 0
    for i, p in enumerate(paths):
        want = golden["prep:" + p]
        assert lens[i] == want.size
        got = out[o:o + lens[i]]
        assert np.abs(got - want).max() <= SIG_TOL, (p, float(np.abs(got - want).max()))
        nf = golden["gate:frame:" + p].size
        assert np.array_equal(fm[f:f + nf], golden["gate:frame:" + p]), p
        n = audio[p].size
        assert np.array_equal(sm[s:s + n], golden["gate:sample:" + p]), p
        o += lens[i]; f += nf; s += n
    # sp0 and mu1 have two silent stretches (removed), onesil one (marked, not removed), short is doubled
    assert list(ns) == [2, 0, 2, 0, 0, 1]


def test_prep_reference_signature_and_mix(ctx, golden, audio):
    for p in audio:
        got, fs = pp.load_and_preprocess_signal(p, 25, 10, loader=lambda q: audio[q].copy())
        assert fs == 16000 and got.dtype == np.float32
        assert np.abs(got - golden["prep:" + p]).max() <= SIG_TOL
    mix = pp.mix_signals(golden["prep:/d/speech/sp0.wav"], golden["prep:/d/music/mu0.wav"], 5)
    assert mix.dtype == np.float32 and np.abs(mix - golden["mix:sp0+mu0@5"]).max() <= SIG_TOL
    # music shorter than the speech (looped) and longer (cut), several ratios, against the oracle
    rng = np.random.default_rng(3)
    for n_sp, n_mu, db in [(20000, 7000, -5), (9000, 30000, 20), (16000, 16000, 0), (12345, 1000, 10)]:
        sp = po.normalize_signal(rng.standard_normal(n_sp).astype(np.float32))
        mu = po.normalize_signal(rng.standard_normal(n_mu).astype(np.float32))
        assert np.abs(pp.mix_signals(sp, mu, db) - po.mix_signals(sp, mu, db)).max() <= SIG_TOL


def test_prep_int16_pcm_equals_float_input(ctx):
    """16-bit PCM uploaded as int16 (half the bytes) gives bit-identical results to librosa.load's x / 32768."""
    rng = np.random.default_rng(11)
    clips = [(gappy_clip(50 + i, n, g) * 30000).astype(np.int16) for i, (n, g) in
             enumerate([(16000, 3), (40001, 5), (1700, 0), (900, 0), (23456, 2)])]
    a, la = pp.prepare_signals_device(ctx, clips, 25, 10)
    b, lb = pp.prepare_signals_device(ctx, [c.astype(np.float32) / np.float32(32768.0) for c in clips], 25, 10)
    assert la == lb and torch.equal(a, b)
    assert la[3] == 1800                       # 900 samples = 0.056 s: doubled once


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_prep_random_ragged_batch_vs_oracle(ctx, seed):
    """Ragged batch of clips with random silent stretches (0 .. 6 per clip, incl. stretches touching the clip
    ends): markers bit-exact, signals within 2e-7 of the oracle (which is pinned to the reference above)."""
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(1000, 90000, size=24)] + [160, 400, 401, 8193, 16384]
    clips = [gappy_clip(1000 * seed + i, n, int(rng.integers(0, 7))) for i, n in enumerate(lens)]
    clips[0][:3000] *= np.float32(1e-5)        # silence at the very start
    clips[1][-3000:] *= np.float32(1e-5)       # ... and at the very end (the j == nFrames-1 quirk)
    out, olen, fm, sm, ns = pp.prepare_signals_device(ctx, clips, 25, 10, markers=True)
    engine.ctx_check(ctx)
    out, fm, sm = out.cpu().numpy(), fm.cpu().numpy(), sm.cpu().numpy()
    o = f = s = 0
    n_applied = 0
    for i, x in enumerate(clips):
        want, smark, fmark, _ = po.load_and_preprocess_signal(x, 25, 10, details=True)
        assert olen[i] == want.size
        assert np.array_equal(fm[f:f + fmark.size], fmark), i
        assert np.array_equal(sm[s:s + x.size], smark), i
        err = float(np.abs(out[o:o + want.size] - want).max())
        assert err <= SIG_TOL, (i, x.size, err)
        n_applied += int((smark == 0).any())
        o += want.size; f += fmark.size; s += x.size
    assert n_applied >= 5                       # the test really exercises the excision


def test_prep_long_clip_many_chunks(ctx):
    """One 5-minute clip (586 chunks of 8192 samples, 30 001 RMS frames: the gate kernel loops) next to a short one."""
    x = gappy_clip(77, 4_800_000, 60)
    y = gappy_clip(78, 20000, 2)
    out, olen, fm, sm, ns = pp.prepare_signals_device(ctx, [x, y], 25, 10, markers=True)
    out, fm, sm = out.cpu().numpy(), fm.cpu().numpy(), sm.cpu().numpy()
    want, smark, fmark, _ = po.load_and_preprocess_signal(x, 25, 10, details=True)
    assert np.array_equal(fm[:fmark.size], fmark) and np.array_equal(sm[:x.size], smark)
    assert np.abs(out[:x.size] - want).max() <= SIG_TOL
    want2 = po.load_and_preprocess_signal(y, 25, 10)
    assert np.abs(out[x.size:] - want2).max() <= SIG_TOL


@pytest.mark.parametrize("Tw,Ts", [(25, 10), (20, 10), (32, 8), (64, 16)])
def test_prep_other_frame_sizes_and_edge_clips(ctx, Tw, Ts):
    """Frame sizes other than 25 / 10 ms (win 320 / 512 / 1024, hop 128 / 160 / 256), clips of 2 .. a few hundred samples
    (shorter than one frame: the reflect padding wraps several times; all below 0.1 s: doubled 1 .. 10 times), a
    clip that is silent except for one burst, and a constant clip (peak 0: the reference divides by zero -> NaN)."""
    rng = np.random.default_rng(Tw * 100 + Ts)
    clips = [gappy_clip(900 + i, n, g) for i, (n, g) in enumerate([(30000, 4), (16000, 2), (50000, 6)])]
    clips += [rng.standard_normal(n).astype(np.float32) for n in (2, 3, 17, 161, 399, 1599)]
    burst = np.full(24000, 1e-6, np.float32) * rng.standard_normal(24000).astype(np.float32)
    burst[9000:9800] = rng.standard_normal(800).astype(np.float32)
    clips.append(burst)
    out, olen, fm, sm, ns = pp.prepare_signals_device(ctx, clips, Tw, Ts, markers=True)
    engine.ctx_check(ctx)
    out, fm, sm = out.cpu().numpy(), fm.cpu().numpy(), sm.cpu().numpy()
    o = f = s = 0
    for i, x in enumerate(clips):
        want, smark, fmark, _ = po.load_and_preprocess_signal(x, Tw, Ts, details=True)
        assert olen[i] == want.size and want.size >= 1600
        assert np.array_equal(fm[f:f + fmark.size], fmark), (i, x.size)
        assert np.array_equal(sm[s:s + x.size], smark), (i, x.size)
        # a mean over 2 .. 17 float32 samples carries a relative rounding error of ~6e-8 in numpy, which the two
        # normalisations turn into a few ulp of the result: 1e-6 for those, the usual 2e-7 otherwise
        tol = SIG_TOL if x.size >= 100 else 1e-6
        assert np.abs(out[o:o + want.size] - want).max() <= tol, (i, x.size)
        o += want.size; f += fmark.size; s += x.size
    # constant clip: mean-subtracted signal is all zero, peak 0 -> 0 / 0 like numpy (NaN everywhere), reported by the features
    const = np.full(4000, 0.25, np.float32)
    got, _ = pp.prepare_signals_device(ctx, [const], Tw, Ts)
    with np.errstate(all="ignore"):
        want = po.load_and_preprocess_signal(const, Tw, Ts)
    assert np.isnan(want).all() and bool(torch.isnan(got).all())


def test_pipeline_ragged_prepare_features_and_errors(ctx):
    """hpss_pipeline with ragged decoded files, preparation, features AND moments in one run; argument checking;
    non-finite PCM is reported by the run itself."""
    lens = [1000, 52000, 16000, 700, 33333, 16000, 120000, 2500]
    clips = [gappy_clip(700 + i, n, 2) for i, n in enumerate(lens)]
    cls = [i % 2 for i in range(len(lens))]
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=40)
    pl = engine.Pipeline(ctx, lens, prm, pcm_dtype=np.float32, prepare=True, n_chunks=3)
    pcm = np.concatenate(clips)
    feat, mom = pl.run(pcm, clip_class=cls, n_classes=2)
    D = pl.rows
    for c, x in enumerate(clips):
        y = po.load_and_preprocess_signal(x, 25, 10)
        want = po.featuregram(y, 16000, 25, 10, 21, 11, 400, 40, "LogMelHarmPercSpec")
        a, b = D * int(pl.frame_offsets[c]), D * int(pl.frame_offsets[c + 1])
        got = feat[a:b].reshape(D, -1)
        assert got.shape == want.shape and rel_l2(got, want) < 1e-4, c
    counts = mom[2 * D + D:2 * D + D + 2]
    assert [int(v) for v in counts] == [int(sum(pl.frame_offsets[c + 1] - pl.frame_offsets[c] for c in range(len(lens)) if cls[c] == k))
                                        for k in (0, 1)]
    with pytest.raises(ValueError):
        pl.run(pcm[:-1])
    with pytest.raises(ValueError):
        pl.run(pcm, clip_class=cls[:-1], n_classes=2)
    with pytest.raises(_lib.HpssError):
        pl.run(pcm, clip_class=[5] * len(lens), n_classes=2)              # class outside [0, n_classes)
    bad = pcm.copy()
    bad[60000] = np.inf
    with pytest.raises(_lib.ParameterError, match="not finite"):
        pl.run(bad)
    pl.run(pcm)                                                           # the status word was cleared
    pl.close()
    with pytest.raises(_lib.ParameterError):                              # n_fft larger than a (prepared) clip: like librosa.stft
        engine.Pipeline(ctx, [100], engine.make_params(n_fft=2048, win_length=2048, hop_length=512), prepare=True)
    with pytest.raises(_lib.HpssError):                                   # 16-bit PCM needs the preparation
        engine.Pipeline(ctx, lens, prm, pcm_dtype=np.int16, prepare=False)


def test_nonfinite_audio_is_reported(ctx, audio):
    x = audio["/d/speech/sp1.wav"].copy()
    x[5000] = np.nan
    with pytest.raises(_lib.ParameterError, match="not finite"):
        pp.load_and_preprocess_signal("x", 25, 10, loader=lambda q: x)
    with pytest.raises(_lib.ParameterError, match="not finite"):          # host pipeline, prepared waveform
        pp.featuregram_batch([x], 16000, PARAMS, 400, 40, "LogMelHarmPercSpec")
    from sm_hpss_mtl_b200 import librosa_compat as lc
    with pytest.raises(_lib.ParameterError, match="not finite"):
        lc.stft(x, n_fft=400, hop_length=160, win_length=400)
    # the status word was cleared: the next clean call passes
    pp.featuregram_batch([audio["/d/speech/sp1.wav"]], 16000, PARAMS, 400, 40, "LogMelHarmPercSpec")
    # direct C-ABI callers: validate + check
    w = torch.from_numpy(x).cuda()
    engine.validate_audio(ctx, w)
    with pytest.raises(_lib.ParameterError):
        engine.ctx_check(ctx)
    engine.ctx_check(ctx)


def test_negative_spectrogram_is_reported(ctx):
    rng = np.random.default_rng(0)
    S = np.abs(rng.standard_normal((201, 50))).astype(np.float32)
    S[17, 3] = -1e-3
    with pytest.raises(_lib.ParameterError, match="non-negative"):
        pp.featuregram_from_spec(S, PARAMS, 40, "LogMelHarmPercSpec")
    from sm_hpss_mtl_b200 import librosa_compat as lc
    with pytest.raises(_lib.ParameterError, match="non-negative"):
        lc.hpss(S, kernel_size=(21, 11))
    pp.featuregram_from_spec(np.abs(S), PARAMS, 40, "LogMelHarmPercSpec")


# ------------------------------------------------------------------------------------------------ host pipeline
@pytest.mark.parametrize("dtype", [np.float32, np.int16])
def test_pipeline_from_decoded_files(ctx, dtype):
    """hpss_pipeline_run with prepare: decoded files in host memory -> features + raw moments, against the
    staged device calls (prep_signals -> featuregram_moments) on the same data; several chunks."""
    lens = [16000, 30000, 1000, 48000, 20000, 16000, 9000, 64000]
    clips = [gappy_clip(300 + i, n, 3) for i, n in enumerate(lens)]
    if dtype == np.int16:
        clips = [(c * 30000).astype(np.int16) for c in clips]
    cls = [i % 3 for i in range(len(lens))]
    prm = engine.make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=40)
    pl = engine.Pipeline(ctx, lens, prm, pcm_dtype=dtype, prepare=True, n_chunks=4)
    assert pl.n_chunks >= 1
    pcm = engine.host_alloc(sum(lens), dtype)
    pcm[:] = np.concatenate(clips)
    feat, mom = pl.run(pcm, clip_class=cls, n_classes=3)
    # staged reference on the device
    wave, olen = engine.prep_signals(ctx, torch.from_numpy(np.concatenate(clips)).cuda(), lens)
    batch = engine.Batch(ctx, clip_lengths=olen, n_fft=400, hop_length=160)
    out, acc = engine.featuregram_moments(batch, wave, prm, cls, 3)
    assert np.array_equal(feat, out.cpu().numpy())
    assert np.allclose(mom, acc.cpu().numpy(), rtol=1e-12, atol=1e-9)
    assert np.array_equal(pl.frame_offsets, batch.frame_offsets)
    # moments only (the get_data_stats shape: nothing but 8 KB comes back), accumulated into the same vector
    _, mom2 = pl.run(pcm, clip_class=cls, n_classes=3, moments=mom.copy(), want_features=False)
    assert np.allclose(mom2, 2 * acc.cpu().numpy(), rtol=1e-12, atol=1e-9)
    pl.close()


def test_get_featuregram_runs_prep_on_gpu(ctx, golden, audio, tmp_path, monkeypatch):
    """get_featuregram through the reference signature: the signal preparation goes through hpss_prep_signals
    (counted), and the 5-class variant with the noise argument works."""
    calls = {"n": 0}
    real = engine.prep_signals

    def counted(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    monkeypatch.setattr(engine, "prep_signals", counted)
    loader = lambda q: audio[q].copy()
    got = pp.get_featuregram(PARAMS, "speech", str(tmp_path), "/d/speech/sp0.wav", "", -1, 400, 40, "LogMelHarmPercSpec",
                             save_feat=False, loader=loader)
    assert calls["n"] == 1 and rel_l2(got, golden["fv:speech:LogMelHarmPercSpec"]) < 1e-4
    P31 = dict(PARAMS, l_harm={MODEL: 31}, l_perc={MODEL: 31})
    got = pp.get_featuregram(P31, "speech_music", str(tmp_path), "/d/speech/sp1.wav", "/d/music/mu1.wav", -5, 400, 120,
                             "LogMelHarmPercSpec", save_feat=False, loader=loader)
    want = golden["fv120:speech_music:LogMelHarmPercSpec"]               # headline shape: 120 mels, k = 31 / 31
    assert got.shape == want.shape == (240, 118) and rel_l2(got, want) < 1e-4
    got = pp.get_featuregram(P31, "speech", str(tmp_path), "/d/speech/sp1.wav", "", -1, 400, 120, "LogMelHarmPercSpec",
                             save_feat=False, loader=loader)
    assert rel_l2(got, golden["fv120:speech:LogMelHarmPercSpec"]) < 1e-4
    # 5-class copy (5_class_classification.py:314): noise file as the third path
    a = pp.get_featuregram_5class(PARAMS, "noise", str(tmp_path), "", "", "/d/music/mu1.wav", -1, 400, 40,
                                  "LogMelHarmPercSpec", loader=loader)
    assert rel_l2(a, golden["fv:music:LogMelHarmPercSpec"]) < 1e-4     # same file, same features, other class name
    assert os.path.exists(os.path.join(str(tmp_path), "noise", "mu1.npy"))
    b = pp.get_featuregram_5class(PARAMS, "speech_noise", str(tmp_path), "/d/speech/sp0.wav", "", "/d/music/mu0.wav", 5,
                                  400, 40, "LogMelHarmPercSpec", loader=loader)
    assert rel_l2(b, golden["fv:speech_music:LogMelHarmPercSpec"]) < 1e-4
    assert os.path.exists(os.path.join(str(tmp_path), "speech_noise", "sp0_mu0_5dB.npy"))
    with pytest.raises(ValueError):
        pp.get_featuregram_5class(PARAMS, "noise", str(tmp_path), "", "", "/d/music/mu1.wav", -1, 400, 40, "Spec", loader=loader)


def _stats_setup(golden, audio, tmp_path, classes):
    P = dict(PARAMS, classes=classes, feature_opDir=str(tmp_path), folder="/d",
             featName={MODEL: "LogMelHarmPercSpec"}, n_fft={MODEL: 400}, n_mels={MODEL: 40})
    files = {"music": ["mu0.wav", "mu1.wav"], "speech": ["sp0.wav", "sp1.wav"],
             "speech+music": [{"speech": "sp0.wav", "music": "mu0.wav", "SMR": 5},
                              {"speech": "sp1.wav", "music": "mu1.wav", "SMR": -5}]}
    return P, files, (lambda q: audio[q].copy())


def test_get_data_stats_from_files_matches_reference(ctx, golden, audio, tmp_path):
    """get_data_stats from decoded files (PCM upload -> prep -> features -> moments on the device, caches written),
    then again from the caches: the reference's (mean, stdev, nMu, nSp, nSpMu)."""
    P, files, loader = _stats_setup(golden, audio, tmp_path, {0: "music", 1: "speech", 2: "speech_music"})
    for _ in range(2):
        mean, std, nMu, nSp, nSpMu = pp.get_data_stats(P, files, loader=loader)
        assert [nMu, nSp, nSpMu] == list(golden["stats:counts"])
        assert mean.dtype == np.float32 and np.allclose(mean, golden["stats:mean"], rtol=2e-5, atol=2e-5)
        assert np.allclose(std, golden["stats:std"], rtol=2e-5, atol=2e-5)
    for cls in ("music", "speech", "speech_music"):
        for f in os.listdir(os.path.join(str(tmp_path), cls)):
            assert rel_l2(np.load(os.path.join(str(tmp_path), cls, f)), golden[f"statsfv:{cls}:{f[:-4]}"]) < 1e-4
    # the statistics pickle of Baseline_Results.py:609-621
    st = pp.load_or_compute_data_stats(P, files, 0, loader=loader)
    path = os.path.join(str(tmp_path), "data_stats_fold0_3class_train.pkl")
    assert os.path.exists(path)
    with open(path, "rb") as f:
        raw = pickle.load(f)
    assert set(raw) == {"mean", "stdev", "nFrames"} and raw["nFrames"] == list(golden["stats:counts"])
    assert np.array_equal(raw["mean"], st["mean"]) and np.array_equal(pp.load_or_compute_data_stats(P, files, 0)["stdev"], st["stdev"])


def test_get_data_stats_two_classes_returns_five_values(ctx, golden, audio, tmp_path):
    """2-class configurations ({0:'music', 1:'speech'}: B3 tuning, DAFx12, t-SNE scripts) still unpack five values;
    a reordered classes dict does not permute the counts."""
    P, files, loader = _stats_setup(golden, audio, tmp_path, {0: "speech", 1: "music"})
    mean, std, nMu, nSp, nSpMu = pp.get_data_stats(P, files, loader=loader)
    assert (nMu, nSp, nSpMu) == (int(golden["stats:counts"][0]), int(golden["stats:counts"][1]), 0)
    fv = {n: [golden[k] for k in golden.files if k.startswith(f"statsfv:{n}:")] for n in ("music", "speech")}
    want_mean, want_std, n0, n1 = po.get_data_stats(fv, ["music", "speech"])
    assert np.allclose(mean, want_mean, rtol=2e-5, atol=2e-5) and np.allclose(std, want_std, rtol=2e-5, atol=2e-5)


def test_data_stats_drops_nonfinite_rows_like_the_reference(ctx):
    """lib/preprocessing.py:507-508 drops, per file, feature rows holding a NaN / Inf: the same rows in every file ->
    statistics of the remaining rows; different rows -> the reference's np.add fails, so do we."""
    rng = np.random.default_rng(5)
    D = 12
    fvs = {n: [(rng.standard_normal((D, int(rng.integers(30, 80)))) * 4 - 20).astype(np.float32) for _ in range(3)]
           for n in ("music", "speech")}
    for n in fvs:
        for fv in fvs[n]:
            fv[3, int(rng.integers(0, fv.shape[1]))] = np.nan
            fv[7, 0] = np.inf
    mean, std, n0, n1 = pp.data_stats_from_featuregrams(fvs, ["music", "speech"])
    want_mean, want_std, m0, m1 = po.get_data_stats(fvs, ["music", "speech"])
    assert mean.shape == (D - 2,) and (n0, n1) == (m0, m1)
    assert np.allclose(mean, want_mean, rtol=1e-6, atol=1e-6) and np.allclose(std, want_std, rtol=1e-6, atol=1e-6)
    fvs["speech"][1][5, 2] = np.nan                      # now one file drops a different set of rows
    with pytest.raises(ValueError, match="broadcast"):
        pp.data_stats_from_featuregrams(fvs, ["music", "speech"])
    with pytest.raises(ValueError):
        po.get_data_stats(fvs, ["music", "speech"])


# ------------------------------------------------------------------------------------------------ N1 / N4
@pytest.mark.parametrize("model", ["Lemaire_et_al_MTL", "Doukhan_et_al_MTL"])
@pytest.mark.parametrize("fn", ["LogMelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec"])
def test_patch_tensor_device_resident(ctx, golden, model, fn):
    """Batch of featuregrams on the device -> model-ready tensor in one call: equals get_feature_patches of the
    reference per file (golden), transposed for the TCN models, incl. a clip shorter than the patch (tiled)."""
    fvs = [golden["fv:speech:LogMelHarmPercSpec"], golden["fv:music:LogMelHarmPercSpec"],
           golden["fv:speech_music:LogMelHarmPercSpec"][:, :40]]
    P = dict(PARAMS, Model=model)
    for (W, sh) in [(68, 68), (49, 24)]:
        batch = engine.Batch(ctx, clip_frames=[fv.shape[1] for fv in fvs])
        feat = torch.from_numpy(np.concatenate([fv.ravel() for fv in fvs])).cuda()
        out = pp.feature_patches_device(P, batch, feat, 80, W, sh, fn, dtype=torch.float64)
        off = engine.patch_offsets(batch, W, sh)
        assert out.shape[0] == off[-1]
        for c, fv in enumerate(fvs):
            want = po.get_feature_patches(fv, W, sh, fn, model)
            got = out[off[c]:off[c + 1]].cpu().numpy()
            if "Lemaire" in model:
                got = np.transpose(got, (0, 2, 1))                       # (n, W, nFeat) -> (n, nFeat, W)
            assert got.shape == want.shape, (c, got.shape, want.shape)
            assert np.allclose(got, want, rtol=0, atol=2e-6)
        # float32 output (what the network consumes) = the float64 one rounded
        feat = torch.from_numpy(np.concatenate([fv.ravel() for fv in fvs])).cuda()
        out32 = pp.feature_patches_device(P, batch, feat, 80, W, sh, fn)
        assert out32.dtype == torch.float32 and torch.equal(out32, out.to(torch.float32))
        # DLPack hand-over (zero copy)
        again = torch.utils.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(out32))
        assert again.data_ptr() == out32.data_ptr()
    first = golden[f"patch:{model}:{fn}:68:68"]
    assert np.allclose(pp.get_feature_patches(P, fvs[0], 68, 68, fn), first, rtol=0, atol=2e-6)


def test_patch_statistics_match_reference(ctx, golden):
    pat = golden["patch:Lemaire_et_al_MTL:LogMelHarmPercSpec:49:24"]
    for st in ("mean", "variance", "skew", "kurtosis"):
        for ax in (0, 1):
            got = pp.get_data_statistics(pat, st, ax)
            want = golden[f"pstat:{st}:{ax}"]
            assert got.shape == want.shape and got.dtype == np.float64
            assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (st, ax, float(np.abs(got - want).max()))
    cnn = golden["patch:Doukhan_et_al_MTL:LogMelHarmPercSpec:49:24"]       # (n, f, t, 1): squeezed like np.squeeze
    assert np.allclose(pp.get_data_statistics(cnn, "skew", 1), golden["pstat:skew:1"], rtol=1e-9, atol=1e-9)


def test_workspace_is_safe_across_streams(ctx):
    """Two torch streams call hpss_featuregram on the same context at the same time: the context's scratch is handed
    over through an event, so both results equal the single-stream result."""
    Ls = [16000] * 64
    prm = engine.make_params(n_mels=40)
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=400, hop_length=160)
    w1 = torch.from_numpy(synth.synth_batch_fast(64, 16000, first_index=1).ravel()).cuda()
    w2 = torch.from_numpy(synth.synth_batch_fast(64, 16000, first_index=99).ravel()).cuda()
    r1, r2 = engine.featuregram(batch, w1, prm).clone(), engine.featuregram(batch, w2, prm).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(5):
        with torch.cuda.stream(s1):
            a = engine.featuregram(batch, w1, prm)
        with torch.cuda.stream(s2):
            b = engine.featuregram(batch, w2, prm)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, r1) and torch.equal(b, r2)


def test_long_stream_time_split_equals_whole(ctx):
    """configs[3] across ranks: a stream cut along time with l_harm // 2 halo frames per side and a MAX all-reduce of
    the two per-stream maxima gives, shard by shard, exactly the columns of the unsharded featuregram."""
    from sm_hpss_mtl_b200 import dist as hd
    n_fft, hop, k = 2048, 512, 31
    L = 16000 * 40
    wave = torch.from_numpy(synth.synth_clip(5, L)).cuda()
    prm = engine.make_params(n_fft=n_fft, win_length=n_fft, hop_length=hop, l_harm=k, l_perc=k, n_mels=120)
    batch = engine.Batch(ctx, clip_lengths=[L], n_fft=n_fft, hop_length=hop)
    whole = engine.featuregram(batch, wave, prm).view(240, -1)
    for world in (2, 3, 8):
        shards = [hd.stream_shard(L, n_fft, hop, k, r, world) for r in range(world)]
        assert shards[0][0][0] == 0 and shards[-1][0][1] == whole.shape[1]
        # pass 1: every "rank" publishes its maxima; pass 2 applies the stream-wide maximum (what the all-reduce does)
        maxima = []
        for sh in shards:
            hd.featuregram_stream_sharded(ctx, wave[sh[2][0]:sh[2][1]], sh, prm, allreduce_max=lambda m: maxima.append(m.clone()))
        gmax = torch.stack(maxima).amax(dim=0)
        for sh in shards:
            got = hd.featuregram_stream_sharded(ctx, wave[sh[2][0]:sh[2][1]], sh, prm, allreduce_max=lambda m: m.copy_(gmax))
            assert torch.equal(got, whole[:, sh[0][0]:sh[0][1]]), (world, sh[0])
