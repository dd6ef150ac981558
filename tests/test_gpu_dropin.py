"""GPU tests of the drop-in layer: the reference-signature functions of sm_hpss_mtl_b200.preprocessing
against the golden vectors produced by the reference's own lib/preprocessing.py, the librosa-shaped
helpers against the oracle, and size-independent properties at the full benchmark size."""
import os

import numpy as np
import pytest
import torch

from oracle import librosa_restated as lr
from oracle import preprocessing_oracle as po
from sm_hpss_mtl_b200 import engine, synth
from sm_hpss_mtl_b200 import librosa_compat as lc
from sm_hpss_mtl_b200 import preprocessing as pp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_glue.npz")
TOL = 1e-4
MODEL = "Lemaire_et_al_MTL"
PARAMS = {"Tw": 25, "Ts": 10, "Model": MODEL, "l_harm": {MODEL: 21}, "l_perc": {MODEL: 11},
          "frame_level_scaling": False}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def loader(golden):
    audio = {k[len("audio:"):]: golden[k] for k in golden.files if k.startswith("audio:")}
    return lambda path: audio[path].copy()


FEATS = ["Spec", "LogSpec", "MelSpec", "LogMelSpec", "PercSpec", "HarmPercSpec", "LogHarmPercSpec",
         "MelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec", "LogMelHarmPercSpec"]


@pytest.mark.parametrize("fn", FEATS)
def test_get_featuregram_matches_reference(ctx, golden, loader, fn, tmp_path):
    """File in -> featuregram out through the reference's own signature (load, normalise, silence
    removal, [mix], features), compared with what the reference's get_featuregram returned."""
    n_fft, n_mels = (512, 21) if fn == "LogHarmPercSpec" else (400, 40)
    got = pp.get_featuregram(PARAMS, "speech", str(tmp_path), "/d/speech/sp0.wav", "", -1, n_fft, n_mels, fn,
                             save_feat=True, loader=loader)
    want = golden[f"fv:speech:{fn}"]
    assert got.dtype == np.float32 and got.shape == want.shape
    assert rel_l2(got, want) < TOL, f"{fn}: rel-L2 {rel_l2(got, want):.2e} max-abs {np.abs(got - want).max():.2e}"
    # cache: same directory scheme as the reference, second call returns the stored array
    assert os.path.exists(os.path.join(str(tmp_path), "speech", "sp0.npy"))
    again = pp.get_featuregram(PARAMS, "speech", str(tmp_path), "/d/speech/sp0.wav", "", -1, n_fft, n_mels, fn,
                               loader=None)
    assert np.array_equal(again, got)
    got = pp.get_featuregram(PARAMS, "speech_music", str(tmp_path), "/d/speech/sp0.wav", "/d/music/mu0.wav", 5, n_fft,
                             n_mels, fn, save_feat=False, loader=loader)
    want = golden[f"fv:speech_music:{fn}"]
    assert got.shape == want.shape and rel_l2(got, want) < TOL


@pytest.mark.parametrize("model", ["Lemaire_et_al_MTL", "Doukhan_et_al_MTL"])
@pytest.mark.parametrize("fn", ["LogMelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec"])
@pytest.mark.parametrize("W,sh", [(68, 68), (49, 24), (249, 24)])
def test_get_feature_patches_matches_reference(ctx, golden, model, fn, W, sh):
    FV = golden["fv:speech:LogMelHarmPercSpec"]
    before = FV.copy()
    want = golden[f"patch:{model}:{fn}:{W}:{sh}"]
    got = pp.get_feature_patches(dict(PARAMS, Model=model), FV, W, sh, fn)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.allclose(got, want, rtol=0, atol=2e-6), float(np.abs(got - want).max())
    assert np.array_equal(FV, before)        # unlike the reference we do not mutate the caller's array


def test_get_feature_patches_dafx_variant(ctx, golden):
    """DAFx12 copy: patches of an already standardised featuregram, float32, stride 1."""
    FV = po._standard_scale_rows(golden["fv:speech:LogMelHarmPercSpec"])
    got = pp.get_feature_patches_dafx(dict(PARAMS, Model="Lemaire_et_al_MTL"), FV, 99, 1, "LogMelHarmPercSpec")
    want = po.extract_patches(FV, 99, 1).astype(np.float32)
    assert got.dtype == np.float32 and got.shape == want.shape == (FV.shape[1] - 98, 80, 99) and np.array_equal(got, want)


def test_get_feature_patches_spec_and_fls(ctx, golden):
    got = pp.get_feature_patches(dict(PARAMS, Model="Doukhan_et_al_MTL"), golden["fv:speech:Spec"], 21, 21, "Spec")
    assert np.allclose(got, golden["patch:Doukhan_et_al_MTL:Spec:21:21"], rtol=0, atol=4e-6)
    P = dict(PARAMS, frame_level_scaling=True)
    got = pp.get_feature_patches(P, golden["fv:speech:LogMelHarmPercSpec"], 49, 24, "LogMelHarmPercSpec")
    assert np.array_equal(got, golden["patch:fls:LogMelHarmPercSpec:49:24"])


def test_get_data_stats_matches_reference(ctx, golden, loader, tmp_path):
    P = dict(PARAMS, classes={0: "music", 1: "speech", 2: "speech_music"}, feature_opDir=str(tmp_path), folder="/d",
             featName={MODEL: "LogMelHarmPercSpec"}, n_fft={MODEL: 400}, n_mels={MODEL: 40})
    files = {"music": ["mu0.wav", "mu1.wav"], "speech": ["sp0.wav", "sp1.wav"],
             "speech+music": [{"speech": "sp0.wav", "music": "mu0.wav", "SMR": 5},
                              {"speech": "sp1.wav", "music": "mu1.wav", "SMR": -5}]}
    mean, std, nMu, nSp, nSpMu = pp.get_data_stats(P, files, loader=loader)
    assert [nMu, nSp, nSpMu] == list(golden["stats:counts"])
    assert mean.dtype == np.float32 and std.dtype == np.float32
    assert np.allclose(mean, golden["stats:mean"], rtol=1e-4, atol=1e-3)
    assert np.allclose(std, golden["stats:std"], rtol=1e-4, atol=1e-3)
    FV = golden["fv:speech:LogMelHarmPercSpec"]
    assert np.allclose(pp.cscale_data(FV, golden["stats:mean"], golden["stats:std"]), golden["scaled:cy"], rtol=1e-12)
    py = pp.scale_data(FV, golden["stats:mean"], golden["stats:std"])
    assert py.dtype == np.float32 and np.array_equal(py, golden["scaled:py"])      # numpy's float32 evaluation, bit for bit
    # frame_level_scaling: the generators hand the float64 output of the Cython scale_data to get_feature_patches
    P = dict(PARAMS, Model="Lemaire_et_al_MTL", frame_level_scaling=True)
    cy = golden["scaled:cy"]
    got = pp.get_feature_patches(P, cy, 49, 24, "LogMelHarmPercSpec")
    assert got.dtype == np.float64 and np.array_equal(got, po.extract_patches(cy, 49, 24))


def test_librosa_compat_surface(ctx):
    y = synth.synth_clip(77, 24000)
    D = lc.stft(y, n_fft=400, hop_length=160, win_length=400)
    want = lr.stft(y, n_fft=400, hop_length=160, win_length=400)
    assert D.dtype == np.complex64 and D.shape == want.shape
    assert np.linalg.norm(D - want) / np.linalg.norm(want) < 1e-5
    S = np.abs(want)
    H, P = lc.hpss(S, kernel_size=(21, 11))
    Hw, Pw = lr.hpss(S, kernel_size=(21, 11))
    assert np.array_equal(H, Hw) and np.array_equal(P, Pw)                 # bit-exact given the same S
    m = lc.melspectrogram(S=Hw, n_mels=120)
    assert rel_l2(m, lr.melspectrogram(S=Hw, n_mels=120)) < 1e-6
    m2 = lc.melspectrogram(y=y, sr=16000, n_fft=400, win_length=400, hop_length=160, center=False, n_mels=40)
    assert rel_l2(m2, lr.melspectrogram(y=y, sr=16000, n_fft=400, win_length=400, hop_length=160, center=False,
                                        n_mels=40)) < 1e-5
    db = lc.power_to_db(m ** 2)
    assert rel_l2(db, lr.power_to_db(m ** 2)) < 1e-6
    assert np.array_equal(lc.mel(22050, 400, 120), lr.mel(22050, 400, 120))
    with pytest.raises(lc.ParameterError):
        lc.stft(y[:300], n_fft=400, hop_length=160, win_length=400)
    with pytest.raises(lc.ParameterError):
        lc.hpss(-S)
    with pytest.raises(lc.ParameterError):
        lc.hpss(S, margin=0.5)


# ------------------------------------------------------------------ full benchmark size (BASELINE.json configs[1])
def test_full_size_batch_properties(ctx):
    """4096 x 1 s clips, k = 31/31: spot clips against the oracle, batch-composition invariance
    (a clip's features do not depend on its neighbours), value ranges, and the top_db identity."""
    n, L = 4096, 16000
    waves = synth.synth_batch_fast(n, L)
    prm = engine.make_params(l_harm=31, l_perc=31, n_mels=120)
    batch = engine.Batch(ctx, clip_lengths=[L] * n, n_fft=400, hop_length=160)
    wave = torch.from_numpy(waves.ravel()).cuda()
    out = engine.featuregram(batch, wave, prm)
    torch.cuda.synchronize()
    fv = out.view(n, 240, 98)
    assert torch.isfinite(fv).all()
    mx = fv.view(n, 2, 120 * 98).amax(dim=2)
    mn = fv.view(n, 2, 120 * 98).amin(dim=2)
    assert (mn >= mx - 80.0 - 1e-3).all()                       # power_to_db top_db per clip and stream
    for c in (0, 1, 2047, 4095):
        want = po.featuregram(waves[c], 16000, 25, 10, 31, 31, 400, 120, "LogMelHarmPercSpec")
        got = fv[c].cpu().numpy()
        assert rel_l2(got, want) < TOL, (c, rel_l2(got, want))
    # the same clips in a different batch (different neighbours, different tile alignment)
    idx = [4095, 17, 2047, 1, 0]
    small = engine.Batch(ctx, clip_lengths=[L] * len(idx), n_fft=400, hop_length=160)
    out2 = engine.featuregram(small, torch.from_numpy(waves[idx].ravel()).cuda(), prm).view(len(idx), 240, 98)
    for j, c in enumerate(idx):
        assert torch.equal(out2[j], fv[c]), c
    # medians at full size: bit-exact on spot clips, idempotent under a second identical launch
    S = engine.stft_mag(batch, wave, 400, 400, 160)
    harm = engine.median_time(batch, S, 201, 31)
    perc = engine.median_freq(batch, S, 201, 31)
    for c in (3, 4000):
        Sc = batch.clip(S, 201, c).cpu().numpy()
        assert np.array_equal(batch.clip(harm, 201, c).cpu().numpy(), lr.median_filter_scipy(Sc, 31, axis=1))
        assert np.array_equal(batch.clip(perc, 201, c).cpu().numpy(), lr.median_filter_scipy(Sc, 31, axis=0))
    assert torch.equal(harm, engine.median_time(batch, S, 201, 31))
    # order statistics stay inside the data range of their line
    Sv = S.view(n, 201, 98)
    assert (harm.view(n, 201, 98) <= Sv.amax(dim=2, keepdim=True)).all()
    assert (perc.view(n, 201, 98) >= Sv.amin(dim=1, keepdim=True)).all()


def test_long_stream_tiling(ctx):
    """One long clip (n_fft 2048, hop 512, k 31): time-axis tiles with halos must agree with the oracle
    at the seams; T is not a multiple of any tile size."""
    L = 2048 + 512 * 1499
    y = synth.synth_clip(5, L)
    batch = engine.Batch(ctx, clip_lengths=[L], n_fft=2048, hop_length=512)
    S = engine.stft_mag(batch, torch.from_numpy(y).cuda(), 2048, 2048, 512)
    Sc = S.view(1025, -1).cpu().numpy()
    assert Sc.shape == (1025, 1500)
    harm = engine.median_time(batch, S, 1025, 31).view(1025, -1).cpu().numpy()
    perc = engine.median_freq(batch, S, 1025, 31).view(1025, -1).cpu().numpy()
    assert np.array_equal(harm, lr.median_filter_scipy(Sc, 31, axis=1))
    assert np.array_equal(perc, lr.median_filter_scipy(Sc, 31, axis=0))
    prm = engine.make_params(n_fft=2048, win_length=2048, hop_length=512, l_harm=31, l_perc=31, n_mels=120)
    got = engine.featuregram(batch, torch.from_numpy(y).cuda(), prm).view(240, -1).cpu().numpy()
    want = po.featuregram(y, 16000, 128, 32, 31, 31, 2048, 120, "LogMelHarmPercSpec")
    assert rel_l2(got, want) < TOL


def test_ragged_batch_musan_like(ctx):
    """Variable clip lengths (0.1 s .. 20 s) in one batch, k = 21/11."""
    lens = [1600, 16000, 320000, 2000, 48000, 1600 + 159, 100000]
    waves = [synth.synth_clip(200 + i, L) for i, L in enumerate(lens)]
    prm = engine.make_params(l_harm=21, l_perc=11, n_mels=120)
    batch = engine.Batch(ctx, clip_lengths=lens, n_fft=400, hop_length=160)
    out = engine.featuregram(batch, torch.from_numpy(np.concatenate(waves)).cuda(), prm)
    for c, y in enumerate(waves):
        want = po.featuregram(y, 16000, 25, 10, 21, 11, 400, 120, "LogMelHarmPercSpec")
        got = batch.clip(out, 240, c).cpu().numpy()
        assert got.shape == want.shape and rel_l2(got, want) < TOL, c


def test_corpus_shaped_batch_properties(ctx):
    """BASELINE.json configs[2] shape at 1/8 scale (136 MUSAN-length clips, ~12 h, k = 21/11) in ONE ragged batch:
    every selected clip must come out bit-identical to the same clip processed alone (a different tiling: the ragged
    batch walks a tile prefix table over contiguous tile ranges, the single clip takes the uniform path), and the
    medians of the shortest clip are checked against scipy."""
    rng = np.random.default_rng(2024)
    d = np.concatenate([rng.gamma(4.0, 232.0 / 4.0, size=83), rng.gamma(3.0, 511.0 / 3.0, size=53)])
    lens = [int(x * 16000) for x in np.clip(d, 5.0, 1800.0)]
    offs = np.concatenate([[0], np.cumsum(lens)])
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    wave = torch.randn(int(offs[-1]), device="cuda", generator=g) * 0.2
    wave += 0.4 * torch.sin(torch.arange(wave.numel(), device="cuda", dtype=torch.float32) * (2 * np.pi * 440.0 / 16000))
    prm = engine.make_params(l_harm=21, l_perc=11, n_mels=120)
    batch = engine.Batch(ctx, clip_lengths=lens, n_fft=400, hop_length=160)
    out = engine.featuregram(batch, wave, prm)
    assert torch.isfinite(out).all()
    order = np.argsort(lens)
    picks = sorted({0, len(lens) - 1, int(order[0]), int(order[-1]), int(order[len(order) // 2]), 17, 99})
    for c in picks:
        y = wave[offs[c]:offs[c + 1]].clone()
        one = engine.Batch(ctx, clip_lengths=[lens[c]], n_fft=400, hop_length=160)
        alone = engine.featuregram(one, y, prm)
        assert torch.equal(batch.clip(out, 240, c).reshape(-1), alone), f"clip {c} ({lens[c]} samples)"
    c = int(order[0])
    S = engine.stft_mag(batch, wave, 400, 400, 160)
    Sc = batch.clip(S, 201, c).cpu().numpy()
    harm = batch.clip(engine.median_time(batch, S, 201, 21), 201, c).cpu().numpy()
    perc = batch.clip(engine.median_freq(batch, S, 201, 11), 201, c).cpu().numpy()
    assert np.array_equal(harm, lr.median_filter_scipy(Sc, 21, axis=1))
    assert np.array_equal(perc, lr.median_filter_scipy(Sc, 11, axis=0))


def test_one_hour_stream_full_size(ctx):
    """BASELINE.json configs[3] at full size: one 1-hour 16 kHz stream, n_fft 2048, hop 512, k = 31
    (112 497 frames x 1025 bins).  Medians bit-exact against scipy on row / column subsets (time-axis tiles
    with halos, frequency walk with 65 steps), STFT against the oracle on a slice, features finite and
    consistent with the top_db floor."""
    L = 57_600_000
    g = torch.Generator(device="cuda").manual_seed(7)
    wave = torch.randn(L, generator=g, device="cuda") * 0.3
    wave += 0.5 * torch.sin(torch.arange(L, device="cuda", dtype=torch.float32) * (2 * np.pi * 440.0 / 16000))
    wave /= wave.abs().max()
    batch = engine.Batch(ctx, clip_lengths=[L], n_fft=2048, hop_length=512)
    T = batch.total_frames
    assert T == 1 + (L - 2048) // 512 == 112497
    S = engine.stft_mag(batch, wave, 2048, 2048, 512)
    harm = engine.median_time(batch, S, 1025, 31).view(1025, T)
    perc = engine.median_freq(batch, S, 1025, 31).view(1025, T)
    Sv = S.view(1025, T)
    rows = [0, 1, 511, 1024]
    Sr = Sv[rows].cpu().numpy()
    assert np.array_equal(harm[rows].cpu().numpy(), lr.median_filter_scipy(Sr, 31, axis=1))
    cols = slice(56000, 56300)
    Sc = Sv[:, cols].cpu().numpy()
    assert np.array_equal(perc[:, cols].cpu().numpy(), lr.median_filter_scipy(Sc, 31, axis=0))
    # STFT of a slice of the stream against the oracle (frame indexing at a large offset)
    t0 = 100_000
    y = wave[t0 * 512:t0 * 512 + 2048 + 512 * 63].cpu().numpy()
    want = np.abs(lr.stft(y, n_fft=2048, hop_length=512, win_length=2048))
    assert rel_l2(Sv[:, t0:t0 + 64].cpu().numpy(), want) < TOL
    prm = engine.make_params(n_fft=2048, win_length=2048, hop_length=512, l_harm=31, l_perc=31, n_mels=120)
    out = engine.featuregram(batch, wave, prm).view(2, 120, T)
    assert torch.isfinite(out).all()
    mx, mn = out.amax(dim=(1, 2)), out.amin(dim=(1, 2))
    assert (mn >= mx - 80.0 - 1e-3).all()


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
@pytest.mark.parametrize("k", [17, 31, 63])
def test_sweep_config_parity(ctx, n_fft, k):
    """BASELINE.json configs[4]: n_fft in {512, 1024, 2048} (hop n_fft/4) x median kernel in {17, 31, 63}:
    medians bit-exact against scipy, features against the oracle, two ragged clips per case."""
    hop = n_fft // 4
    Ls = [3 * 16000 + 123, 16000]
    waves = [synth.synth_clip(900 + i, L) for i, L in enumerate(Ls)]
    batch = engine.Batch(ctx, clip_lengths=Ls, n_fft=n_fft, hop_length=hop)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    F = n_fft // 2 + 1
    S = engine.stft_mag(batch, wave, n_fft, n_fft, hop)
    harm = engine.median_time(batch, S, F, k)
    perc = engine.median_freq(batch, S, F, k)
    prm = engine.make_params(n_fft=n_fft, win_length=n_fft, hop_length=hop, l_harm=k, l_perc=k, n_mels=120)
    out = engine.featuregram(batch, wave, prm)
    frame_ms, hop_ms = n_fft / 16.0, hop / 16.0
    for c, y in enumerate(waves):
        Sc = batch.clip(S, F, c).cpu().numpy()
        assert np.array_equal(batch.clip(harm, F, c).cpu().numpy(), lr.median_filter_scipy(Sc, k, axis=1))
        assert np.array_equal(batch.clip(perc, F, c).cpu().numpy(), lr.median_filter_scipy(Sc, k, axis=0))
        want = po.featuregram(y, 16000, frame_ms, hop_ms, k, k, n_fft, 120, "LogMelHarmPercSpec")
        got = batch.clip(out, 240, c).cpu().numpy()
        assert got.shape == want.shape and rel_l2(got, want) < TOL, (c, rel_l2(got, want))
