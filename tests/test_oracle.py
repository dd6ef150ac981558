"""CPU tests of the oracle: against the golden vectors produced by the reference's own glue
(tests/golden/make_golden.py), against the reference's compiled Cython leaf (oracle/_ref) and
against independent implementations of the librosa leaves available in this image."""
import os
import warnings

import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import preprocessing_oracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_glue.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------ golden: reference glue
FEATS = ["Spec", "LogSpec", "MelSpec", "LogMelSpec", "PercSpec", "HarmPercSpec", "LogHarmPercSpec",
         "MelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec", "LogMelHarmPercSpec"]


@pytest.mark.parametrize("fn", FEATS)
def test_featuregram_matches_reference_glue(golden, fn):
    n_fft, n_mels = (512, 21) if fn == "LogHarmPercSpec" else (400, 40)
    for cls, sig in (("speech", "prep:/d/speech/sp0.wav"), ("speech_music", "mix:sp0+mu0@5")):
        want = golden[f"fv:{cls}:{fn}"]
        got = po.featuregram(golden[sig], 16000, 25, 10, 21, 11, n_fft, n_mels, fn)
        assert got.dtype == np.float32 and got.shape == want.shape
        # the golden mix is float64 inside the reference (numpy 2 promotion); ours is its float32 cast
        assert rel_l2(got, want) < 2e-6, (cls, fn)
    if fn == "LogMelHarmPercSpec":
        got = po.featuregram(golden["prep:/d/music/mu1.wav"], 16000, 25, 10, 21, 11, 400, 40, fn)
        assert np.array_equal(got, golden["fv:music:LogMelHarmPercSpec"])


def test_speech_featuregrams_bit_exact(golden):
    """No mixing involved -> identical dtype flow -> the restated glue must be bit-identical."""
    for fn in FEATS:
        n_fft, n_mels = (512, 21) if fn == "LogHarmPercSpec" else (400, 40)
        got = po.featuregram(golden["prep:/d/speech/sp0.wav"], 16000, 25, 10, 21, 11, n_fft, n_mels, fn)
        assert np.array_equal(got, golden[f"fv:speech:{fn}"]), fn


def test_harm_perc_rows_order(golden):
    both = golden["fv:speech:HarmPercSpec"]
    assert np.array_equal(both, golden["fv:speech:PercSpec"])          # startswith dispatch: same array
    assert both.shape[0] == 2 * 201


@pytest.mark.parametrize("model", ["Lemaire_et_al_MTL", "Doukhan_et_al_MTL"])
@pytest.mark.parametrize("fn", ["LogMelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec"])
@pytest.mark.parametrize("W,sh", [(68, 68), (49, 24), (249, 24)])
def test_patches_match_reference_glue(golden, model, fn, W, sh):
    FV = golden["fv:speech:LogMelHarmPercSpec"]
    want = golden[f"patch:{model}:{fn}:{W}:{sh}"]
    got = po.get_feature_patches(FV, W, sh, fn, model)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.allclose(got, want, rtol=0, atol=1e-6)


def test_patches_spec_and_frame_level_scaling(golden):
    got = po.get_feature_patches(golden["fv:speech:Spec"], 21, 21, "Spec", "Doukhan_et_al_MTL")
    assert np.allclose(got, golden["patch:Doukhan_et_al_MTL:Spec:21:21"], rtol=0, atol=1e-6)
    got = po.get_feature_patches(golden["fv:speech:LogMelHarmPercSpec"], 49, 24, "LogMelHarmPercSpec",
                                 "Lemaire_et_al_MTL", frame_level_scaling=True)
    assert np.array_equal(got, golden["patch:fls:LogMelHarmPercSpec:49:24"])


def test_data_stats_match_reference_glue(golden):
    names = ["music", "speech", "speech_music"]
    groups = {n: [] for n in names}
    for k in golden.files:
        if k.startswith("statsfv:"):
            _, cls, _ = k.split(":", 2)
            groups[cls].append(golden[k])
    mean, std, n0, n1, n2 = po.get_data_stats(groups, names)
    assert [n0, n1, n2] == list(golden["stats:counts"])
    assert np.allclose(mean, golden["stats:mean"], rtol=1e-6, atol=1e-6)
    assert np.allclose(std, golden["stats:std"], rtol=1e-6, atol=1e-6)
    FV = golden["fv:speech:LogMelHarmPercSpec"]
    assert np.allclose(po.scale_data(FV, golden["stats:mean"], golden["stats:std"]), golden["scaled:py"])
    assert np.array_equal(po.cscale_data(FV, golden["stats:mean"], golden["stats:std"]), golden["scaled:cy"])


# ------------------------------------------------------------------ reference Cython leaf
def test_against_compiled_reference_leaf():
    from oracle import build_ref
    tools = build_ref.load()
    if tools is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(3)
    FV = rng.standard_normal((80, 333)).astype(np.float32)
    for (W, sh) in [(249, 24), (68, 68), (99, 1), (21, 5), (68, 34)]:
        assert np.array_equal(tools.extract_patches(FV, FV.shape, W, sh), po.extract_patches(FV, W, sh))
    mean, std = FV.mean(1), FV.std(1)
    assert np.array_equal(tools.scale_data(FV, mean, std), po.cscale_data(FV, mean, std))


# ------------------------------------------------------------------ leaves vs independent implementations
@pytest.mark.parametrize("n_fft,win,hop", [(400, 400, 160), (512, 400, 160), (2048, 2048, 512)])
def test_stft_vs_torch_float64(n_fft, win, hop):
    import torch
    rng = np.random.default_rng(0)
    y = rng.standard_normal(3 * n_fft + 777).astype(np.float32)
    S = lr.stft(y, n_fft=n_fft, hop_length=hop, win_length=win)
    w = torch.hann_window(win, periodic=True, dtype=torch.float64)
    St = torch.stft(torch.from_numpy(y).double(), n_fft=n_fft, hop_length=hop, win_length=win, window=w, center=False,
                    return_complex=True).numpy()
    assert S.dtype == np.complex64 and S.shape == St.shape == (1 + n_fft // 2, 1 + (len(y) - n_fft) // hop)
    assert np.linalg.norm(S - St) / np.linalg.norm(St) < 1e-7


def test_stft_errors_like_librosa():
    with pytest.raises(lr.ParameterError):
        lr.stft(np.zeros(399, np.float32), n_fft=400, hop_length=160, win_length=400)
    with pytest.raises(lr.ParameterError):
        lr.stft(np.full(1000, np.nan, np.float32), n_fft=400, hop_length=160, win_length=400)


@pytest.mark.parametrize("sr,n_fft,n_mels", [(22050, 400, 120), (16000, 400, 120), (22050, 512, 21)])
def test_mel_vs_torchaudio_and_transformers(sr, n_fft, n_mels):
    m = lr.mel(sr, n_fft, n_mels)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchaudio
        mt = torchaudio.functional.melscale_fbanks(1 + n_fft // 2, 0.0, sr / 2, n_mels, sr, norm="slaney",
                                                   mel_scale="slaney").T.numpy()
        from transformers.audio_utils import mel_filter_bank
        mh = mel_filter_bank(1 + n_fft // 2, n_mels, 0.0, sr / 2, sr, norm="slaney", mel_scale="slaney").T
    assert np.abs(m - mt).max() < 5e-7
    assert np.abs(m - mh).max() < 5e-7
    assert m.dtype == np.float32 and (m >= 0).all()


def test_mel_sr22050_basis_is_banded():
    m = lr.mel(22050, 400, 120)
    assert (m > 0).sum() == 393 and (m > 0).sum(axis=1).max() == 11 and (m[0] == 0).all()   # SURVEY.md probe
    for row in m:                                                    # contiguous support
        nz = np.flatnonzero(row)
        if nz.size:
            assert nz[-1] - nz[0] + 1 == nz.size


@pytest.mark.parametrize("n,k", [(98, 31), (998, 21), (201, 11), (8, 21), (5, 31), (40, 16), (9, 63), (33, 17)])
def test_median_vs_scipy(n, k):
    rng = np.random.default_rng(n * 100 + k)
    m = np.abs(rng.standard_normal((6, n))).astype(np.float32)
    m = np.where(rng.random(m.shape) < 0.3, np.round(m * 4) / 4, m).astype(np.float32)
    assert lr.scipy_median_well_defined(n, k)
    assert np.array_equal(lr.median_filter_1d(m, k, 1), lr.median_filter_scipy(m, k, 1))
    assert np.array_equal(lr.median_filter_1d(m.T.copy(), k, 0), lr.median_filter_scipy(m.T.copy(), k, 0))


def test_scipy_reflect_overshoot_bug():
    """Where k//2 >= 4n scipy reads the element before the line; documents why the oracle states the
    median itself.  (If a future scipy fixes this, the flagged regime simply becomes equal too.)"""
    rng = np.random.default_rng(0)
    m = np.abs(rng.standard_normal((5, 2))).astype(np.float32)
    assert not lr.scipy_median_well_defined(2, 63)
    ours = lr.median_filter_1d(m, 63, 1)
    # mathematically: 63 taps over the period-4 extension of (x0, x1) -> x1 at t=0, x0 at t=1
    assert np.array_equal(ours, m[:, ::-1])


def test_softmask_properties():
    rng = np.random.default_rng(1)
    a = np.abs(rng.standard_normal((20, 30))).astype(np.float32)
    b = np.abs(rng.standard_normal((20, 30))).astype(np.float32)
    a[0, :5] = 0
    b[0, :5] = 0
    m1 = lr.softmask(a, b, power=2, split_zeros=True)
    m2 = lr.softmask(b, a, power=2, split_zeros=True)
    assert m1.dtype == np.float32
    assert np.allclose(m1 + m2, 1.0, atol=1e-6)
    assert (m1[0, :5] == 0.5).all()
    with pytest.raises(lr.ParameterError):
        lr.softmask(-a, b)
    with pytest.raises(lr.ParameterError):
        lr.hpss(a, margin=0.5)


def test_power_to_db_vs_transformers():
    from transformers.audio_utils import power_to_db
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((40, 50)) ** 2).astype(np.float32)
    x[3, 4] = 0.0
    got = lr.power_to_db(x)
    want = power_to_db(x, reference=1.0, min_value=1e-10, db_range=80.0)
    assert got.dtype == np.float32
    assert np.allclose(got, want, atol=1e-4)
    assert got.min() >= got.max() - 80.0 - 1e-4


def test_hpss_uses_scipy_and_own_median_identically():
    y = np.random.default_rng(4).standard_normal(16000).astype(np.float32)
    S = np.abs(lr.stft(y, n_fft=400, hop_length=160, win_length=400))
    H1, P1 = lr.hpss(S, kernel_size=(21, 11), use_scipy=True)
    H2, P2 = lr.hpss(S, kernel_size=(21, 11), use_scipy=False)
    assert np.array_equal(H1, H2) and np.array_equal(P1, P2)


# ------------------------------------------------------------------------------ MFCC extension (not in the reference)
@pytest.mark.parametrize("M,n", [(120, 20), (21, 13), (128, 40), (7, 7)])
def test_dct_restatement_vs_scipy(M, n):
    """Parity pin of the MFCC extension: scipy's own orthonormal DCT-II (what librosa.feature.mfcc calls)."""
    import scipy.fft
    rng = np.random.default_rng(M)
    S = rng.standard_normal((M, 33)) * 30 - 40
    want = scipy.fft.dct(S, axis=0, type=2, norm="ortho")[:n]
    got = lr.dct_ortho(S, n)
    assert np.allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())
