"""Known-answer tests of the librosa leaves whose expected values do NOT pass through the oracle: closed-form
spectra, exact rational soft masks, a hand-computed Slaney basis, power_to_db edge cases, and a whole-pipeline
cross-check against an independent third-party implementation (transformers.audio_utils.spectrogram).

librosa itself cannot be installed here (oracle/__init__.py), so these are what pins the restated leaves -- and, in
the ``gpu`` half, the CUDA kernels -- to something other than our own restatement.  The same expectations are
applied to the oracle (CPU tests) and to the kernels (GPU tests)."""
import warnings

import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import preprocessing_oracle as po


# ------------------------------------------------------------------------------------------------ closed forms
def hann_periodic_closed_form(win, n_fft):
    """scipy.signal.get_window('hann', win, fftbins=True) = 0.5 - 0.5 cos(2 pi n / win), centred in n_fft samples
    (librosa.util.pad_center: left pad (n_fft - win) // 2)."""
    w = np.zeros(n_fft)
    lpad = (n_fft - win) // 2
    w[lpad:lpad + win] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win) / win)
    return w


def impulse_case(n_fft, win, hop, n_frames=9, pos=None):
    """x = delta at sample `pos`: frame t sees it at n = pos - t*hop, so |X[k, t]| = w[pos - t*hop] for EVERY bin k
    (0 when the impulse is outside the frame).  Pins hop, frame alignment (center=False), the window shape and its
    centring in the FFT frame -- a symmetric Hann or an un-centred window give different numbers."""
    L = n_fft + (n_frames - 1) * hop
    pos = (L // 2 + 3) if pos is None else pos
    x = np.zeros(L, np.float32)
    x[pos] = 1.0
    w = hann_periodic_closed_form(win, n_fft)
    want = np.zeros((n_fft // 2 + 1, n_frames))
    for t in range(n_frames):
        n = pos - t * hop
        if 0 <= n < n_fft:
            want[:, t] = w[n]
    return x, want


def sinusoid_case(n_fft, hop, k0, amp=0.75, phase=0.4, n_frames=6):
    """x[n] = A cos(2 pi k0 n / n_fft + phi) with win = n_fft: the periodic Hann window's DFT is N/2 at bin 0, -N/4 at
    bins +-1 and exactly 0 elsewhere, so |X[k0]| = A N / 4, |X[k0 +- 1]| = A N / 8 and every other bin is 0
    (2 <= k0 <= N/2 - 2), in every frame."""
    N = n_fft
    L = N + (n_frames - 1) * hop
    n = np.arange(L)
    x = (amp * np.cos(2 * np.pi * k0 * n / N + phase)).astype(np.float32)
    want = np.zeros((N // 2 + 1, n_frames))
    want[k0] = amp * N / 4
    want[k0 - 1] = want[k0 + 1] = amp * N / 8
    return x, want


STFT_CASES = [(400, 400, 160), (512, 400, 160), (512, 512, 128), (1024, 1024, 256), (2048, 2048, 512)]
SOFTMASK_PAIRS = [(3, 4), (4, 3), (1, 2), (1, 4), (5, 8), (7, 8), (1, 1), (0, 5), (5, 0), (0, 0), (1, 1024)]


def softmask_expected(x, xr):
    """mask = (x/Z)^2 / ((x/Z)^2 + (xr/Z)^2), Z = max: for these integer pairs x/Z and its square are exact in float32
    (denominators are powers of two), so the only rounding is the last division: RN of an exact rational.
    (0, 0) is the split_zeros case: 0.5."""
    if x == 0 and xr == 0:
        return np.float32(0.5)
    z = max(x, xr)
    a, b = (x / z) ** 2, (xr / z) ** 2                 # exact in binary64, and representable in binary32
    assert np.float32(a) == a and np.float32(b) == b
    return np.float32(a / (a + b))                      # single rounding of the exact quotient (53 -> 24 bits: a + b is exact)


def slaney_hand_basis():
    """sr = 2000, n_fft = 16, 3 filters: fmax = 1000 Hz = 15 mel is the end of Slaney's LINEAR region (200/3 Hz per mel),
    so the five band edges are 0, 250, 500, 750, 1000 Hz and the FFT bins sit every 125 Hz: every triangle is
    (0.5, 1, 0.5) on three bins, times the Slaney area normalisation 2 / (f[i+2] - f[i]) = 2 / 500."""
    m = np.zeros((3, 9))
    for i in range(3):
        m[i, 2 * i + 1:2 * i + 4] = np.array([0.5, 1.0, 0.5]) * (2.0 / 500.0)
    return m


def check_mel_log_region(mel_fn):
    """Slaney's log region: 1000 Hz * 6.4^((mel - 15) / 27), i.e. 6400 Hz = 42 mel.  sr = 12800, 13 filters -> edges every
    3 mel: 0, 200, ..., 1000 Hz, then 1000 * 6.4^(j/9).  Each filter must be non-zero exactly on the bins strictly
    inside (f[i], f[i+2]) and integrate to one (Slaney norm)."""
    sr, n_fft, n_mels = 12800, 2560, 13                # 5 Hz bins
    m = mel_fn(sr, n_fft, n_mels)
    mels = 3.0 * np.arange(n_mels + 2)
    edges = np.where(mels <= 15.0, mels * 200.0 / 3.0, 1000.0 * 6.4 ** ((mels - 15.0) / 27.0))
    assert abs(edges[-1] - 6400.0) < 1e-9
    freqs = np.arange(n_fft // 2 + 1) * (sr / n_fft)
    for i in range(n_mels):
        inside = (freqs > edges[i] + 1e-6) & (freqs < edges[i + 2] - 1e-6)
        outside = (freqs < edges[i] - 1e-6) | (freqs > edges[i + 2] + 1e-6)
        assert (m[i][inside] > 0).all() and (m[i][outside] == 0).all(), i
        assert abs(m[i].sum() * (sr / n_fft) - 1.0) < 2e-3, (i, m[i].sum() * (sr / n_fft))     # unit area
        peak = freqs[np.argmax(m[i])]
        assert abs(peak - edges[i + 1]) <= sr / n_fft, i


DB_IN = np.array([[1.0, 1e-3, 1e-9, 0.0, 10.0, 1e-12]], dtype=np.float32)
DB_NOCLIP = np.array([[0.0, -30.0, -90.0, -100.0, 10.0, -100.0]])        # amin = 1e-10 floors 0 and 1e-12 at -100 dB
DB_TOP80 = np.array([[0.0, -30.0, -70.0, -70.0, 10.0, -70.0]])           # max is 10 dB: floor at 10 - 80


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("n_fft,win,hop", STFT_CASES)
def test_oracle_stft_impulse(n_fft, win, hop):
    x, want = impulse_case(n_fft, win, hop)
    got = np.abs(lr.stft(x, n_fft=n_fft, hop_length=hop, win_length=win))
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-6


@pytest.mark.parametrize("n_fft,hop,k0", [(400, 160, 37), (400, 160, 2), (400, 160, 198), (512, 128, 100), (2048, 512, 777)])
def test_oracle_stft_bin_centred_sinusoid(n_fft, hop, k0):
    x, want = sinusoid_case(n_fft, hop, k0)
    got = np.abs(lr.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft))
    assert np.abs(got - want).max() < 2e-6 * want.max()                 # float32 samples: 6e-8 relative input noise


def test_oracle_softmask_exact_rationals():
    X = np.array([[p[0] for p in SOFTMASK_PAIRS]], dtype=np.float32)
    R = np.array([[p[1] for p in SOFTMASK_PAIRS]], dtype=np.float32)
    want = np.array([[softmask_expected(*p) for p in SOFTMASK_PAIRS]], dtype=np.float32)
    assert np.array_equal(lr.softmask(X, R, power=2, split_zeros=True), want)
    # scale invariance by powers of two (exact): the same masks at 2^-20 and 2^+20
    for s in (np.float32(2.0 ** -20), np.float32(2.0 ** 20)):
        assert np.array_equal(lr.softmask(X * s, R * s, power=2, split_zeros=True), want)


def test_oracle_mel_hand_computed_and_log_region():
    assert np.abs(lr.mel(2000, 16, 3) - slaney_hand_basis()).max() < 1e-9
    check_mel_log_region(lr.mel)


def test_oracle_power_to_db_edges():
    assert np.abs(lr.power_to_db(DB_IN, top_db=None) - DB_NOCLIP).max() < 1e-4
    assert np.abs(lr.power_to_db(DB_IN) - DB_TOP80).max() < 1e-4
    with pytest.raises(lr.ParameterError):
        lr.power_to_db(DB_IN, amin=0)
    with pytest.raises(lr.ParameterError):
        lr.power_to_db(DB_IN, top_db=-1)


def _transformers_logmel(y, sr, n_fft, win, hop, n_mels):
    """The reference's LogMelSpec branch (lib/preprocessing.py:397-402) written with transformers.audio_utils only:
    mel of the POWER spectrogram (sr = fs), squared again, power_to_db with top_db = 80."""
    from transformers.audio_utils import mel_filter_bank, power_to_db, spectrogram, window_function
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # librosa frames n_fft samples and multiplies by the window centred in the frame; transformers would frame
        # win samples and pad on the right, so it is given the centred n_fft-long window (closed form above)
        window = window_function(win, "hann", periodic=True) if win == n_fft else hann_periodic_closed_form(win, n_fft)
        fb = mel_filter_bank(1 + n_fft // 2, n_mels, 0.0, sr / 2, sr, norm="slaney", mel_scale="slaney")
        mel_pow = spectrogram(y.astype(np.float64), window, frame_length=n_fft, hop_length=hop, fft_length=n_fft, power=2.0,
                              center=False, mel_filters=fb, mel_floor=0.0, dtype=np.float64)
    return mel_pow, power_to_db(mel_pow ** 2, reference=1.0, min_value=1e-10, db_range=80.0)


@pytest.mark.parametrize("n_fft,win", [(400, 400), (512, 400)])
def test_oracle_full_pipeline_vs_transformers(n_fft, win):
    """Framing + window + FFT + |.|^2 + Slaney mel + power_to_db end to end against a third-party implementation."""
    from sm_hpss_mtl_b200 import synth
    y = synth.synth_clip(11, 24000)
    mel_pow, logmel = _transformers_logmel(y, 16000, n_fft, win, 160, 40)
    got_mel = po.featuregram(y, 16000, 25, 10, 21, 11, n_fft, 40, "MelSpec")
    got_log = po.featuregram(y, 16000, 25, 10, 21, 11, n_fft, 40, "LogMelSpec")
    assert got_mel.shape == mel_pow.shape
    assert np.linalg.norm(got_mel - mel_pow) / np.linalg.norm(mel_pow) < 1e-6
    assert np.abs(got_log - logmel).max() < 1e-3


# ------------------------------------------------------------------------------------------------ kernels (GPU)
gpu = pytest.mark.gpu


def _stft_gpu(ctx, x, n_fft, win, hop):
    import torch
    from sm_hpss_mtl_b200 import engine
    batch = engine.Batch(ctx, clip_lengths=[len(x)], n_fft=n_fft, hop_length=hop)
    S = engine.stft_mag(batch, torch.from_numpy(x).cuda(), n_fft, win, hop)
    return S.cpu().numpy().reshape(n_fft // 2 + 1, -1)


@gpu
@pytest.mark.parametrize("n_fft,win,hop", STFT_CASES)
def test_gpu_stft_impulse(ctx, n_fft, win, hop):
    x, want = impulse_case(n_fft, win, hop)
    got = _stft_gpu(ctx, x, n_fft, win, hop)
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-6
    x, want = impulse_case(n_fft, win, hop, pos=0)                     # first sample: only frame 0, at the window's edge
    assert np.abs(_stft_gpu(ctx, x, n_fft, win, hop) - want).max() < 2e-6


@gpu
@pytest.mark.parametrize("n_fft,hop,k0", [(400, 160, 37), (400, 160, 2), (400, 160, 198), (512, 128, 100), (2048, 512, 777)])
def test_gpu_stft_bin_centred_sinusoid(ctx, n_fft, hop, k0):
    x, want = sinusoid_case(n_fft, hop, k0)
    got = _stft_gpu(ctx, x, n_fft, n_fft, hop)
    assert np.abs(got - want).max() < 3e-6 * want.max()                 # fp32 FFT: ~1e-7 relative of the peak


@gpu
def test_gpu_softmask_exact_rationals(ctx):
    import torch
    from sm_hpss_mtl_b200 import engine
    n = len(SOFTMASK_PAIRS)
    for scale in (1.0, 2.0 ** -20, 2.0 ** 20):
        harm = np.tile(np.array([p[0] for p in SOFTMASK_PAIRS], np.float32) * np.float32(scale), (3, 1))
        perc = np.tile(np.array([p[1] for p in SOFTMASK_PAIRS], np.float32) * np.float32(scale), (3, 1))
        S = np.full((3, n), 4.0, np.float32)                            # a power of two: S * mask is exact
        batch = engine.Batch(ctx, clip_frames=[n])
        out, _ = engine.mask_mel_log(batch, torch.from_numpy(S.ravel()).cuda(), torch.from_numpy(harm.ravel()).cuda(),
                                     torch.from_numpy(perc.ravel()).cuda(), 3)
        out = out.cpu().numpy().reshape(6, n)
        want_h = np.array([softmask_expected(*p) for p in SOFTMASK_PAIRS], np.float32) * np.float32(4.0)
        want_p = np.array([softmask_expected(p[1], p[0]) for p in SOFTMASK_PAIRS], np.float32) * np.float32(4.0)
        assert np.array_equal(out[:3], np.tile(want_h, (3, 1))) and np.array_equal(out[3:], np.tile(want_p, (3, 1)))


@gpu
def test_gpu_mel_hand_computed_and_log_region(ctx):
    from sm_hpss_mtl_b200 import engine
    assert np.abs(engine.mel_filterbank(2000, 16, 3) - slaney_hand_basis()).max() < 1e-9
    check_mel_log_region(engine.mel_filterbank)


@gpu
def test_gpu_power_to_db_edges(ctx):
    from sm_hpss_mtl_b200 import librosa_compat as lc
    assert np.abs(lc.power_to_db(DB_IN, top_db=None) - DB_NOCLIP).max() < 1e-4
    assert np.abs(lc.power_to_db(DB_IN) - DB_TOP80).max() < 1e-4


@gpu
@pytest.mark.parametrize("n_fft,win", [(400, 400), (512, 400)])
def test_gpu_full_pipeline_vs_transformers(ctx, n_fft, win):
    from sm_hpss_mtl_b200 import preprocessing as pp
    from sm_hpss_mtl_b200 import synth
    y = synth.synth_clip(11, 24000)
    mel_pow, logmel = _transformers_logmel(y, 16000, n_fft, win, 160, 40)
    P = {"Tw": 25, "Ts": 10, "Model": "m", "l_harm": {"m": 21}, "l_perc": {"m": 11}}
    got_mel = pp.featuregram_from_signal(y, 16000, P, n_fft, 40, "MelSpec")
    got_log = pp.featuregram_from_signal(y, 16000, P, n_fft, 40, "LogMelSpec")
    assert np.linalg.norm(got_mel - mel_pow) / np.linalg.norm(mel_pow) < 1e-5
    assert np.abs(got_log - logmel).max() < 1e-3
