"""Restatement of the librosa / scipy leaf functions the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  librosa is an un-vendored,
unpinned dependency of the reference (no requirements file; era ~0.8.x, see
SURVEY.md section 8c) and is not installable here, so the functions below restate
its published algorithm.  Each one names the reference call site it serves.

All float32 arithmetic is written so that numpy reproduces librosa's own
rounding sequence (same dtypes at every step).
"""
from __future__ import annotations

import numpy as np


class ParameterError(ValueError):
    """Mirror of librosa.util.exceptions.ParameterError."""


# --------------------------------------------------------------------------
# a1  librosa.core.stft(..., center=False)    lib/preprocessing.py:381,387,407,417,429,439
# --------------------------------------------------------------------------
def hann_periodic(win_length: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win_length, fftbins=True): float64, periodic."""
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def pad_center(w: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center on a 1-D window: zero-pad, left pad = (size-n)//2."""
    n = w.shape[0]
    if n > size:
        raise ParameterError(f"Target size ({size}) must be at least input size ({n})")
    lpad = (size - n) // 2
    out = np.zeros(size, dtype=w.dtype)
    out[lpad:lpad + n] = w
    return out


def n_frames(length: int, n_fft: int, hop_length: int) -> int:
    """librosa.util.frame: 1 + (len - frame_length) // hop (no centring)."""
    return 1 + (length - n_fft) // hop_length


def stft(y, n_fft=2048, hop_length=None, win_length=None, center=False):
    """complex64 (1+n_fft//2, T) STFT; only ``center=False`` (all reference call sites).

    float64 window x float32 frame -> float64 rfft -> complex64, as librosa does.
    """
    if center:
        raise NotImplementedError("the reference always passes center=False")
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim != 1:
        raise ParameterError("only mono input is on the reference path")
    if not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    if n_fft > y.shape[-1]:
        raise ParameterError(
            f"n_fft={n_fft} is too large for input signal of length={y.shape[-1]}")
    w = pad_center(hann_periodic(win_length), n_fft).reshape(-1, 1)   # float64
    T = n_frames(y.shape[0], n_fft, hop_length)
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(T)[None, :]
    frames = y[idx]                                                   # (n_fft, T)
    out = np.empty((1 + n_fft // 2, T), dtype=np.complex64, order="F")
    # blockwise like librosa (keeps memory bounded on the 1-hour stream)
    step = max(1, (2 ** 22) // n_fft)
    for s in range(0, T, step):
        out[:, s:s + step] = np.fft.rfft(w * frames[:, s:s + step], axis=0)
    return out


# --------------------------------------------------------------------------
# a2  scipy.ndimage.median_filter(S, size=(1,k)|(k,1), mode='reflect')
#     called inside librosa.decompose.hpss   lib/preprocessing.py:408,418,430,440
# --------------------------------------------------------------------------
def reflect_index(i: np.ndarray, n: int) -> np.ndarray:
    """scipy 'reflect' (half-sample symmetric, d c b a | a b c d | d c b a), any overshoot."""
    i = np.mod(i, 2 * n)
    return np.where(i < n, i, 2 * n - 1 - i)


def median_filter_1d(S: np.ndarray, k: int, axis: int) -> np.ndarray:
    """Own statement of the sliding median: window offsets -k//2 .. k-1-k//2,
    reflect boundary, element of rank k//2.  Pure selection: bit-exact.
    (Checked bit-exact against scipy.ndimage.median_filter in tests.)"""
    S = np.asarray(S)
    n = S.shape[axis]
    offs = np.arange(k) - (k // 2)
    idx = reflect_index(np.arange(n)[:, None] + offs[None, :], n)     # (n, k)
    Sm = np.moveaxis(S, axis, -1)                                     # (..., n)
    win = Sm[..., idx]                                                # (..., n, k)
    med = np.partition(win, k // 2, axis=-1)[..., k // 2]
    return np.moveaxis(med, -1, axis)


def scipy_median_well_defined(n: int, k: int) -> bool:
    """scipy.ndimage's reflect handling returns index -1 (it reads the element *before* the
    line, i.e. the previous row or out-of-bounds memory) once a window offset reaches a negative
    multiple of 2n below -2n, i.e. when k//2 >= 4n.  Observed with scipy 1.18 on both the 1-D
    fast path and the legacy N-D rank filter; the reference never gets there (shortest clip is
    8 frames, lib/preprocessing.py:345-347; largest swept kernel 51).  Outside that regime scipy
    and ``median_filter_1d`` agree bit for bit; inside it the oracle (and the CUDA kernel)
    follow the mathematical definition."""
    return n <= 1 or (k // 2) < 4 * n


def median_filter_scipy(S: np.ndarray, k: int, axis: int) -> np.ndarray:
    """The routine librosa itself calls (fast path for the CPU baseline)."""
    from scipy.ndimage import median_filter
    size = [1] * S.ndim
    size[axis] = k
    return median_filter(S, size=tuple(size), mode="reflect")


# --------------------------------------------------------------------------
# a3  librosa.util.softmask + librosa.decompose.hpss
# --------------------------------------------------------------------------
def softmask(X, X_ref, power=1, split_zeros=False):
    X = np.asarray(X)
    X_ref = np.asarray(X_ref)
    if X.shape != X_ref.shape:
        raise ParameterError(f"Shape mismatch: {X.shape}!={X_ref.shape}")
    if np.any(X < 0) or np.any(X_ref < 0):
        raise ParameterError("X and X_ref must be non-negative")
    if power <= 0:
        raise ParameterError("power must be strictly positive")
    dtype = X.dtype
    if not np.issubdtype(dtype, np.floating):
        dtype = np.float32
    Z = np.maximum(X, X_ref).astype(dtype)
    bad_idx = Z < np.finfo(dtype).tiny
    Z[bad_idx] = 1
    if np.isfinite(power):
        mask = (X / Z) ** power
        ref_mask = (X_ref / Z) ** power
        good_idx = ~bad_idx
        mask[good_idx] /= mask[good_idx] + ref_mask[good_idx]
        mask[bad_idx] = 0.5 if split_zeros else 0.0
    else:
        mask = X > X_ref
    return mask


def hpss(S, kernel_size=31, power=2.0, mask=False, margin=1.0, use_scipy=True):
    """librosa.decompose.hpss on a real magnitude spectrogram (F, T)."""
    S = np.asarray(S)
    if np.iscomplexobj(S):
        phase = np.exp(1j * np.angle(S))
        S = np.abs(S)
    else:
        phase = 1
    if np.isscalar(kernel_size):
        win_harm = win_perc = kernel_size
    else:
        win_harm, win_perc = kernel_size
    if np.isscalar(margin):
        margin_harm = margin_perc = margin
    else:
        margin_harm, margin_perc = margin
    if margin_harm < 1 or margin_perc < 1:
        raise ParameterError("Margins must be >= 1.0. A typical range is between 1 and 10.")
    med = median_filter_scipy if use_scipy else median_filter_1d
    harm = np.empty_like(S)
    harm[:] = med(S, int(win_harm), axis=1)      # size=(1, win_harm): along time
    perc = np.empty_like(S)
    perc[:] = med(S, int(win_perc), axis=0)      # size=(win_perc, 1): along frequency
    split_zeros = margin_harm == 1 and margin_perc == 1
    mask_harm = softmask(harm, perc * margin_harm, power=power, split_zeros=split_zeros)
    mask_perc = softmask(perc, harm * margin_perc, power=power, split_zeros=split_zeros)
    if mask:
        return mask_harm, mask_perc
    return ((S * mask_harm) * phase, (S * mask_perc) * phase)


# --------------------------------------------------------------------------
# a4  librosa.filters.mel / librosa.feature.melspectrogram
# --------------------------------------------------------------------------
def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """Slaney-scale, slaney-normalised float32 (n_mels, 1+n_fft//2) basis."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]          # f32 *= f64: computed in f64, rounded to f32
    return weights


def melspectrogram(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                   center=True, power=2.0, n_mels=128):
    """librosa.feature.melspectrogram.  With ``S=`` (HPSS call sites,
    lib/preprocessing.py:409-410,419,421) ``n_fft = 2*(S.shape[0]-1)`` and ``sr``
    keeps its default 22050; with ``y=`` (lib/preprocessing.py:394,400) the power
    spectrogram is |stft|**power."""
    if S is not None:
        S = np.asarray(S)
        n_fft = 2 * (S.shape[0] - 1)
    else:
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                        center=center)) ** power
    mel_basis = mel(sr, n_fft, n_mels=n_mels)
    return np.dot(mel_basis, S)


# --------------------------------------------------------------------------
# a5  librosa.core.power_to_db      lib/preprocessing.py:388,401,420,422,441-442
# --------------------------------------------------------------------------
def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    magnitude = np.abs(S) if np.iscomplexobj(S) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ParameterError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def dct_ortho(S, n_out=None):
    """Orthonormal DCT-II along axis 0 in float64: what librosa.feature.mfcc applies to its log-mel input
    (scipy.fftpack.dct(S, axis=0, type=2, norm='ortho')[:n_mfcc]).  EXTENSION: the reference has no MFCC step, so
    this restates scipy's published definition and is pinned against scipy.fft.dct in tests/test_oracle.py."""
    S = np.asarray(S, dtype=np.float64)
    M = S.shape[0]
    n_out = M if n_out is None else n_out
    k = np.arange(n_out)[:, None]
    m = np.arange(M)[None, :]
    D = np.sqrt(2.0 / M) * np.cos(np.pi * k * (2 * m + 1) / (2.0 * M))
    D[0] *= np.sqrt(0.5)
    return D @ S
