"""librosa-shaped entry points for the calls the reference makes (numpy in, numpy out, GPU inside).

Mirrors the argument meaning and error behaviour of the librosa (~0.8) functions that
lib/preprocessing.py calls, restricted to the configurations on the reference path:
``stft(center=False)``, ``hpss(power=2.0, margin=1.0)``, ``melspectrogram``, ``power_to_db``,
``filters.mel`` (Slaney scale and norm), plus ``mfcc`` as an extension the reference does not call.  Anything else raises NotImplementedError rather than
silently computing something different.
"""
from __future__ import annotations

import numpy as np

from ._lib import ParameterError


def _ctx():
    from . import engine
    return engine.get_context()


def mel(sr, n_fft, n_mels=128):
    """librosa.filters.mel(sr, n_fft, n_mels) (fmin=0, fmax=sr/2, htk=False, norm='slaney')."""
    from . import engine
    return engine.mel_filterbank(int(sr), int(n_fft), int(n_mels))


def stft(y, n_fft=2048, hop_length=None, win_length=None, center=False):
    """librosa.core.stft -> complex64 (1 + n_fft//2, T); only center=False (every reference call site)."""
    import torch
    from . import engine
    if center:
        raise NotImplementedError("center=True is not on the reference path")
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim != 1:
        raise ParameterError("only mono input is supported")
    ctx = _ctx()
    batch = engine.Batch(ctx, clip_lengths=[y.shape[0]], n_fft=n_fft, hop_length=hop_length)   # raises if n_fft > len(y)
    wave = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda()
    engine.validate_audio(ctx, wave)                       # librosa.util.valid_audio, on the device
    _, cplx = engine.stft_mag(batch, wave, n_fft, win_length, hop_length, return_complex=True)
    engine.ctx_check(ctx)                                  # raises ParameterError("Audio buffer is not finite everywhere")
    out = cplx.cpu().numpy().reshape(1 + n_fft // 2, -1)
    batch.close()
    return out


def hpss(S, kernel_size=31, power=2.0, mask=False, margin=1.0):
    """librosa.decompose.hpss on a magnitude spectrogram (F, T) -> (H, P) float32."""
    import torch
    from . import engine
    S = np.asarray(S)
    if np.iscomplexobj(S):
        raise NotImplementedError("complex input: pass np.abs(D) (the reference always does)")
    if np.isscalar(kernel_size):
        win_harm = win_perc = kernel_size
    else:
        win_harm, win_perc = kernel_size
    mh, mp = (margin, margin) if np.isscalar(margin) else margin
    if mh < 1 or mp < 1:
        raise ParameterError("Margins must be >= 1.0. A typical range is between 1 and 10.")
    if power != 2.0 or mh != 1.0 or mp != 1.0 or mask:
        raise NotImplementedError("only power=2.0, margin=1.0, mask=False (librosa's defaults, used by the reference)")
    rows, T = S.shape
    ctx = _ctx()
    batch = engine.Batch(ctx, clip_frames=[T])
    Sd = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32).ravel()).cuda()
    engine.validate_nonneg(ctx, Sd)                        # softmask's input check, on the device (medians of
    harm = engine.median_time(batch, Sd, rows, int(win_harm))   # non-negative data are non-negative)
    perc = engine.median_freq(batch, Sd, rows, int(win_perc))
    out, _ = engine.mask_mel_log(batch, Sd, harm, perc, rows)
    engine.ctx_check(ctx)                                  # raises ParameterError("X and X_ref must be non-negative")
    res = out.cpu().numpy().reshape(2 * rows, T)
    batch.close()
    return res[:rows].copy(), res[rows:].copy()


def melspectrogram(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, center=True, power=2.0,
                   n_mels=128):
    """librosa.feature.melspectrogram: with ``S`` the basis is built for n_fft = 2*(S.shape[0]-1); with ``y``
    the power spectrogram |stft|**2 is projected."""
    import torch
    from . import engine
    ctx = _ctx()
    if S is not None:
        S = np.ascontiguousarray(S, dtype=np.float32)
        rows, T = S.shape
        n_fft = 2 * (rows - 1)
        batch = engine.Batch(ctx, clip_frames=[T])
        Sd = torch.from_numpy(S.ravel()).cuda()
        pre_square = False
    else:
        if center:
            raise NotImplementedError("center=True is not on the reference path")
        if power != 2.0:
            raise NotImplementedError("only power=2.0")
        y = np.ascontiguousarray(y, dtype=np.float32)
        win_length = n_fft if win_length is None else win_length
        batch = engine.Batch(ctx, clip_lengths=[y.shape[0]], n_fft=n_fft, hop_length=hop_length)
        Sd = engine.stft_mag(batch, torch.from_numpy(y).cuda(), n_fft, win_length, hop_length)
        rows, T = 1 + n_fft // 2, batch.total_frames
        pre_square = True
    basis = torch.from_numpy(engine.mel_filterbank(int(sr), n_fft, int(n_mels))).cuda()
    out, _ = engine.mask_mel_log(batch, Sd, None, None, rows, mel=basis, pre_square=pre_square)
    res = out.cpu().numpy().reshape(int(n_mels), T)
    batch.close()
    return res


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.core.power_to_db for ref=1.0 on a real (rows, T) array: 10*log10(max(amin, S)), clipped at
    (global max - top_db)."""
    import torch
    from . import engine
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if ref != 1.0:
        raise NotImplementedError("only ref=1.0 (the reference's default)")
    if top_db is not None and top_db < 0:
        raise ParameterError("top_db must be non-negative")
    S = np.asarray(S)
    if np.iscomplexobj(S):
        raise NotImplementedError("complex input")
    rows, T = S.shape
    ctx = _ctx()
    batch = engine.Batch(ctx, clip_frames=[T])
    Sd = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32).ravel()).cuda()
    out, cmax = engine.mask_mel_log(batch, Sd, None, None, rows, log_power=2, amin=float(amin))   # 2: S is a power
    if top_db is not None:
        engine.topdb_clip(batch, out, rows, 1, cmax, float(top_db))
    res = out.cpu().numpy().reshape(rows, T)
    batch.close()
    return res


def mfcc(y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", **kwargs):
    """librosa.feature.mfcc (an extension: the reference never calls it).  ``S`` is a log-power mel spectrogram
    (n_mels, T); with ``y`` it is power_to_db(melspectrogram(y, **kwargs)).  Returns
    scipy.fftpack.dct(S, axis=0, type=2, norm='ortho')[:n_mfcc]."""
    import torch
    from . import engine
    if dct_type != 2 or norm != "ortho":
        raise NotImplementedError("only dct_type=2, norm='ortho' (librosa's defaults)")
    if S is None:
        S = power_to_db(melspectrogram(y=y, sr=sr, **kwargs))
    S = np.ascontiguousarray(S, dtype=np.float32)
    n_mels, T = S.shape
    if not 1 <= n_mfcc <= n_mels:
        raise ParameterError("n_mfcc must be in [1, n_mels]")
    ctx = _ctx()
    batch = engine.Batch(ctx, clip_frames=[T])
    out = engine.dct_mfcc(batch, torch.from_numpy(S.ravel()).cuda(), n_mels, 1, int(n_mfcc))
    res = out.cpu().numpy().reshape(int(n_mfcc), T)
    batch.close()
    return res
