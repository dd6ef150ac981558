"""CPU restatement of the reference's feature front-end glue (lib/preprocessing.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works on waveforms already in
memory (the reference's file loading / silence removal sit *before* the hot path,
SURVEY.md section 8 row N2).  Function names follow the reference.
"""
from __future__ import annotations

import numpy as np

from . import librosa_restated as lr


# ---- lib/preprocessing.py:114-132 -----------------------------------------
def normalize_signal(Xin):
    Xin = Xin - np.mean(Xin)
    Xin = Xin / np.max(np.abs(Xin))
    return Xin


# ---- lib/preprocessing.py:378-444 (body of get_featuregram after the signal is loaded)
def featuregram(Xin, fs, Tw, Ts, l_harm, l_perc, n_fft, n_mels, featName, use_scipy=True):
    """Waveform -> float32 (nFeat, T) featuregram, all feature names of the reference.

    Name dispatch follows the reference's ``if / if-elif`` chain: ``startswith`` tests,
    so e.g. 'LogMelHarmSpec', 'LogMelPercSpec' and 'LogMelHarmPercSpec' all produce
    the stacked (harmonic rows, then percussive rows) array.
    """
    frameSize = int(Tw * fs / 1000)
    frameShift = int(Ts * fs / 1000)
    kw = dict(n_fft=n_fft, win_length=frameSize, hop_length=frameShift, center=False)
    ks = (l_harm, l_perc)
    fv = None
    if featName == 'Spec':                                            # :378-382
        fv = np.abs(lr.stft(Xin, **kw)).astype(np.float32)
    if featName == 'LogSpec':                                         # :384-389
        fv = np.abs(lr.stft(Xin, **kw))
        fv = lr.power_to_db(fv ** 2).astype(np.float32)
    elif featName == 'MelSpec':                                       # :391-395
        fv = lr.melspectrogram(y=Xin, sr=fs, n_mels=n_mels, **kw).astype(np.float32)
    elif featName == 'LogMelSpec':                                    # :397-402
        fv = lr.melspectrogram(y=Xin, sr=fs, n_mels=n_mels, **kw)
        fv = lr.power_to_db(fv ** 2).astype(np.float32)
    elif featName.startswith('MelHarm') or featName.startswith('MelPerc'):      # :404-412
        Spec = np.abs(lr.stft(Xin, **kw))
        H, P = lr.hpss(Spec, kernel_size=ks, use_scipy=use_scipy)
        fv_H = lr.melspectrogram(S=H, n_mels=n_mels)
        fv_P = lr.melspectrogram(S=P, n_mels=n_mels)
        fv = np.append(fv_H, fv_P, axis=0).astype(np.float32)
    elif featName.startswith('LogMelHarm') or featName.startswith('LogMelPerc'):  # :414-424
        Spec = np.abs(lr.stft(Xin, **kw))
        H, P = lr.hpss(Spec, kernel_size=ks, use_scipy=use_scipy)
        fv_H = lr.power_to_db(lr.melspectrogram(S=H, n_mels=n_mels) ** 2)
        fv_P = lr.power_to_db(lr.melspectrogram(S=P, n_mels=n_mels) ** 2)
        fv = np.append(fv_H, fv_P, axis=0).astype(np.float32)
    elif featName.startswith('Harm') or featName.startswith('Perc'):            # :426-434
        Spec = np.abs(lr.stft(Xin, **kw))
        H, P = lr.hpss(Spec, kernel_size=ks, use_scipy=use_scipy)
        fv = np.append(H, P, axis=0).astype(np.float32)
    elif featName.startswith('LogHarm') or featName.startswith('LogPerc'):      # :436-444
        Spec = np.abs(lr.stft(Xin, **kw))
        H, P = lr.hpss(Spec, kernel_size=ks, use_scipy=use_scipy)
        fv = np.append(lr.power_to_db(H ** 2), lr.power_to_db(P ** 2), axis=0).astype(np.float32)
    if fv is None:
        raise ValueError(f"unknown featName {featName!r}")
    return fv


def featuregram_from_spec(Spec, l_harm, l_perc, n_mels, featName, use_scipy=True):
    """DAFx12 variant: precomputed magnitude spectrogram in
    (DAFx12_Speech_Music_Detection_B3_MTL_v2.py:230-246)."""
    if featName == 'LogMelSpec':
        fv = lr.melspectrogram(S=Spec, n_mels=n_mels)
        return lr.power_to_db(fv ** 2).astype(np.float32)
    if featName.startswith('LogMelHarm') or featName.startswith('LogMelPerc'):
        H, P = lr.hpss(Spec, kernel_size=(l_harm, l_perc), use_scipy=use_scipy)
        fv_H = lr.power_to_db(lr.melspectrogram(S=H, n_mels=n_mels) ** 2)
        fv_P = lr.power_to_db(lr.melspectrogram(S=P, n_mels=n_mels) ** 2)
        return np.append(fv_H, fv_P, axis=0).astype(np.float32)
    raise ValueError(featName)


# ---- lib/cython_impl/tools.pyx:21-38 ---------------------------------------
def extract_patches(FV, patch_size, patch_shift):
    nFeat, nFrames = FV.shape
    half_win = int(patch_size / 2)
    centres = list(range(half_win, nFrames - half_win, patch_shift))
    patches = np.zeros((len(centres), nFeat, patch_size))            # float64
    for n, i in enumerate(centres):
        frmStart = i - half_win
        frmEnd = min(frmStart + patch_size, nFrames)
        if (frmEnd - frmStart) < patch_size:
            frmStart = frmEnd - patch_size
        patches[n] = FV[:, frmStart:frmEnd]
    return patches


def _standard_scale_rows(FV):
    """StandardScaler(copy=False).fit_transform(FV.T).T on float32 rows, as the sklearn installed
    here (1.9, the only pin available) does it: per-row mean / variance (ddof=0) in float64, constant
    rows keep scale 1, then ``X -= mean.astype(X.dtype); X /= scale.astype(X.dtype)`` in float32.
    (sklearn of the reference's era applied the float64 mean/scale directly; the two differ by at
    most ~2 float32 ulp.)"""
    FV = np.array(FV, dtype=np.float32, copy=True)
    n = FV.shape[1]
    mean = FV.astype(np.float64).mean(axis=1)
    var = FV.astype(np.float64).var(axis=1)
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2          # sklearn _is_constant_feature
    scale = np.sqrt(var)
    scale[constant] = 1.0
    FV -= mean.astype(np.float32)[:, None]
    FV /= scale.astype(np.float32)[:, None]
    return FV


# ---- lib/preprocessing.py:137-292 ------------------------------------------
def get_feature_patches(FV, patch_size, patch_shift, featName, model, frame_level_scaling=False):
    FV = np.asarray(FV)
    if FV.shape[1] < patch_size:                                      # :139-142
        FV1 = FV.copy()
        while FV.shape[1] <= patch_size:
            FV = np.append(FV, FV1, axis=1)

    def one(block):
        if not frame_level_scaling:
            block = _standard_scale_rows(block)
        p = extract_patches(block, patch_size, patch_shift)
        if 'Lemaire_et_al' not in model:
            p = np.expand_dims(p, axis=3)
        return p

    if featName in ('Spec', 'LogSpec', 'MelSpec', 'LogMelSpec'):
        return one(FV)
    half = int(FV.shape[0] / 2)
    stem = featName
    for pre in ('LogMel', 'Mel', 'Log'):
        if stem.startswith(pre):
            stem = stem[len(pre):]
            break
    if stem == 'HarmPercSpec':
        return np.append(one(FV[:half]), one(FV[half:]), axis=1)
    if stem == 'HarmSpec':
        return one(FV[:half]).copy()
    if stem == 'PercSpec':
        return one(FV[half:]).copy()
    raise ValueError(featName)


# ---- lib/preprocessing.py:461-586 (statistics over in-memory featuregrams) --
def get_data_stats(class_to_fvs, classes):
    """``class_to_fvs``: {class_name: [FV (D,T), ...]}; ``classes``: ordered class names.
    Pass 1: per-class frame sums -> class means -> unweighted mean of class means.
    Pass 2: sum (x-mean)^2 over all frames / (N-1), sqrt."""
    sums, counts = {}, {}
    for c in classes:
        s, n = None, 0
        for FV in class_to_fvs[c]:
            FV = FV[~np.isnan(FV).any(axis=1), :]
            FV = FV[~np.isinf(FV).any(axis=1), :]
            FVt = FV.T
            n += FVt.shape[0]
            s = np.sum(FVt, axis=0) if s is None else np.add(s, np.sum(FVt, axis=0))
        sums[c], counts[c] = s, n
    means = [sums[c] / (counts[c] + 1e-10) for c in classes]
    overall_mean = means[0]
    for m in means[1:]:
        overall_mean = np.add(overall_mean, m)
    overall_mean = overall_mean / len(classes)
    stdev, nFrames = None, 0
    for c in classes:
        for FV in class_to_fvs[c]:
            FV = FV[~np.isnan(FV).any(axis=1), :]
            FV = FV[~np.isinf(FV).any(axis=1), :]
            FVt = FV.T
            nFrames += FVt.shape[0]
            mean_arr = np.repeat(np.array(overall_mean, ndmin=2), FVt.shape[0], axis=0)
            d = np.sum(np.power(np.subtract(FVt, mean_arr), 2), axis=0)
            stdev = d if stdev is None else np.add(stdev, d)
    stdev = np.sqrt(stdev / (nFrames - 1))
    return (overall_mean.astype(np.float32), stdev.astype(np.float32),
            *[counts[c] for c in classes])


# ---- lib/preprocessing.py:590-614 and lib/cython_impl/tools.pyx:138-166 -----
def scale_data(FV, mean, stdev):
    M = np.repeat(np.array(mean, ndmin=2).T, FV.shape[1], axis=1)
    S = np.repeat(np.array(stdev, ndmin=2).T, FV.shape[1], axis=1)
    return np.divide(np.subtract(FV.copy(), M), S)


def cscale_data(FV, mean, stdev):
    M = np.repeat(np.array(mean, ndmin=2, dtype=np.float64).T, FV.shape[1], axis=1)
    S = np.repeat(np.array(stdev, ndmin=2, dtype=np.float64).T, FV.shape[1], axis=1)
    return np.divide(np.subtract(FV.copy().astype(np.float64), M), S + 1e-10)
