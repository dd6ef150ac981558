"""CPU restatement of the reference's feature front-end glue (lib/preprocessing.py) and of the signal
preparation in front of it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works on waveforms already in memory.
Function names follow the reference.
"""
from __future__ import annotations

import numpy as np

from . import librosa_restated as lr


# ---- lib/preprocessing.py:114-132 -----------------------------------------
def normalize_signal(Xin):
    Xin = Xin - np.mean(Xin)
    Xin = Xin / np.max(np.abs(Xin))
    return Xin


# ---- librosa.feature.rms(y, frame_length, hop_length, center=True, pad_mode='reflect')[0] (lib/preprocessing.py:337)
def frame_rms(y, frame_length, hop_length):
    yp = np.pad(y, int(frame_length // 2), mode='reflect')
    n = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n)[None, :]
    x = yp[idx]
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0))


# ---- lib/cython_impl/tools.pyx:42-134 (the Cython leaf load_and_preprocess_signal calls) -------------------
def removeSilence(Xin, nSamples, energy, nFrames, fs, Tw, Ts, alpha=0.025, beta=0.075):
    """Quirks kept: the energy threshold is a C float, nothing is removed unless MORE than one silent stretch
    qualifies, and the returned signal keeps its original length -- the kept samples are packed to the front of a
    float32 buffer of ones."""
    from scipy.signal import medfilt
    frameSize = int((Tw * fs) / 1000)
    frameShift = int((Ts * fs) / 1000)
    thresh = np.float32(alpha * np.max(energy))
    marker = (np.asarray(energy) >= thresh).astype(np.float64)
    marker = (medfilt(marker, 5) > 0.5).astype(np.int64)
    sample_marker = np.ones(nSamples, dtype=np.int64)
    total, nSil, i = 0, 0, 0
    last = nFrames - 1
    while i < nFrames:
        # i: first silent frame at/after i (or the last frame); j: first active frame after it (or the last)
        nz = np.flatnonzero(marker[i:] == 0)
        i = i + int(nz[0]) if nz.size else last
        nz = np.flatnonzero(marker[i:] == 1)
        j = i + int(nz[0]) if nz.size else last
        k = max(frameShift * (i - 1) + frameSize, 1)
        l = min(frameShift * (j - 1) + frameSize, nSamples)
        if (l - k) / fs > beta:
            sample_marker[k:l] = 0
            nSil += 1
            total += int((l - k) / fs)          # the reference accumulates into a C int
        i = j + 1
    if nSil > 1:
        keep = np.flatnonzero(sample_marker == 1)
        out = np.ones(nSamples, dtype=np.float32)
        out[:keep.size] = Xin[keep]
    else:
        out = Xin
    return out, sample_marker, marker, total


# ---- lib/preprocessing.py:297-325 ---------------------------------------------------------------------------
def mix_signals(Xin_sp, Xin_mu, target_dB):
    n_sp = len(Xin_sp)
    reps = int(np.ceil(n_sp / len(Xin_mu))) if len(Xin_mu) < n_sp else 1
    mu = np.tile(Xin_mu, reps) if reps > 1 else Xin_mu.copy()
    common = min(n_sp, len(mu))
    sp, mu = Xin_sp[:common], mu[:common]
    e_sp = np.sum(np.power(sp, 2)) / len(sp)
    e_mu = np.sum(np.power(mu, 2)) / len(mu)
    g_mu = np.sqrt((e_sp / np.power(10, (target_dB / 10))) / e_mu)
    g_sp = 1
    tot = g_mu + g_sp
    g_mu /= tot
    g_sp /= tot
    return normalize_signal(g_sp * sp + g_mu * mu)


# ---- lib/preprocessing.py:330-350 after the decode ----------------------------------------------------------
def load_and_preprocess_signal(Xin, Tw, Ts, fs=16000, details=False):
    Xin = normalize_signal(np.asarray(Xin, dtype=np.float32))
    frameSize = int((Tw * fs) / 1000)
    frameShift = int((Ts * fs) / 1000)
    energy = frame_rms(Xin, frameSize, frameShift)
    silrem, sample_marker, frame_marker, _ = removeSilence(Xin, len(Xin), energy, len(energy), fs, Tw, Ts)
    out = silrem.copy()
    if len(out) / fs < 0.1:
        while len(out) / fs < 0.1:
            out = np.append(out, out)
    out = normalize_signal(out)
    if details:
        return out, sample_marker, frame_marker, energy
    return out


# ---- lib/cython_impl/tools.pyx:169-211 ----------------------------------------------------------------------
def get_data_statistics(FV, stat_type='skew', axis=0):
    from scipy.stats import kurtosis, skew
    fn = {'mean': np.mean, 'variance': np.var, 'skew': skew, 'kurtosis': kurtosis}[stat_type]
    return np.stack([fn(np.squeeze(FV[i]), axis=axis) for i in range(FV.shape[0])]).astype(np.float64)


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
