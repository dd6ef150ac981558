"""Drop-in mirror of the reference's ``lib/preprocessing.py`` feature front-end.

Same function names, argument meaning, return layout and dtypes as
/root/reference/lib/preprocessing.py (get_featuregram :355, get_feature_patches :137,
get_data_stats :461, scale_data :590) and lib/cython_impl/tools.pyx (scale_data :138,
extract_patches :21).  All feature arithmetic (STFT, HPSS medians, soft masks, mel, power_to_db,
row standardisation, patch gather, moments) runs in libhpss_b200.so on the GPU; there is no CPU
fallback.  What stays on the host is file I/O, the ``.npy`` feature cache and the signal
preparation that precedes the hot path (normalise / silence removal / SMR mixing).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import ParameterError

FS = 16000


# ============================================================================ feature names
def feature_family(featName: str) -> Tuple[str, bool]:
    """Reference name dispatch (lib/preprocessing.py:378-444) -> (library feature id, mel basis uses fs).

    The reference tests ``==`` for the four plain names and ``startswith`` for the HPSS families, so
    e.g. 'LogMelHarmSpec', 'LogMelPercSpec' and 'LogMelHarmPercSpec' all compute both streams."""
    if featName == 'Spec':
        return 'SPEC', False
    if featName == 'LogSpec':
        return 'LOGSPEC', False
    if featName == 'MelSpec':
        return 'MELSPEC', True
    if featName == 'LogMelSpec':
        return 'LOGMELSPEC', True
    if featName.startswith('MelHarm') or featName.startswith('MelPerc'):
        return 'MEL_HARMPERC', False
    if featName.startswith('LogMelHarm') or featName.startswith('LogMelPerc'):
        return 'LOGMEL_HARMPERC', False
    if featName.startswith('Harm') or featName.startswith('Perc'):
        return 'HARMPERC', False
    if featName.startswith('LogHarm') or featName.startswith('LogPerc'):
        return 'LOG_HARMPERC', False
    raise ValueError(f"unknown featName {featName!r} (the reference leaves 'fv' unbound here)")


def _kernel_sizes(PARAMS) -> Tuple[int, int]:
    lh, lp = PARAMS.get('l_harm', 31), PARAMS.get('l_perc', 31)
    if isinstance(lh, dict):
        lh = lh[PARAMS['Model']]
    if isinstance(lp, dict):
        lp = lp[PARAMS['Model']]
    return int(lh), int(lp)


def make_params(PARAMS, fs, n_fft, n_mels, featName):
    from . import engine
    fam, mel_uses_fs = feature_family(featName)
    frameSize = int(PARAMS['Tw'] * fs / 1000)
    frameShift = int(PARAMS['Ts'] * fs / 1000)
    lh, lp = _kernel_sizes(PARAMS) if 'HARMPERC' in fam else (1, 1)
    return engine.make_params(n_fft=n_fft, win_length=frameSize, hop_length=frameShift, l_harm=lh, l_perc=lp,
                              n_mels=max(int(n_mels), 1), mel_sr=int(fs) if mel_uses_fs else 22050, feature=fam)


# ============================================================================ GPU feature extraction
def featuregram_batch(signals: Sequence[np.ndarray], fs: int, PARAMS, n_fft: int, n_mels: int, featName: str,
                      device: Optional[int] = None) -> List[np.ndarray]:
    """Featuregrams of many already-prepared signals in one batched GPU call.

    Returns one float32 (nFeat, T_c) array per signal, exactly what get_featuregram returns per file."""
    import torch
    from . import engine
    sigs = [np.ascontiguousarray(x, dtype=np.float32) for x in signals]
    for x in sigs:
        if x.ndim != 1:
            raise ParameterError("only mono input is on the reference path")
        if not np.isfinite(x).all():
            raise ParameterError("Audio buffer is not finite everywhere")      # librosa.util.valid_audio
    ctx = engine.get_context(device)
    prm = make_params(PARAMS, fs, n_fft, n_mels, featName)
    batch = engine.Batch(ctx, clip_lengths=[len(x) for x in sigs], n_fft=n_fft, hop_length=prm.hop_length)
    rows = engine.feature_rows(prm)
    with torch.cuda.device(ctx.device):
        wave = torch.from_numpy(np.concatenate(sigs) if sigs else np.zeros(0, np.float32)).cuda()
        out = engine.featuregram(batch, wave, prm)
        host = out.cpu().numpy()
    res = []
    for c in range(batch.n_clips):
        a, b = rows * int(batch.frame_offsets[c]), rows * int(batch.frame_offsets[c + 1])
        res.append(host[a:b].reshape(rows, -1).copy())
    batch.close()
    return res


def featuregram_from_signal(Xin: np.ndarray, fs: int, PARAMS, n_fft: int, n_mels: int, featName: str) -> np.ndarray:
    """Body of get_featuregram after the signal is loaded (lib/preprocessing.py:378-444)."""
    return featuregram_batch([Xin], fs, PARAMS, n_fft, n_mels, featName)[0]


def featuregram_from_spec(Spec: np.ndarray, PARAMS, n_mels: int, featName: str) -> np.ndarray:
    """DAFx12 variant: precomputed magnitude spectrogram (DAFx12_..._v2.py:230-246)."""
    import torch
    from . import engine
    Spec = np.ascontiguousarray(Spec, dtype=np.float32)
    if (Spec < 0).any():
        raise ParameterError("X and X_ref must be non-negative")             # librosa.util.softmask
    rows, T = Spec.shape
    if featName == 'LogMelSpec':
        fam, sr = 'LOGMELSPEC', 22050     # melspectrogram(S=Spec): default sr; S is used as is (no squaring)
    elif featName.startswith('LogMelHarm') or featName.startswith('LogMelPerc'):
        fam, sr = 'LOGMEL_HARMPERC', 22050
    else:
        raise ValueError(featName)
    ctx = engine.get_context()
    lh, lp = _kernel_sizes(PARAMS)
    prm = engine.make_params(n_fft=2 * (rows - 1), l_harm=lh, l_perc=lp, n_mels=n_mels, mel_sr=sr, feature=fam)
    batch = engine.Batch(ctx, clip_frames=[T])
    S = torch.from_numpy(Spec.ravel()).cuda()
    if fam == 'LOGMELSPEC':
        mel = torch.from_numpy(engine.mel_filterbank(sr, 2 * (rows - 1), n_mels)).cuda()
        out, cmax = engine.mask_mel_log(batch, S, None, None, rows, mel=mel, pre_square=False, log_power=True)
        engine.topdb_clip(batch, out, n_mels, 1, cmax, 80.0)
        res = out.cpu().numpy().reshape(n_mels, T)
    else:
        res = engine.featuregram_from_spec(batch, S, rows, prm).cpu().numpy().reshape(2 * n_mels, T)
    batch.close()
    return res


# ============================================================================ signal preparation (host)
def normalize_signal(Xin):
    """lib/preprocessing.py:114-132."""
    Xin = Xin - np.mean(Xin)
    Xin = Xin / np.max(np.abs(Xin))
    return Xin


def _frame_rms(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.feature.rms(y=..., center=True, pad_mode='reflect')[0]."""
    yp = np.pad(y, int(frame_length // 2), mode='reflect')
    n = 1 + (len(yp) - frame_length) // hop_length
    # sliding sum of squares via a cumulative sum would change rounding; frame explicitly like librosa
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n)[None, :]
    x = yp[idx]
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0))


def removeSilence(Xin, nSamples, energy, nFrames, fs, Tw, Ts, alpha=0.025, beta=0.075):
    """Silence excision with the semantics of the reference's Cython leaf
    (lib/cython_impl/tools.pyx:42-134), including its quirks: the energy threshold is a float32,
    nothing is removed unless MORE than one silent stretch qualifies, and the returned signal keeps
    its original length -- the kept samples are packed to the front of a float32 buffer of ones."""
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


def mix_signals(Xin_sp, Xin_mu, target_dB):
    """lib/preprocessing.py:297-325: loop the music up to the speech length, scale it to the target
    speech-to-music ratio, weight both so the factors sum to one, normalise."""
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


def load_audio(fName: str, sr: int = FS) -> np.ndarray:
    """Mono float32 audio at ``sr`` (stand-in for librosa.core.load(fName, mono=True, sr=16000)).
    Reads RIFF/WAVE PCM (scipy.io.wavfile) and ``.npy`` waveforms; MUSAN is 16 kHz wav, which
    librosa returns as int16 / 32768.  Other sample rates are resampled with scipy's polyphase
    filter, which is NOT bit-identical to librosa's resampy kernel."""
    if fName.endswith('.npy'):
        x = np.load(fName)
        rate = sr
    else:
        from scipy.io import wavfile
        rate, x = wavfile.read(fName)
        if x.dtype == np.int16:
            x = x.astype(np.float32) / 32768.0
        elif x.dtype == np.int32:
            x = x.astype(np.float32) / 2147483648.0
        elif x.dtype == np.uint8:
            x = (x.astype(np.float32) - 128.0) / 128.0
        else:
            x = x.astype(np.float32)
    if x.ndim > 1:
        x = np.mean(x, axis=1 if x.shape[1] < x.shape[0] else 0)
    if rate != sr:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(rate), int(sr))
        x = resample_poly(x, sr // g, rate // g)
    return np.ascontiguousarray(x, dtype=np.float32)


def load_and_preprocess_signal(fName, Tw, Ts, loader=load_audio):
    """lib/preprocessing.py:330-350."""
    Xin = loader(fName)
    fs = FS
    Xin = normalize_signal(Xin)
    frameSize = int((Tw * fs) / 1000)
    frameShift = int((Ts * fs) / 1000)
    energy = _frame_rms(Xin, frameSize, frameShift)
    Xin_silrem, _, _, _ = removeSilence(Xin, len(Xin), energy, len(energy), fs, Tw, Ts)
    Xin = Xin_silrem.copy()
    if len(Xin) / fs < 0.1:
        while len(Xin) / fs < 0.1:
            Xin = np.append(Xin, Xin)
    return normalize_signal(Xin), fs


# ============================================================================ get_featuregram
def _feature_name_of_file(fName_path_sp, fName_path_mu, target_dB):
    if (fName_path_sp != '') and (fName_path_mu != ''):
        return (fName_path_sp.split('/')[-1].split('.')[0] + '_' + fName_path_mu.split('/')[-1].split('.')[0]
                + '_' + str(target_dB) + 'dB')
    if fName_path_sp != '':
        return fName_path_sp.split('/')[-1].split('.')[0]
    if fName_path_mu != '':
        return fName_path_mu.split('/')[-1].split('.')[0]
    raise ValueError("both file names are empty")


def prepare_signal(PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader=load_audio):
    """Signal of one get_featuregram call (lib/preprocessing.py:364-376)."""
    if classname == 'speech_music':
        sp, fs = load_and_preprocess_signal(fName_path_sp, PARAMS['Tw'], PARAMS['Ts'], loader)
        mu, fs = load_and_preprocess_signal(fName_path_mu, PARAMS['Tw'], PARAMS['Ts'], loader)
        return mix_signals(sp, mu, target_dB), fs
    if classname in ('speech', 'muspeak'):
        return load_and_preprocess_signal(fName_path_sp, PARAMS['Tw'], PARAMS['Ts'], loader)
    if classname == 'music':
        return load_and_preprocess_signal(fName_path_mu, PARAMS['Tw'], PARAMS['Ts'], loader)
    raise ValueError(f"unknown classname {classname!r}")


def get_featuregram(PARAMS, classname, feature_opDir, fName_path_sp, fName_path_mu, target_dB, n_fft, n_mels,
                    featName, save_feat=True, loader=load_audio):
    """Same signature, cache layout and return value as lib/preprocessing.py:355-457."""
    fName = _feature_name_of_file(fName_path_sp, fName_path_mu, target_dB)
    cache = feature_opDir + '/' + classname + '/' + fName + '.npy'
    if not os.path.exists(cache):
        Xin, fs = prepare_signal(PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader)
        fv = featuregram_from_signal(Xin, fs, PARAMS, n_fft, n_mels, featName)
        if save_feat:
            os.makedirs(feature_opDir + '/' + classname + '/', exist_ok=True)
            np.save(cache, fv)
    else:
        try:
            fv = np.load(cache, allow_pickle=True)
        except Exception:
            print('Error loading: ', cache)
            fv = np.load(cache, allow_pickle=True)
    return fv


# ============================================================================ patches
def _stem(featName: str) -> str:
    for pre in ('LogMel', 'Mel', 'Log'):
        if featName.startswith(pre) and featName != pre + 'Spec':
            return featName[len(pre):]
    return featName


def get_feature_patches(PARAMS, FV, patch_size, patch_shift, featName):
    """lib/preprocessing.py:137-292 on the GPU: per-file StandardScaler of every feature row (unless
    frame_level_scaling), patch gather -> float64 (nPatch, nFeat, W[, 1]); harmonic/percussive halves
    selected or stacked on axis 1 by the feature name."""
    import torch
    from . import engine
    FV = np.asarray(FV)
    if FV.shape[1] < patch_size:                                   # :139-142
        FV1 = FV.copy()
        while FV.shape[1] <= patch_size:
            FV = np.append(FV, FV1, axis=1)
    plain = featName in ('Spec', 'LogSpec', 'MelSpec', 'LogMelSpec')
    stem = _stem(featName)
    if not plain and stem not in ('HarmSpec', 'PercSpec', 'HarmPercSpec'):
        raise ValueError(f"unknown featName {featName!r}")
    half = int(FV.shape[0] / 2)
    if plain or stem == 'HarmPercSpec':
        block = FV                       # row standardisation is per row: both halves at once
    elif stem == 'HarmSpec':
        block = FV[:half]
    else:
        block = FV[half:]
    ctx = engine.get_context()
    D, T = block.shape
    feat = torch.from_numpy(np.ascontiguousarray(block, dtype=np.float32)).cuda()
    if not PARAMS['frame_level_scaling']:
        batch = engine.Batch(ctx, clip_frames=[T])
        engine.row_standardize(batch, feat.view(-1), D)
        batch.close()
    patches = engine.extract_patches(ctx, feat, patch_size, patch_shift).cpu().numpy()
    if 'Lemaire_et_al' not in PARAMS['Model']:
        patches = np.expand_dims(patches, axis=3)
    return patches


# ============================================================================ global statistics
def scale_data(FV, mean, stdev):
    """lib/preprocessing.py:590-614 ((x - mean) / stdev, numpy promotion of the inputs)."""
    return _scale(FV, mean, stdev, 0.0, np.result_type(np.asarray(FV).dtype, np.asarray(mean).dtype))


def cscale_data(FV, mean, stdev):
    """lib/cython_impl/tools.pyx:138-166 ((x - mean) / (stdev + 1e-10) in float64)."""
    return _scale(FV, mean, stdev, 1e-10, np.float64)


def _scale(FV, mean, stdev, eps, out_dtype):
    import torch
    from . import engine
    FV = np.ascontiguousarray(FV, dtype=np.float32)
    D, T = FV.shape
    ctx = engine.get_context()
    batch = engine.Batch(ctx, clip_frames=[T])
    out = engine.scale_data(batch, torch.from_numpy(FV.ravel()).cuda(), D,
                            torch.from_numpy(np.ascontiguousarray(mean, dtype=np.float32)).cuda(),
                            torch.from_numpy(np.ascontiguousarray(stdev, dtype=np.float32)).cuda(), eps)
    res = out.cpu().numpy().reshape(D, T)
    batch.close()
    return res.astype(out_dtype, copy=False)


def data_stats_from_featuregrams(class_to_fvs: Dict[str, List[np.ndarray]], classes: Sequence[str],
                                 group=None, chunk_frames: int = 1 << 20):
    """Global feature mean / stdev of get_data_stats (lib/preprocessing.py:461-586) from featuregrams held
    in memory: raw float64 moments per batch on the GPU, SUM all-reduce over the process group (the only
    collective on the path), closed-form finish.  Returns (mean f32[D], stdev f32[D], *per-class counts)."""
    import torch
    from . import engine
    from .dist import allreduce_moments, finalize_stats
    ctx = engine.get_context()
    n_classes = len(classes)
    D = None
    acc = None
    pend, pend_cls, pend_frames = [], [], 0

    def flush():
        nonlocal acc, pend, pend_cls, pend_frames
        if not pend:
            return
        batch = engine.Batch(ctx, clip_frames=[fv.shape[1] for fv in pend])
        flat = torch.from_numpy(np.concatenate([np.ascontiguousarray(fv, dtype=np.float32).ravel() for fv in pend])).cuda()
        acc = engine.moments(batch, flat, D, pend_cls, n_classes, acc=acc)
        torch.cuda.synchronize()
        batch.close()
        pend, pend_cls, pend_frames = [], [], 0

    for k, name in enumerate(classes):
        for fv in class_to_fvs.get(name, []):
            if D is None:
                D = fv.shape[0]
            pend.append(fv)
            pend_cls.append(k)
            pend_frames += fv.shape[1]
            if pend_frames >= chunk_frames:
                flush()
    flush()
    if acc is None:
        if D is None:
            raise ValueError("no featuregrams given")
        acc = torch.zeros(n_classes * D + D + n_classes + 1, dtype=torch.float64, device='cuda')
    allreduce_moments(acc, group)
    mean, std, counts = finalize_stats(acc.cpu().numpy(), D, n_classes)
    return (mean, std, *[int(c) for c in counts])


def get_data_stats(PARAMS, files, loader=load_audio):
    """Same signature and return value as lib/preprocessing.py:461-586: every file of every class goes
    through get_featuregram (cached .npy or computed), statistics as in data_stats_from_featuregrams."""
    classes = PARAMS['classes']
    folder = PARAMS['feature_opDir']
    model = PARAMS['Model']
    featName, n_fft, n_mels = PARAMS['featName'][model], PARAMS['n_fft'][model], PARAMS['n_mels'][model]
    names = [classes[k] for k in classes.keys()]
    groups: Dict[str, List[np.ndarray]] = {n: [] for n in names}
    for name in names:
        file_list = files['speech+music'] if name == 'speech_music' else files[name]
        for fl in file_list:
            if name == 'speech_music':
                sp = PARAMS['folder'] + '/speech/' + fl['speech']
                mu = PARAMS['folder'] + '/music/' + fl['music']
                FV = get_featuregram(PARAMS, 'speech_music', folder, sp, mu, fl['SMR'], n_fft, n_mels, featName,
                                     loader=loader)
            elif name == 'music':
                mu = PARAMS['folder'] + '/music/' + fl.split('.')[0] + '.wav'
                FV = get_featuregram(PARAMS, 'music', folder, '', mu, -1, n_fft, n_mels, featName, loader=loader)
            else:
                sp = PARAMS['folder'] + '/' + name + '/' + fl.split('.')[0] + '.wav'
                FV = get_featuregram(PARAMS, name, folder, sp, '', -1, n_fft, n_mels, featName, loader=loader)
            groups[name].append(FV)
    return data_stats_from_featuregrams(groups, names)
