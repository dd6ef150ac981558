"""Drop-in mirror of the reference's ``lib/preprocessing.py`` feature front-end.

Same function names, argument meaning, return layout and dtypes as
/root/reference/lib/preprocessing.py (get_featuregram :355, get_feature_patches :137,
get_data_stats :461, scale_data :590) and lib/cython_impl/tools.pyx (scale_data :138,
extract_patches :21).  All feature arithmetic (STFT, HPSS medians, soft masks, mel, power_to_db,
row standardisation, patch gather, moments) and the signal preparation in front of it (normalise, RMS gate,
silence removal, SMR mixing: SURVEY.md row N2) run in libhpss_b200.so on the GPU; there is no CPU
fallback.  What stays on the host is decoding files, the ``.npy`` feature cache and the statistics pickle.
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


# ============================================================================ host staging
_PINNED = {}


def _pinned(n: int, dtype, slot: str) -> np.ndarray:
    """Grow-only pinned staging buffers (one per role), so H2D / D2H run at link rate without per-call
    cudaHostAlloc."""
    from . import engine
    dt = np.dtype(dtype)
    key = (slot, dt.str)
    buf = _PINNED.get(key)
    if buf is None or buf.size < n:
        buf = _PINNED[key] = engine.host_alloc(max(int(n * 1.25), 1 << 16), dt)
    return buf[:n]


_PIPES = {}


def _pipeline(ctx, lengths, prm, pcm_dtype, prepare, fs=FS):
    """Small cache of host pipelines keyed by (clip lengths, parameters): repeated calls with the same shapes
    (mini-batches of equal segments, benchmark loops) reuse chunk layouts and device slots."""
    from . import engine
    key = (ctx.device, tuple(int(x) for x in lengths), bytes(prm), np.dtype(pcm_dtype).str, bool(prepare), int(fs))
    pl = _PIPES.get(key)
    if pl is None:
        if len(_PIPES) >= 8:
            _PIPES.pop(next(iter(_PIPES))).close()
        pl = _PIPES[key] = engine.Pipeline(ctx, lengths, prm, pcm_dtype=pcm_dtype, prepare=prepare, fs=fs)
    return pl


# ============================================================================ GPU feature extraction
def featuregram_batch(signals: Sequence[np.ndarray], fs: int, PARAMS, n_fft: int, n_mels: int, featName: str,
                      device: Optional[int] = None) -> List[np.ndarray]:
    """Featuregrams of many already-prepared signals in one batched GPU call (pinned staging, H2D / kernels /
    D2H pipelined by hpss_pipeline_run).

    Returns one float32 (nFeat, T_c) array per signal, exactly what get_featuregram returns per file."""
    from . import engine
    sigs = [np.asarray(x) for x in signals]
    for x in sigs:
        if x.ndim != 1:
            raise ParameterError("only mono input is on the reference path")
    if not sigs:
        return []
    ctx = engine.get_context(device)
    prm = make_params(PARAMS, fs, n_fft, n_mels, featName)
    lengths = [len(x) for x in sigs]
    pl = _pipeline(ctx, lengths, prm, np.float32, False)
    wave = _pinned(sum(lengths), np.float32, "wave")
    o = 0
    for x in sigs:
        wave[o:o + len(x)] = x                                           # float64 mixes are rounded here, once
        o += len(x)
    rows = pl.rows
    out = _pinned(rows * pl.total_frames, np.float32, "feat")
    pl.run(wave, feat_host=out)                                          # raises ParameterError on non-finite audio
    res = []
    for c in range(pl.n_clips):
        a, b = rows * int(pl.frame_offsets[c]), rows * int(pl.frame_offsets[c + 1])
        res.append(out[a:b].reshape(rows, -1).copy())
    return res


def featuregram_from_signal(Xin: np.ndarray, fs: int, PARAMS, n_fft: int, n_mels: int, featName: str) -> np.ndarray:
    """Body of get_featuregram after the signal is loaded (lib/preprocessing.py:378-444)."""
    return featuregram_batch([Xin], fs, PARAMS, n_fft, n_mels, featName)[0]


def featuregram_from_spec(Spec: np.ndarray, PARAMS, n_mels: int, featName: str) -> np.ndarray:
    """DAFx12 variant: precomputed magnitude spectrogram (DAFx12_..._v2.py:230-246)."""
    import torch
    from . import engine
    Spec = np.ascontiguousarray(Spec, dtype=np.float32)
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
        out = engine.featuregram_from_spec(batch, S, rows, prm)         # flags negative input (softmask raises)
        engine.ctx_check(ctx)
        res = out.cpu().numpy().reshape(2 * n_mels, T)
    batch.close()
    return res


# ============================================================================ signal preparation (GPU, row N2)
def normalize_signal(Xin):
    """lib/preprocessing.py:114-132 (two numpy reductions; the feature path itself normalises on the device)."""
    Xin = Xin - np.mean(Xin)
    Xin = Xin / np.max(np.abs(Xin))
    return Xin


def _as_pcm(x) -> np.ndarray:
    """Decoded audio as the device accepts it: int16 PCM stays int16 (half the upload; librosa's x / 32768 is
    applied on the device), everything else becomes float32."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ParameterError("only mono input is on the reference path")
    if x.dtype != np.int16:
        x = x.astype(np.float32, copy=False)
    return np.ascontiguousarray(x)


def prepare_signals_device(ctx, pcms: Sequence[np.ndarray], Tw, Ts, fs: int = FS, markers: bool = False):
    """load_and_preprocess_signal after the decode for a list of files, on the device (engine.prep_signals).
    Returns (flat CUDA float32 tensor, per-file output lengths[, frame markers, sample markers, n_sil])."""
    import torch
    from . import engine
    pcms = [_as_pcm(x) for x in pcms]
    if len({x.dtype for x in pcms}) > 1:
        pcms = [x.astype(np.float32) / np.float32(32768.0) if x.dtype == np.int16 else x for x in pcms]
    lengths = [len(x) for x in pcms]
    host = _pinned(sum(lengths), pcms[0].dtype, "pcm")
    o = 0
    for x in pcms:
        host[o:o + len(x)] = x
        o += len(x)
    with torch.cuda.device(ctx.device):
        dev = torch.from_numpy(host).cuda(non_blocking=True)
        res = engine.prep_signals(ctx, dev, lengths, fs=fs, win_length=int((Tw * fs) / 1000),
                                  hop_length=int((Ts * fs) / 1000), markers=markers)
        # the pinned buffer may be reused by the next call: the upload has to be through
        torch.cuda.current_stream().synchronize()
    return res


def load_and_preprocess_signal(fName, Tw, Ts, loader=None):
    """lib/preprocessing.py:330-350: decode on the host, everything else (normalise, RMS gate, silence removal with the
    Cython leaf's semantics, doubling below 0.1 s, normalise) on the GPU."""
    from . import engine
    Xin = (loader or load_pcm)(fName)
    ctx = engine.get_context()
    out, _ = prepare_signals_device(ctx, [Xin], Tw, Ts)
    engine.ctx_check(ctx)
    return out.cpu().numpy(), FS


def mix_signals(Xin_sp, Xin_mu, target_dB):
    """lib/preprocessing.py:297-325 on the GPU (hpss_mix_signals): the music is looped up to the speech length and
    scaled to the target speech-to-music ratio, both weights divided by their sum, the mix normalised."""
    import torch
    from . import engine
    ctx = engine.get_context()
    sp = torch.from_numpy(np.ascontiguousarray(Xin_sp, dtype=np.float32)).cuda()
    mu = torch.from_numpy(np.ascontiguousarray(Xin_mu, dtype=np.float32)).cuda()
    return engine.mix_signals(ctx, sp, [sp.numel()], mu, [mu.numel()], [float(target_dB)]).cpu().numpy()


def load_pcm(fName: str, sr: int = FS) -> np.ndarray:
    """Decode one file to mono samples at ``sr``: int16 for 16-bit PCM wav at the right rate (MUSAN), float32
    otherwise (stand-in for librosa.core.load(fName, mono=True, sr=16000); ``.npy`` waveforms are accepted too).
    Other sample rates are resampled with scipy's polyphase filter, which is NOT bit-identical to librosa's
    resampy kernel."""
    if fName.endswith('.npy'):
        x = np.load(fName)
        rate = sr
    else:
        from scipy.io import wavfile
        rate, x = wavfile.read(fName)
    if x.dtype == np.int16 and x.ndim == 1 and rate == sr:
        return np.ascontiguousarray(x)
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


def load_audio(fName: str, sr: int = FS) -> np.ndarray:
    """Mono float32 audio at ``sr`` (what librosa.core.load returns)."""
    x = load_pcm(fName, sr)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / np.float32(32768.0)
    return x


# ============================================================================ get_featuregram
def _stem_of(path):
    return path.split('/')[-1].split('.')[0]


def _feature_name_of_file(fName_path_sp, fName_path_mu, target_dB, fName_path_no=''):
    """Cache-file stem: lib/preprocessing.py:356-361; with the noise path, 5_class_classification.py:315-324."""
    if (fName_path_sp != '') and (fName_path_mu != ''):
        return _stem_of(fName_path_sp) + '_' + _stem_of(fName_path_mu) + '_' + str(target_dB) + 'dB'
    if (fName_path_sp != '') and (fName_path_no != ''):
        return _stem_of(fName_path_sp) + '_' + _stem_of(fName_path_no) + '_' + str(target_dB) + 'dB'
    if fName_path_sp != '':
        return _stem_of(fName_path_sp)
    if fName_path_mu != '':
        return _stem_of(fName_path_mu)
    if fName_path_no != '':
        return _stem_of(fName_path_no)
    raise ValueError("all file names are empty")


def _sources(classname, fName_path_sp, fName_path_mu, fName_path_no=''):
    """(first file, second file or None) of a class (lib/preprocessing.py:364-376; 5_class_classification.py:327-349)."""
    if classname == 'speech_music':
        return fName_path_sp, fName_path_mu
    if classname == 'speech_noise':
        return fName_path_sp, fName_path_no
    if classname in ('speech', 'muspeak'):
        return fName_path_sp, None
    if classname == 'music':
        return fName_path_mu, None
    if classname == 'noise':
        return fName_path_no, None
    raise ValueError(f"unknown classname {classname!r}")


def prepare_signal_device(ctx, PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader=None,
                          fName_path_no=''):
    """Prepared (and, for the mixed classes, mixed) signal of one get_featuregram call as a CUDA float32 tensor."""
    from . import engine
    loader = loader or load_pcm
    first, second = _sources(classname, fName_path_sp, fName_path_mu, fName_path_no)
    pcms = [loader(first)] + ([loader(second)] if second is not None else [])
    flat, lens = prepare_signals_device(ctx, pcms, PARAMS['Tw'], PARAMS['Ts'])
    if second is None:
        return flat
    return engine.mix_signals(ctx, flat[:lens[0]], [lens[0]], flat[lens[0]:], [lens[1]], [float(target_dB)])


def prepare_signal(PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader=None, fName_path_no=''):
    """Signal of one get_featuregram call (lib/preprocessing.py:364-376) as a numpy array."""
    from . import engine
    ctx = engine.get_context()
    x = prepare_signal_device(ctx, PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader, fName_path_no)
    engine.ctx_check(ctx)
    return x.cpu().numpy(), FS


def _featuregram_of_files(PARAMS, classname, fName_path_sp, fName_path_mu, fName_path_no, target_dB, n_fft, n_mels,
                          featName, loader):
    """decode (host) -> upload -> prepare [-> mix] -> features, one download of the result."""
    import torch
    from . import engine
    ctx = engine.get_context()
    with torch.cuda.device(ctx.device):
        wave = prepare_signal_device(ctx, PARAMS, classname, fName_path_sp, fName_path_mu, target_dB, loader,
                                     fName_path_no)
        prm = make_params(PARAMS, FS, n_fft, n_mels, featName)
        batch = engine.Batch(ctx, clip_lengths=[wave.numel()], n_fft=n_fft, hop_length=prm.hop_length)
        out = engine.featuregram(batch, wave, prm)
        rows = engine.feature_rows(prm)
        host = _pinned(out.numel(), np.float32, "feat1")
        torch.from_numpy(host).copy_(out, non_blocking=True)
        engine.ctx_check(ctx)                        # synchronises; raises like librosa on non-finite audio
        batch.close()
    return host.reshape(rows, -1).copy()


def _cached_featuregram(cache, compute, save_feat):
    if not os.path.exists(cache):
        fv = compute()
        if save_feat:
            os.makedirs(os.path.dirname(cache) + '/', exist_ok=True)
            np.save(cache, fv)
    else:
        try:
            fv = np.load(cache, allow_pickle=True)
        except Exception:
            print('Error loading: ', cache)
            fv = np.load(cache, allow_pickle=True)
    return fv


def get_featuregram(PARAMS, classname, feature_opDir, fName_path_sp, fName_path_mu, target_dB, n_fft, n_mels,
                    featName, save_feat=True, loader=None):
    """Same signature, cache layout and return value as lib/preprocessing.py:355-457."""
    fName = _feature_name_of_file(fName_path_sp, fName_path_mu, target_dB)
    cache = feature_opDir + '/' + classname + '/' + fName + '.npy'
    return _cached_featuregram(cache, lambda: _featuregram_of_files(
        PARAMS, classname, fName_path_sp, fName_path_mu, '', target_dB, n_fft, n_mels, featName, loader), save_feat)


def get_featuregram_5class(PARAMS, classname, feature_opDir, fName_path_sp, fName_path_mu, fName_path_no, target_dB,
                           n_fft, n_mels, featName, save_feat=True, loader=None):
    """The 5-class script's copy of get_featuregram (5_class_classification.py:314-385): one more positional, the
    noise file; classes 'noise' and 'speech_noise' (speech mixed with noise at target_dB) next to the other three.
    Only 'LogMelSpec' and the 'LogMelHarm*' / 'LogMelPerc*' names exist there."""
    if not (featName == 'LogMelSpec' or featName.startswith('LogMelHarm') or featName.startswith('LogMelPerc')):
        raise ValueError(f"featName {featName!r} is not computed by the 5-class script")
    fName = _feature_name_of_file(fName_path_sp, fName_path_mu, target_dB, fName_path_no)
    cache = feature_opDir + '/' + classname + '/' + fName + '.npy'
    return _cached_featuregram(cache, lambda: _featuregram_of_files(
        PARAMS, classname, fName_path_sp, fName_path_mu, fName_path_no, target_dB, n_fft, n_mels, featName, loader),
        save_feat)


def get_featuregram_from_spec(PARAMS, feature_opDir, fName, Spec, n_fft, n_mels, featName, save_feat=True):
    """The long-form script's variant (DAFx12_Speech_Music_Detection_B3_MTL_v2.py:230-255): a precomputed whole-file
    magnitude spectrogram in, cache file feature_opDir/fName.npy."""
    cache = feature_opDir + '/' + fName + '.npy'
    return _cached_featuregram(cache, lambda: featuregram_from_spec(Spec, PARAMS, n_mels, featName), save_feat)


# ============================================================================ patches
def _stem(featName: str) -> str:
    for pre in ('LogMel', 'Mel', 'Log'):
        if featName.startswith(pre) and featName != pre + 'Spec':
            return featName[len(pre):]
    return featName


def _patch_rows(featName: str) -> str:
    plain = featName in ('Spec', 'LogSpec', 'MelSpec', 'LogMelSpec')
    stem = _stem(featName)
    if not plain and stem not in ('HarmSpec', 'PercSpec', 'HarmPercSpec'):
        raise ValueError(f"unknown featName {featName!r}")
    if plain or stem == 'HarmPercSpec':
        return "all"                      # row standardisation is per row: both halves at once
    return "harm" if stem == 'HarmSpec' else "perc"


def feature_patches_device(PARAMS, batch, feat, D, patch_size, patch_shift, featName, time_major=None, dtype=None):
    """Row N1, device resident: featuregrams of a whole batch (``feat``: CUDA float32 in the batch layout, e.g. the
    output of engine.featuregram) -> the model-ready patch tensor on the device.  Per-file StandardScaler (unless
    frame_level_scaling; in place on ``feat``), tiling of clips shorter than the patch, patch gather, H/P selection
    by the feature name; the TCN models ('Lemaire_et_al' in the model name) get (n, W, nFeat) -- the transpose of
    Proposed_Work_Results.py:235-236 -- the CNNs (n, nFeat, W, 1).  float32 by default (what the network consumes)."""
    import torch
    from . import engine
    tcn = 'Lemaire_et_al' in PARAMS['Model']
    tm = tcn if time_major is None else bool(time_major)
    out = engine.patch_tensor(batch, feat, D, patch_size, patch_shift, standardize=not PARAMS['frame_level_scaling'],
                              rows=_patch_rows(featName), time_major=tm, dtype=dtype or torch.float32)
    return out if tcn else out.unsqueeze(3)


def get_feature_patches(PARAMS, FV, patch_size, patch_shift, featName):
    """lib/preprocessing.py:137-292 on the GPU: per-file StandardScaler of every feature row (unless
    frame_level_scaling), patch gather -> float64 (nPatch, nFeat, W[, 1]); harmonic/percussive halves
    selected or stacked on axis 1 by the feature name.  (Compatibility wrapper: numpy in, float64 numpy out; the
    device-resident form is feature_patches_device.)"""
    import torch
    from . import engine
    FV = np.asarray(FV)
    rows = _patch_rows(featName)
    ctx = engine.get_context()
    D, T = FV.shape
    # with frame_level_scaling the generators pass the float64 output of the Cython scale_data and nothing is
    # standardised: those values are gathered exactly; otherwise float32 featuregrams, as get_featuregram returns them
    keep64 = bool(PARAMS['frame_level_scaling']) and FV.dtype == np.float64
    feat = torch.from_numpy(np.ascontiguousarray(FV, dtype=np.float64 if keep64 else np.float32)).cuda()
    batch = engine.Batch(ctx, clip_frames=[T])
    patches = engine.patch_tensor(batch, feat.view(-1), D, patch_size, patch_shift,
                                  standardize=not PARAMS['frame_level_scaling'], rows=rows, time_major=False,
                                  dtype=torch.float64).cpu().numpy()
    batch.close()
    if 'Lemaire_et_al' not in PARAMS['Model']:
        patches = np.expand_dims(patches, axis=3)
    return patches


def get_feature_patches_dafx(PARAMS, FV, patch_size, patch_shift, featName):
    """The long-form script's copy (DAFx12_Speech_Music_Detection_B3_MTL_v2.py:257-292): no standardisation inside (the
    script standardises every file before cutting it into blocks, :613-626), float32 patches."""
    return get_feature_patches(dict(PARAMS, frame_level_scaling=True), np.asarray(FV, dtype=np.float32), patch_size,
                               patch_shift, featName).astype(np.float32)


def get_data_statistics(FV, stat_type='skew', axis=0):
    """lib/cython_impl/tools.pyx:169-211 on the GPU: per-patch mean / variance / skew / kurtosis vectors of a
    (N, f, t) patch array along axis 0 (over f -> (N, t)) or 1 (over t -> (N, f))."""
    import torch
    from . import engine
    FV = np.asarray(FV)
    if FV.ndim == 4 and FV.shape[3] == 1:
        FV = FV[:, :, :, 0]                               # np.squeeze of the CNN patches
    ctx = engine.get_context()
    x = torch.from_numpy(np.ascontiguousarray(FV, dtype=np.float64)).cuda()
    return engine.patch_statistics(ctx, x, stat_type, axis).cpu().numpy()


# ============================================================================ global statistics
def scale_data(FV, mean, stdev):
    """lib/preprocessing.py:590-614 ((x - mean) / stdev, numpy promotion of the inputs: float32 arithmetic -- two
    rounded operations, bit-identical to numpy -- when everything is float32, float64 otherwise)."""
    import torch
    from . import engine
    dt = np.result_type(np.asarray(FV).dtype, np.asarray(mean).dtype, np.asarray(stdev).dtype)
    if dt != np.float32:
        return _scale(FV, mean, stdev, 0.0, dt)
    FV = np.ascontiguousarray(FV, dtype=np.float32)
    D, T = FV.shape
    ctx = engine.get_context()
    batch = engine.Batch(ctx, clip_frames=[T])
    out = engine.scale_data_f32(batch, torch.from_numpy(FV.ravel()).cuda(),
                                D, torch.from_numpy(np.ascontiguousarray(mean, dtype=np.float32)).cuda(),
                                torch.from_numpy(np.ascontiguousarray(stdev, dtype=np.float32)).cuda())
    res = out.cpu().numpy().reshape(D, T)
    batch.close()
    return res


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


class _Moments:
    """Raw moments of get_data_stats accumulated on the device over many batches of featuregrams, with the
    reference's per-file dropping of feature rows that hold a NaN / Inf (lib/preprocessing.py:507-508)."""

    def __init__(self, ctx, n_classes, chunk_frames=1 << 20):
        self.ctx, self.n_classes, self.chunk_frames = ctx, n_classes, chunk_frames
        self.D = None
        self.acc = None
        self.bad_rows = None                  # feature rows dropped (must be the same rows in every file)
        self.n_files = 0
        self.pend, self.pend_cls, self.pend_frames = [], [], 0

    def add(self, fv: np.ndarray, cls: int):
        if self.D is None:
            self.D = fv.shape[0]
        self.pend.append(fv)
        self.pend_cls.append(cls)
        self.pend_frames += fv.shape[1]
        if self.pend_frames >= self.chunk_frames:
            self.flush()

    def add_device(self, batch, feat, D, classes):
        """Featuregrams already on the device (batch layout)."""
        from . import engine
        if self.D is None:
            self.D = D
        self.acc = engine.moments(batch, feat, D, classes, self.n_classes, acc=self.acc)
        self._rows(batch, feat, D)

    def _rows(self, batch, feat, D):
        from . import engine
        bad_total = float(self.acc[-1].item())            # synchronises; 8 bytes
        seen = getattr(self, "_bad_seen", 0.0)
        if bad_total != seen:                             # this batch holds non-finite values: which rows, which files?
            flags = engine.row_nonfinite(batch, feat, D).cpu().numpy().astype(bool)
            for c in range(batch.n_clips):
                self._merge(flags[c])
            self._bad_seen = bad_total
        elif self.bad_rows is not None and self.bad_rows.any():
            self._merge(np.zeros(D, dtype=bool))          # a clean file after a file with dropped rows
        self.n_files += batch.n_clips

    def _merge(self, mask):
        if self.bad_rows is None:
            if self.n_files > 0 and mask.any():
                raise ValueError("operands could not be broadcast together: files drop different non-finite "
                                 "feature rows (lib/preprocessing.py:507-529)")
            self.bad_rows = mask.copy()
        elif not np.array_equal(self.bad_rows, mask):
            raise ValueError("operands could not be broadcast together: files drop different non-finite feature rows "
                             "(lib/preprocessing.py:507-529)")

    def flush(self):
        import torch
        from . import engine
        if not self.pend:
            return
        batch = engine.Batch(self.ctx, clip_frames=[fv.shape[1] for fv in self.pend])
        n = sum(fv.size for fv in self.pend)
        host = _pinned(n, np.float32, "statfeat")
        o = 0
        for fv in self.pend:
            host[o:o + fv.size] = np.asarray(fv, dtype=np.float32).ravel()
            o += fv.size
        flat = torch.from_numpy(host).cuda(non_blocking=True)
        self.acc = engine.moments(batch, flat, self.D, self.pend_cls, self.n_classes, acc=self.acc)
        self._rows(batch, flat, self.D)                   # synchronises: the pinned buffer is free again
        batch.close()
        self.pend, self.pend_cls, self.pend_frames = [], [], 0

    def result(self, group=None):
        import torch
        from .dist import allreduce_moments, finalize_stats, moments_size
        self.flush()
        err = None
        if self.acc is None:
            if self.D is None:
                err = ValueError("no featuregrams given")
                self.D = 1
            self.acc = torch.zeros(moments_size(self.D, self.n_classes), dtype=torch.float64, device='cuda')
        allreduce_moments(self.acc, group)                # every rank reaches the collective before anyone raises
        if err is not None:
            raise err
        host = self.acc.cpu().numpy().copy()
        host[-1] = 0.0                                    # non-finite values sit in dropped rows only (checked above)
        mean, std, counts = finalize_stats(host, self.D, self.n_classes)
        if self.bad_rows is not None and self.bad_rows.any():
            mean, std = mean[~self.bad_rows], std[~self.bad_rows]
        return mean, std, [int(c) for c in counts]


def data_stats_from_featuregrams(class_to_fvs: Dict[str, List[np.ndarray]], classes: Sequence[str],
                                 group=None, chunk_frames: int = 1 << 20):
    """Global feature mean / stdev of get_data_stats (lib/preprocessing.py:461-586) from featuregrams held
    in memory: raw float64 moments per batch on the GPU, SUM all-reduce over the process group (the only
    collective on the path; pass featuregrams of THIS rank's files only), closed-form finish.
    Returns (mean f32[D], stdev f32[D], *per-class counts in the order of ``classes``)."""
    from . import engine
    m = _Moments(engine.get_context(), len(classes), chunk_frames)
    for k, name in enumerate(classes):
        for fv in class_to_fvs.get(name, []):
            m.add(fv, k)
    mean, std, counts = m.result(group)
    return (mean, std, *counts)


def _stat_jobs(PARAMS, files):
    """(class name, speech path, music path, SMR) of every file get_data_stats visits (lib/preprocessing.py:476-504)."""
    classes = PARAMS['classes']
    jobs = []
    for clNum in classes.keys():
        name = classes[clNum]
        file_list = files['speech+music'] if name == 'speech_music' else files[name]
        for fl in file_list:
            if name == 'speech_music':
                jobs.append((name, PARAMS['folder'] + '/speech/' + fl['speech'], PARAMS['folder'] + '/music/' + fl['music'],
                             fl['SMR']))
            elif name == 'music':
                jobs.append((name, '', PARAMS['folder'] + '/music/' + fl.split('.')[0] + '.wav', -1))
            else:
                jobs.append((name, PARAMS['folder'] + '/' + name + '/' + fl.split('.')[0] + '.wav', '', -1))
    return jobs


def _shard_jobs(jobs, rank, world):
    """Contiguous slice of the job list for one rank (every job lands on exactly one rank; the file lengths are not
    known before decoding, so the cut is by count)."""
    if world <= 1:
        return jobs
    per = (len(jobs) + world - 1) // world
    return jobs[rank * per:(rank + 1) * per]


def get_data_stats(PARAMS, files, loader=None, group=None, shard=True, batch_samples: int = 1 << 27):
    """Same signature and return value as lib/preprocessing.py:461-586: (mean f32[D], stdev f32[D], nMuFrames,
    nSpFrames, nSpMuFrames).  Every file of every class goes through the feature cache (``.npy`` under
    feature_opDir) or is computed: files without a cache entry are decoded on the host, uploaded as PCM in
    batches, prepared, featurised and reduced to raw moments on the device by one hpss_pipeline_run per batch (the
    features come back only to be written to the cache); speech+music mixes go file by file.  Under
    torch.distributed (``shard``) the job list is cut into one contiguous slice per rank and the moment vector is
    SUM-all-reduced once, at the end."""
    import torch
    from . import engine
    classes = PARAMS['classes']
    folder = PARAMS['feature_opDir']
    model = PARAMS['Model']
    featName, n_fft, n_mels = PARAMS['featName'][model], PARAMS['n_fft'][model], PARAMS['n_mels'][model]
    names = [classes[k] for k in classes.keys()]
    jobs = _stat_jobs(PARAMS, files)
    rank, world = 0, 1
    try:
        import torch.distributed as dist
        if shard and dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
    except Exception:
        pass
    jobs = _shard_jobs(jobs, rank, world)
    ctx = engine.get_context()
    mom = _Moments(ctx, len(names))
    loader = loader or load_pcm
    prm = make_params(PARAMS, FS, n_fft, n_mels, featName)
    pend = []                                             # (class index, cache path, pcm) of plain files to compute

    def run_pending():
        if not pend:
            return
        lens = [len(p[2]) for p in pend]
        dt = np.int16 if all(p[2].dtype == np.int16 for p in pend) else np.float32
        host = _pinned(sum(lens), dt, "pcm")
        o = 0
        for _, _, x in pend:
            host[o:o + len(x)] = x if x.dtype == dt else x.astype(np.float32) / np.float32(32768.0)
            o += len(x)
        pl = engine.Pipeline(ctx, lens, prm, pcm_dtype=dt, prepare=True, fs=FS)
        feat = _pinned(pl.rows * pl.total_frames, np.float32, "feat")
        acc = np.zeros(len(names) * pl.rows + pl.rows + len(names) + 1)
        pl.run(host, feat_host=feat, clip_class=[p[0] for p in pend], n_classes=len(names), moments=acc)
        if acc[-1] == 0:                                  # the common case: moments straight from the pipeline
            if mom.D is None:
                mom.D = pl.rows
            t = torch.from_numpy(acc).cuda()
            mom.acc = t if mom.acc is None else mom.acc + t
            if mom.bad_rows is not None and mom.bad_rows.any():
                mom._merge(np.zeros(pl.rows, dtype=bool))
            mom.n_files += len(pend)
        for c, (k, cache, _) in enumerate(pend):
            a, b = pl.rows * int(pl.frame_offsets[c]), pl.rows * int(pl.frame_offsets[c + 1])
            fv = feat[a:b].reshape(pl.rows, -1)
            os.makedirs(os.path.dirname(cache) + '/', exist_ok=True)
            np.save(cache, fv)
            if acc[-1] != 0:                              # non-finite values: the slow path sorts out which rows
                mom.add(fv.copy(), k)
        pl.close()
        pend.clear()

    pend_samples = 0
    for (name, sp, mu, smr) in jobs:
        k = names.index(name)
        cache = folder + '/' + name + '/' + _feature_name_of_file(sp, mu, smr) + '.npy'
        if os.path.exists(cache) or name == 'speech_music':
            mom.add(get_featuregram(PARAMS, name, folder, sp, mu, smr, n_fft, n_mels, featName, loader=loader), k)
        else:
            x = _as_pcm(loader(sp or mu))
            pend.append((k, cache, x))
            pend_samples += len(x)
            if pend_samples >= batch_samples:
                run_pending()
                pend_samples = 0
    run_pending()
    mean, std, counts = mom.result(group)
    by_name = dict(zip(names, counts))
    return mean, std, by_name.get('music', 0), by_name.get('speech', 0), by_name.get('speech_music', 0)


def save_data_stats(PARAMS, fold, mean, stdev, nFrames, split='train'):
    """The statistics pickle of Baseline_Results.py:609-616 (lib/misc.py:20-22): same file name, keys and protocol."""
    import pickle
    name = 'data_stats_fold' + str(fold) + '_' + str(len(PARAMS['classes'])) + 'class_' + split
    stats = {'mean': mean, 'stdev': stdev, 'nFrames': list(nFrames)}
    with open(PARAMS['feature_opDir'] + '/' + name + '.pkl', 'wb') as f:
        pickle.dump(stats, f, pickle.HIGHEST_PROTOCOL)
    return stats


def load_or_compute_data_stats(PARAMS, files, fold, loader=None, split='train'):
    """Baseline_Results.py:607-621: reuse feature_opDir/data_stats_fold<k>_<n>class_train.pkl when present, else
    run get_data_stats and write it.  Returns the dict the reference pickles."""
    import pickle
    name = 'data_stats_fold' + str(fold) + '_' + str(len(PARAMS['classes'])) + 'class_' + split
    path = PARAMS['feature_opDir'] + '/' + name + '.pkl'
    if os.path.exists(path):
        with open(path, 'rb') as f:
            return pickle.load(f)
    mean, stdev, nMu, nSp, nSpMu = get_data_stats(PARAMS, files, loader=loader)
    return save_data_stats(PARAMS, fold, mean, stdev, [nMu, nSp, nSpMu], split)
