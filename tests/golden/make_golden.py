#!/usr/bin/env python3
"""Generate tests/golden/reference_glue.npz by running the REFERENCE's own code.

Runs only in the authoring container (needs /root/reference).  It imports the reference's
``lib/preprocessing.py`` unmodified and calls its public functions (get_featuregram,
get_feature_patches, get_data_stats, scale_data, load_and_preprocess_signal, mix_signals).
Two things the reference needs are not installable here and are shimmed:

* ``librosa`` -- replaced by a module whose leaves are the restatements in
  oracle/librosa_restated.py (stft, hpss, melspectrogram, power_to_db, rms, load).  So these
  vectors pin the reference's GLUE (feature-name dispatch, H/P stacking, dtype casts, silence
  removal + mixing, per-file StandardScaler, patch gather, two-pass statistics), not the librosa
  leaves themselves (see oracle/__init__.py for what pins those).
* ``lib.cython_impl.tools`` -- the reference's Cython leaf compiled from its own source into
  oracle/_ref (oracle/build_ref.py).  Its module-level ``medfilt`` is wrapped to return float64,
  which is what the scipy of the reference's era did (scipy >= 1.11 keeps the int dtype and the
  typed buffer assignment at tools.pyx:98 then fails).

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import build_ref                      # noqa: E402
from oracle import librosa_restated as lr         # noqa: E402

_AUDIO = {}


def _rms(y=None, frame_length=2048, hop_length=512, center=True, pad_mode="reflect"):
    if center:
        y = np.pad(y, int(frame_length // 2), mode=pad_mode)
    n = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n)[None, :]
    x = y[idx]
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0, keepdims=True))


def install_shims():
    librosa = types.ModuleType("librosa")
    core = types.ModuleType("librosa.core")
    core.stft = lambda y, n_fft=2048, hop_length=None, win_length=None, center=True: lr.stft(
        y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, center=center)
    core.power_to_db = lr.power_to_db
    core.load = lambda fName, mono=True, sr=16000: (_AUDIO[fName].copy(), sr)
    decompose = types.ModuleType("librosa.decompose")
    decompose.hpss = lambda S, kernel_size=31, power=2.0, mask=False, margin=1.0: lr.hpss(
        S, kernel_size=kernel_size, power=power, mask=mask, margin=margin)
    feature = types.ModuleType("librosa.feature")
    feature.melspectrogram = lr.melspectrogram
    feature.rms = _rms
    filters = types.ModuleType("librosa.filters")
    filters.mel = lr.mel
    librosa.core, librosa.decompose, librosa.feature, librosa.filters = core, decompose, feature, filters
    for name, mod in [("librosa", librosa), ("librosa.core", core), ("librosa.decompose", decompose),
                      ("librosa.feature", feature), ("librosa.filters", filters)]:
        sys.modules[name] = mod
    tools = build_ref.load() if build_ref.build() else None
    if tools is None:
        raise SystemExit("oracle/_ref could not be built (is /root/reference present?)")
    from scipy.signal import medfilt as _medfilt
    tools.medfilt = lambda x, k: _medfilt(np.asarray(x, dtype=np.float64), k)
    sys.modules["lib.cython_impl.tools"] = tools
    sys.path.insert(0, REF)
    import lib.preprocessing as preproc
    return preproc


def make_wave(seed, seconds, silent=()):
    """Synthetic 16 kHz clip with optional silent stretches (to exercise removeSilence)."""
    from sm_hpss_mtl_b200 import synth
    n = int(seconds * 16000)
    x = synth.synth_clip(seed, n)
    for (a, b) in silent:
        x[int(a * 16000):int(b * 16000)] *= 1e-4
    return x.astype(np.float32)


def main():
    preproc = install_shims()
    out = {}
    _AUDIO["/d/speech/sp0.wav"] = make_wave(1, 1.4, silent=[(0.3, 0.5), (0.9, 1.05)])
    _AUDIO["/d/music/mu0.wav"] = make_wave(2, 1.1)
    _AUDIO["/d/music/mu1.wav"] = make_wave(3, 0.9, silent=[(0.2, 0.45), (0.6, 0.8)])
    _AUDIO["/d/speech/sp1.wav"] = make_wave(4, 1.2)
    for k, v in _AUDIO.items():
        out["audio:" + k] = v

    model = "Lemaire_et_al_MTL"
    PARAMS = {"Tw": 25, "Ts": 10, "Model": model, "l_harm": {model: 21}, "l_perc": {model: 11},
              "frame_level_scaling": False}
    # signal preparation (the step before the hot path)
    _AUDIO["/d/speech/short.wav"] = make_wave(5, 0.06)            # below 0.1 s: exercises the doubling (:345-347)
    _AUDIO["/d/music/onesil.wav"] = make_wave(6, 1.0, silent=[(0.4, 0.7)])   # ONE stretch: marked, not removed
    out["audio:/d/speech/short.wav"] = _AUDIO["/d/speech/short.wav"]
    out["audio:/d/music/onesil.wav"] = _AUDIO["/d/music/onesil.wav"]
    tools = sys.modules["lib.cython_impl.tools"]
    for k in _AUDIO:
        x, fs = preproc.load_and_preprocess_signal(k, 25, 10)
        out["prep:" + k] = np.asarray(x, dtype=np.float32)
        # the gate's intermediates straight from the reference's Cython leaf (tools.pyx:42-134)
        y = preproc.normalize_signal(_AUDIO[k].copy())
        energy = _rms(y=y, frame_length=400, hop_length=160)[0]
        _, smark, fmark, _ = tools.removeSilence(y, len(y), energy, len(energy), 16000, 25, 10)
        out["gate:frame:" + k] = np.asarray(fmark, dtype=np.int32)
        out["gate:sample:" + k] = np.asarray(smark, dtype=np.uint8)
        out["gate:energy:" + k] = np.asarray(energy, dtype=np.float32)
    mix = preproc.mix_signals(out["prep:/d/speech/sp0.wav"], out["prep:/d/music/mu0.wav"], 5)
    out["mix:sp0+mu0@5"] = np.asarray(mix, dtype=np.float32)

    feats = ["Spec", "LogSpec", "MelSpec", "LogMelSpec", "PercSpec", "HarmPercSpec",
             "LogHarmPercSpec", "MelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec", "LogMelHarmPercSpec"]
    for fn in feats:
        n_fft, n_mels = (512, 21) if fn in ("LogHarmPercSpec",) else (400, 40)
        fv = preproc.get_featuregram(PARAMS, "speech", "/nonexistent", "/d/speech/sp0.wav", "", -1, n_fft, n_mels,
                                     fn, save_feat=False)
        out[f"fv:speech:{fn}"] = fv
        fv = preproc.get_featuregram(PARAMS, "speech_music", "/nonexistent", "/d/speech/sp0.wav", "/d/music/mu0.wav",
                                     5, n_fft, n_mels, fn, save_feat=False)
        out[f"fv:speech_music:{fn}"] = fv
    fvm = preproc.get_featuregram(PARAMS, "music", "/nonexistent", "", "/d/music/mu1.wav", -1, 400, 40,
                                  "LogMelHarmPercSpec", save_feat=False)
    out["fv:music:LogMelHarmPercSpec"] = fvm
    # the headline configuration of BASELINE.json configs[1]: 120 mel bands, median kernels 31 / 31
    P31 = dict(PARAMS, l_harm={model: 31}, l_perc={model: 31})
    out["fv120:speech:LogMelHarmPercSpec"] = preproc.get_featuregram(
        P31, "speech", "/nonexistent", "/d/speech/sp1.wav", "", -1, 400, 120, "LogMelHarmPercSpec", save_feat=False)
    out["fv120:speech_music:LogMelHarmPercSpec"] = preproc.get_featuregram(
        P31, "speech_music", "/nonexistent", "/d/speech/sp1.wav", "/d/music/mu1.wav", -5, 400, 120,
        "LogMelHarmPercSpec", save_feat=False)
    # per-patch statistics of the skewness ablation (tools.pyx:169-211), from the reference's Cython leaf

    # NOTE: the reference standardises IN PLACE (StandardScaler(copy=False) on views of FV), i.e.
    # get_feature_patches mutates its argument; copies keep the golden featuregrams pristine.
    # patches: Lemaire (no trailing axis) and a CNN model (expand_dims), incl. the T < patch_size tiling
    FV = out["fv:speech:LogMelHarmPercSpec"]
    for mdl in ("Lemaire_et_al_MTL", "Doukhan_et_al_MTL"):
        P = dict(PARAMS, Model=mdl)
        for fn in ("LogMelHarmPercSpec", "LogMelHarmSpec", "LogMelPercSpec"):
            for (W, sh) in [(68, 68), (49, 24), (249, 24)]:
                out[f"patch:{mdl}:{fn}:{W}:{sh}"] = preproc.get_feature_patches(P, FV.copy(), W, sh, fn)
    out["patch:Doukhan_et_al_MTL:Spec:21:21"] = preproc.get_feature_patches(
        dict(PARAMS, Model="Doukhan_et_al_MTL"), out["fv:speech:Spec"].copy(), 21, 21, "Spec")
    P = dict(PARAMS, Model="Lemaire_et_al_MTL", frame_level_scaling=True)
    out["patch:fls:LogMelHarmPercSpec:49:24"] = preproc.get_feature_patches(P, FV.copy(), 49, 24, "LogMelHarmPercSpec")

    # global statistics over cached featuregrams (the reference np.load()s them from feature_opDir)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        P = dict(PARAMS, classes={0: "music", 1: "speech", 2: "speech_music"}, feature_opDir=td, folder="/d",
                 featName={model: "LogMelHarmPercSpec"}, n_fft={model: 400}, n_mels={model: 40})
        files = {"music": ["mu0.wav", "mu1.wav"], "speech": ["sp0.wav", "sp1.wav"],
                 "speech+music": [{"speech": "sp0.wav", "music": "mu0.wav", "SMR": 5},
                                  {"speech": "sp1.wav", "music": "mu1.wav", "SMR": -5}]}
        devnull = open(os.devnull, "w")
        old = sys.stdout
        sys.stdout = devnull
        try:
            mean, std, nMu, nSp, nSpMu = preproc.get_data_stats(P, files)
        finally:
            sys.stdout = old
        out["stats:mean"], out["stats:std"] = mean, std
        out["stats:counts"] = np.array([nMu, nSp, nSpMu], dtype=np.int64)
        for cls in ("music", "speech", "speech_music"):
            for f in sorted(os.listdir(os.path.join(td, cls))):
                out[f"statsfv:{cls}:{f[:-4]}"] = np.load(os.path.join(td, cls, f))
    out["scaled:py"] = preproc.scale_data(FV, mean, std)
    out["scaled:cy"] = tools.scale_data(FV, mean, std)
    pat = out["patch:Lemaire_et_al_MTL:LogMelHarmPercSpec:49:24"]
    for st in ("mean", "variance", "skew", "kurtosis"):
        for ax in (0, 1):
            out[f"pstat:{st}:{ax}"] = tools.get_data_statistics(pat, stat_type=st, axis=ax)

    path = os.path.join(HERE, "reference_glue.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
