"""Tensor-level host API over the C ABI: contexts, batch layouts and the stage / fused calls.

torch is used here only as the owner of device memory and streams; all arithmetic happens in
libhpss_b200.so.  Arrays follow the batch layout of include/hpss_b200.h: a flat tensor holding,
clip after clip, a C-ordered (rows, T_c) matrix -- the layout of the reference's numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import FEATURES, Params, ParameterError, check

_contexts = {}
_ctx_lock = threading.Lock()


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_ptr(t: Optional[torch.Tensor], dtype=None, name="tensor") -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (sm_hpss_mtl_b200 has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


class Context:
    """Per-device library context (plan caches + workspace)."""

    def __init__(self, device: Optional[int] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("sm_hpss_mtl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(self.lib.hpss_ctx_create(self.device, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hpss_ctx_destroy(self.handle)
            self.handle = None

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.hpss_ctx_workspace_bytes(self.handle))


def get_context(device: Optional[int] = None) -> Context:
    """Process-wide context of a device (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("sm_hpss_mtl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    with _ctx_lock:
        ctx = _contexts.get(dev)
        if ctx is None:
            ctx = _contexts[dev] = Context(dev)
        return ctx


def launch_count() -> int:
    return int(_lib.load().hpss_launch_count())


class Batch:
    """Layout of one batch of clips (see include/hpss_b200.h)."""

    def __init__(self, ctx: Context, clip_lengths: Optional[Sequence[int]] = None,
                 clip_frames: Optional[Sequence[int]] = None, n_fft: int = 0, hop_length: int = 0):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        if clip_lengths is not None:
            arr = np.ascontiguousarray(clip_lengths, dtype=np.int64)
            check(self.lib.hpss_batch_from_samples(ctx.handle, arr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   arr.size, int(n_fft), int(hop_length), C.byref(h)))
        elif clip_frames is not None:
            arr = np.ascontiguousarray(clip_frames, dtype=np.int64)
            check(self.lib.hpss_batch_from_frames(ctx.handle, arr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                  arr.size, C.byref(h)))
        else:
            raise ValueError("give clip_lengths (samples) or clip_frames")
        self.handle = h
        self.n_clips = int(self.lib.hpss_batch_n_clips(h))
        self.total_frames = int(self.lib.hpss_batch_total_frames(h))
        self.total_samples = int(self.lib.hpss_batch_total_samples(h))
        self.n_fft, self.hop_length = int(n_fft), int(hop_length)
        fo = np.zeros(self.n_clips + 1, dtype=np.int64)
        so = np.zeros(self.n_clips + 1, dtype=np.int64)
        check(self.lib.hpss_batch_frame_offsets(h, fo.ctypes.data_as(C.POINTER(C.c_int64))))
        check(self.lib.hpss_batch_sample_offsets(h, so.ctypes.data_as(C.POINTER(C.c_int64))))
        self.frame_offsets, self.sample_offsets = fo, so

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hpss_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frames(self, c: int) -> int:
        return int(self.frame_offsets[c + 1] - self.frame_offsets[c])

    def split(self, flat: torch.Tensor, rows: int) -> List[torch.Tensor]:
        """Views (rows, T_c) of every clip of a flat batch array."""
        out = []
        for c in range(self.n_clips):
            a, b = rows * int(self.frame_offsets[c]), rows * int(self.frame_offsets[c + 1])
            out.append(flat[a:b].view(rows, -1) if b > a else flat[a:b].view(rows, 0))
        return out

    def clip(self, flat: torch.Tensor, rows: int, c: int) -> torch.Tensor:
        a, b = rows * int(self.frame_offsets[c]), rows * int(self.frame_offsets[c + 1])
        return flat[a:b].view(rows, -1)


def make_params(n_fft=400, win_length=400, hop_length=160, l_harm=21, l_perc=11, n_mels=120, mel_sr=22050,
                feature="LOGMEL_HARMPERC", amin=1e-10, top_db=80.0) -> Params:
    fid = FEATURES[feature] if isinstance(feature, str) else int(feature)
    return Params(int(n_fft), int(win_length), int(hop_length), int(l_harm), int(l_perc), int(n_mels),
                  int(mel_sr), fid, float(amin), float(-1.0 if top_db is None else top_db))


def feature_rows(params: Params) -> int:
    return int(_lib.load().hpss_feature_rows(C.byref(params)))


# ---------------------------------------------------------------------------- tables
def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    out = np.empty((n_mels, 1 + n_fft // 2), dtype=np.float32)
    check(_lib.load().hpss_mel_filterbank(int(sr), int(n_fft), int(n_mels), C.c_void_p(out.ctypes.data)))
    return out


def stft_window(n_fft: int, win_length: int) -> np.ndarray:
    out = np.empty(n_fft, dtype=np.float32)
    check(_lib.load().hpss_stft_window(int(n_fft), int(win_length), C.c_void_p(out.ctypes.data)))
    return out


# ---------------------------------------------------------------------------- stages
def stft_mag(batch: Batch, wave: torch.Tensor, n_fft: int, win_length: int, hop_length: int, power: bool = False,
             return_complex: bool = False):
    F = n_fft // 2 + 1
    S = torch.empty(F * batch.total_frames, dtype=torch.float32, device=wave.device)
    cplx = torch.empty(F * batch.total_frames, dtype=torch.complex64, device=wave.device) if return_complex else None
    check(batch.lib.hpss_stft_mag(batch.ctx.handle, batch.handle, _dev_ptr(wave, torch.float32, "wave"), int(n_fft),
                                  int(win_length), int(hop_length), int(bool(power)), _dev_ptr(S),
                                  _dev_ptr(cplx), _stream_ptr()))
    return (S, cplx) if return_complex else S


def median_time(batch: Batch, S: torch.Tensor, rows: int, k: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty_like(S) if out is None else out
    if out.numel() != S.numel():
        raise ValueError("median_time: out must have the size of S")
    check(batch.lib.hpss_median_time(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"), int(rows), int(k),
                                     _dev_ptr(out, torch.float32, "out"), _stream_ptr()))
    return out


def median_freq(batch: Batch, S: torch.Tensor, rows: int, k: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty_like(S) if out is None else out
    if out.numel() != S.numel():
        raise ValueError("median_freq: out must have the size of S")
    check(batch.lib.hpss_median_freq(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"), int(rows), int(k),
                                     _dev_ptr(out, torch.float32, "out"), _stream_ptr()))
    return out


def mask_mel_log(batch: Batch, S: torch.Tensor, harm: Optional[torch.Tensor], perc: Optional[torch.Tensor], rows: int,
                 mel: Optional[torch.Tensor] = None, pre_square: bool = False, log_power: bool = False,
                 amin: float = 1e-10, mel_sr: Optional[int] = None, n_mels: int = 0):
    """Returns (out, clip_max): out rows = streams * (n_mels or rows); clip_max (uint32-coded) or None.
    The basis is either an explicit dense ``mel`` (n_mels, rows) tensor or, with ``mel_sr`` and ``n_mels``,
    the library's cached Slaney filterbank for that sample rate (what melspectrogram itself builds)."""
    ns = 2 if harm is not None else 1
    clip_max = torch.empty(ns * max(1, batch.n_clips), dtype=torch.int32, device=S.device) if log_power else None
    if mel is None and mel_sr is not None and n_mels > 0:
        out = torch.empty(ns * n_mels * batch.total_frames, dtype=torch.float32, device=S.device)
        check(batch.lib.hpss_mask_mel_log_sr(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"),
                                             _dev_ptr(harm, torch.float32, "harm"), _dev_ptr(perc, torch.float32, "perc"),
                                             int(rows), int(mel_sr), int(n_mels), int(bool(pre_square)), int(log_power),
                                             float(amin), _dev_ptr(out), _dev_ptr(clip_max), _stream_ptr()))
        return out, clip_max
    n_mels = 0 if mel is None else int(mel.shape[0])
    if mel is not None and tuple(mel.shape) != (n_mels, rows):
        raise ValueError(f"mel must be (n_mels, {rows})")
    rows_out = ns * (n_mels if mel is not None else rows)
    out = torch.empty(rows_out * batch.total_frames, dtype=torch.float32, device=S.device)
    check(batch.lib.hpss_mask_mel_log(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"),
                                      _dev_ptr(harm, torch.float32, "harm"), _dev_ptr(perc, torch.float32, "perc"),
                                      int(rows), _dev_ptr(mel, torch.float32, "mel"), n_mels, int(bool(pre_square)),
                                      int(log_power), float(amin), _dev_ptr(out), _dev_ptr(clip_max),
                                      _stream_ptr()))
    return out, clip_max


def topdb_clip(batch: Batch, out: torch.Tensor, rows_per_stream: int, n_streams: int, clip_max: torch.Tensor,
               top_db: float = 80.0) -> torch.Tensor:
    if top_db < 0:
        raise ParameterError("top_db must be non-negative")
    check(batch.lib.hpss_topdb_clip(batch.ctx.handle, batch.handle, _dev_ptr(out, torch.float32, "out"),
                                    int(rows_per_stream), int(n_streams), _dev_ptr(clip_max), float(top_db),
                                    _stream_ptr()))
    return out


def dct_basis(n_mels: int, n_mfcc: int) -> np.ndarray:
    """(n_mfcc, n_mels) float32 orthonormal DCT-II basis (scipy.fftpack.dct(type=2, norm='ortho') along the mel axis)."""
    out = np.empty((n_mfcc, n_mels), dtype=np.float32)
    check(_lib.load().hpss_dct_basis(int(n_mels), int(n_mfcc), C.c_void_p(out.ctypes.data)))
    return out


def dct_mfcc(batch: Batch, feat: torch.Tensor, rows_per_stream: int, n_streams: int, n_mfcc: int = 20,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Extension (not in the reference): per stream, the first n_mfcc rows of the orthonormal DCT-II along the mel
    axis of a (n_streams * rows_per_stream, T_c) log-mel featuregram -> (n_streams * n_mfcc, T_c) per clip."""
    if out is None:
        out = torch.empty(n_streams * n_mfcc * batch.total_frames, dtype=torch.float32, device=feat.device)
    check(batch.lib.hpss_dct_mfcc(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"),
                                  int(rows_per_stream), int(n_streams), int(n_mfcc), _dev_ptr(out, torch.float32, "out"),
                                  _stream_ptr()))
    return out


# ---------------------------------------------------------------------------- fused
def featuregram(batch: Batch, wave: torch.Tensor, params: Params, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = feature_rows(params)
    if out is None:
        out = torch.empty(rows * batch.total_frames, dtype=torch.float32, device=wave.device)
    check(batch.lib.hpss_featuregram(batch.ctx.handle, batch.handle, _dev_ptr(wave, torch.float32, "wave"),
                                     C.byref(params), _dev_ptr(out, torch.float32, "out"), _stream_ptr()))
    return out


def featuregram_from_spec(batch: Batch, S: torch.Tensor, rows: int, params: Params) -> torch.Tensor:
    p = Params.from_buffer_copy(params)
    p.n_fft = 2 * (rows - 1)
    out = torch.empty(feature_rows(p) * batch.total_frames, dtype=torch.float32, device=S.device)
    check(batch.lib.hpss_featuregram_from_spec(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"),
                                               int(rows), C.byref(p), _dev_ptr(out), _stream_ptr()))
    return out


def host_alloc(n: int, dtype=np.float32) -> np.ndarray:
    """numpy array (float32 by default) over pinned host memory owned by the library (freed with the array)."""
    lib = _lib.load()
    dt = np.dtype(dtype)
    p = C.c_void_p()
    check(lib.hpss_host_alloc(C.byref(p), int(n) * dt.itemsize))
    buf = (C.c_byte * (int(n) * dt.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt)
    weakref.finalize(arr, lib.hpss_host_free, C.c_void_p(p.value))    # views keep `arr` alive via .base
    return arr


def ctx_check(ctx: Context) -> None:
    """Synchronise the current stream and raise what the kernels flagged since the last check
    (ParameterError for non-finite audio / negative spectrogram input, as librosa does)."""
    check(ctx.lib.hpss_ctx_check(ctx.handle, _stream_ptr()))


def validate_audio(ctx: Context, wave: torch.Tensor) -> None:
    """librosa.util.valid_audio as a device pass; the verdict arrives with the next ctx_check."""
    check(ctx.lib.hpss_validate_audio(ctx.handle, _dev_ptr(wave, torch.float32, "wave"), int(wave.numel()), _stream_ptr()))


def validate_nonneg(ctx: Context, x: torch.Tensor) -> None:
    """softmask's "X and X_ref must be non-negative" as a device pass (verdict with the next ctx_check)."""
    check(ctx.lib.hpss_validate_nonneg(ctx.handle, _dev_ptr(x, torch.float32, "x"), int(x.numel()), _stream_ptr()))


# ---------------------------------------------------------------------------- N2: signal preparation
def prep_out_length(n_samples: int, fs: int = 16000) -> int:
    return int(_lib.load().hpss_prep_out_length(int(n_samples), int(fs)))


def prep_num_frames(n_samples: int, win_length: int, hop_length: int) -> int:
    return int(_lib.load().hpss_prep_num_frames(int(n_samples), int(win_length), int(hop_length)))


def prep_signals(ctx: Context, pcm: torch.Tensor, clip_lengths: Sequence[int], fs: int = 16000, win_length: int = 400,
                 hop_length: int = 160, alpha: float = 0.025, beta: float = 0.075, markers: bool = False):
    """load_and_preprocess_signal after the decode (normalise, RMS gate, silence excision, doubling of clips
    shorter than 0.1 s, normalise) for a batch of files on the device.

    pcm: CUDA float32 or int16 tensor holding the files back to back.  Returns the prepared float32 signals back to
    back (clip c has prep_out_length(len_c) samples); with ``markers`` also (frame_marker int32, sample_marker uint8,
    n_sil int32) -- the reference's frame_silMarker / sample_silMarker and the number of qualifying stretches."""
    lens = np.ascontiguousarray(clip_lengths, dtype=np.int64)
    if pcm.dtype == torch.float32:
        fmt = _lib.PCM_F32
    elif pcm.dtype == torch.int16:
        fmt = _lib.PCM_S16
    else:
        raise ValueError(f"pcm must be float32 or int16, got {pcm.dtype}")
    if int(lens.sum()) != pcm.numel():
        raise ValueError(f"pcm has {pcm.numel()} samples, clip_lengths sum to {int(lens.sum())}")
    ol = lens.copy()                                     # hpss_prep_out_length, vectorised (lib/preprocessing.py:345-347)
    while True:
        short = ol / float(fs) < 0.1
        if not short.any():
            break
        ol[short] *= 2
    out_len = [int(v) for v in ol]
    out = torch.empty(int(ol.sum()), dtype=torch.float32, device=pcm.device)
    fm = sm = ns = None
    if markers:
        nfr = int((1 + (lens + 2 * (win_length // 2) - win_length) // hop_length).sum())      # hpss_prep_num_frames
        fm = torch.zeros(nfr, dtype=torch.int32, device=pcm.device)
        sm = torch.zeros(pcm.numel(), dtype=torch.uint8, device=pcm.device)
        ns = torch.zeros(max(1, lens.size), dtype=torch.int32, device=pcm.device)
    check(ctx.lib.hpss_prep_signals(ctx.handle, _dev_ptr(pcm, None, "pcm"), fmt, lens.ctypes.data_as(C.POINTER(C.c_int64)),
                                    int(lens.size), int(fs), int(win_length), int(hop_length), float(alpha), float(beta),
                                    _dev_ptr(out), _dev_ptr(fm), _dev_ptr(sm), _dev_ptr(ns), _stream_ptr()))
    if markers:
        return out, out_len, fm, sm, ns
    return out, out_len


def mix_signals(ctx: Context, sp: torch.Tensor, sp_lengths: Sequence[int], mu: torch.Tensor, mu_lengths: Sequence[int],
                target_db: Sequence[float]) -> torch.Tensor:
    """mix_signals (lib/preprocessing.py:297-325) for pairs of prepared signals on the device."""
    sl = np.ascontiguousarray(sp_lengths, dtype=np.int64)
    ml = np.ascontiguousarray(mu_lengths, dtype=np.int64)
    db = np.ascontiguousarray(target_db, dtype=np.float64)
    if not (sl.size == ml.size == db.size):
        raise ValueError("one speech length, music length and target ratio per pair")
    if int(sl.sum()) != sp.numel() or int(ml.sum()) != mu.numel():
        raise ValueError("signal buffers do not match the given lengths")
    out = torch.empty(sp.numel(), dtype=torch.float32, device=sp.device)
    check(ctx.lib.hpss_mix_signals(ctx.handle, _dev_ptr(sp, torch.float32, "sp"), sl.ctypes.data_as(C.POINTER(C.c_int64)),
                                   _dev_ptr(mu, torch.float32, "mu"), ml.ctypes.data_as(C.POINTER(C.c_int64)),
                                   C.c_void_p(db.ctypes.data), int(sl.size), _dev_ptr(out), _stream_ptr()))
    return out


class Pipeline:
    """Host-buffer pipeline (hpss_pipeline_*): decoded files or prepared waveforms in host memory -> features in
    host memory and / or the raw moments of get_data_stats, H2D / kernels / D2H overlapped over clip chunks."""

    def __init__(self, ctx: Context, clip_lengths: Sequence[int], params: Params, pcm_dtype=np.float32,
                 prepare: bool = False, fs: int = 16000, alpha: float = 0.025, beta: float = 0.075, n_chunks: int = 0):
        self.ctx, self.lib, self.params = ctx, ctx.lib, params
        self.pcm_dtype = np.dtype(pcm_dtype)
        fmt = {np.dtype(np.float32): _lib.PCM_F32, np.dtype(np.int16): _lib.PCM_S16}[self.pcm_dtype]
        lens = np.ascontiguousarray(clip_lengths, dtype=np.int64)
        h = C.c_void_p()
        check(self.lib.hpss_pipeline_create(ctx.handle, lens.ctypes.data_as(C.POINTER(C.c_int64)), int(lens.size),
                                            C.byref(params), fmt, int(bool(prepare)), int(fs), float(alpha), float(beta),
                                            int(n_chunks), C.byref(h)))
        self.handle = h
        self.n_clips = int(lens.size)
        self.total_samples = int(lens.sum())
        self.total_frames = int(self.lib.hpss_pipeline_total_frames(h))
        self.n_chunks = int(self.lib.hpss_pipeline_n_chunks(h))
        self.rows = feature_rows(params)
        fo = np.zeros(self.n_clips + 1, dtype=np.int64)
        check(self.lib.hpss_pipeline_frame_offsets(h, fo.ctypes.data_as(C.POINTER(C.c_int64))))
        self.frame_offsets = fo

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hpss_pipeline_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, pcm_host: np.ndarray, feat_host: Optional[np.ndarray] = None, clip_class: Optional[Sequence[int]] = None,
            n_classes: int = 0, moments: Optional[np.ndarray] = None, want_features: bool = True):
        """Returns (feat_host or None, moments or None).  ``moments`` (float64, moments layout of :func:`moments`)
        is accumulated into when given, created zeroed when ``clip_class`` is given without it."""
        if pcm_host.dtype != self.pcm_dtype or not pcm_host.flags.c_contiguous or pcm_host.size != self.total_samples:
            raise ValueError(f"pcm_host must be C-contiguous {self.pcm_dtype} with {self.total_samples} samples")
        if want_features and feat_host is None:
            feat_host = np.empty(self.rows * self.total_frames, dtype=np.float32)
        if feat_host is not None and (feat_host.dtype != np.float32 or feat_host.size != self.rows * self.total_frames):
            raise ValueError("feat_host has the wrong dtype/size")
        cls = None
        if clip_class is not None:
            cls = np.ascontiguousarray(clip_class, dtype=np.int32)
            if cls.size != self.n_clips:
                raise ValueError("clip_class must have one entry per clip")
            n = n_classes * self.rows + self.rows + n_classes + 1
            if moments is None:
                moments = np.zeros(n, dtype=np.float64)
            if moments.dtype != np.float64 or moments.size != n or not moments.flags.c_contiguous:
                raise ValueError(f"moments must be a contiguous float64 array of {n} entries")
        check(self.lib.hpss_pipeline_run(self.handle, C.c_void_p(pcm_host.ctypes.data),
                                         C.c_void_p(feat_host.ctypes.data) if feat_host is not None else C.c_void_p(0),
                                         C.c_void_p(cls.ctypes.data) if cls is not None else C.c_void_p(0),
                                         int(n_classes), C.c_void_p(moments.ctypes.data) if cls is not None else C.c_void_p(0)))
        return feat_host, moments


def featuregram_host(batch: Batch, wave_host: np.ndarray, params: Params, out_host: Optional[np.ndarray] = None):
    """Host buffers in, host buffers out (H2D + kernels + D2H pipelined inside the library)."""
    if wave_host.dtype != np.float32 or not wave_host.flags.c_contiguous:
        raise ValueError("wave_host must be C-contiguous float32")
    if wave_host.size != batch.total_samples:
        raise ValueError(f"wave_host has {wave_host.size} samples, batch expects {batch.total_samples}")
    rows = feature_rows(params)
    if out_host is None:
        out_host = np.empty(rows * batch.total_frames, dtype=np.float32)
    if out_host.dtype != np.float32 or out_host.size != rows * batch.total_frames:
        raise ValueError("out_host has the wrong dtype/size")
    check(batch.lib.hpss_featuregram_host(batch.ctx.handle, batch.handle, C.c_void_p(wave_host.ctypes.data),
                                          C.byref(params), C.c_void_p(out_host.ctypes.data)))
    return out_host


# ---------------------------------------------------------------------------- statistics / patches
def moments(batch: Batch, feat: torch.Tensor, D: int, clip_class: Sequence[int], n_classes: int, acc=None):
    """Accumulate raw moments into ``acc`` = float64 tensor [n_classes*D + D + n_classes + 1]
    (sums | sumsq | counts | nonfinite), created zeroed when None."""
    n = n_classes * D + D + n_classes + 1
    if acc is None:
        acc = torch.zeros(n, dtype=torch.float64, device=feat.device)
    cls = np.ascontiguousarray(clip_class, dtype=np.int32)
    if cls.size != batch.n_clips:
        raise ValueError("clip_class must have one entry per clip")
    base = acc.data_ptr()
    check(batch.lib.hpss_moments(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                 C.c_void_p(cls.ctypes.data), int(cls.size), int(n_classes), C.c_void_p(base),
                                 C.c_void_p(base + 8 * n_classes * D), C.c_void_p(base + 8 * (n_classes * D + D)),
                                 C.c_void_p(base + 8 * (n_classes * D + D + n_classes)), _stream_ptr()))
    return acc


def topdb_moments(batch: Batch, out: torch.Tensor, rows_per_stream: int, n_streams: int, clip_max: torch.Tensor,
                  top_db: float, clip_class: Sequence[int], n_classes: int, acc: Optional[torch.Tensor] = None):
    """top_db clip (in place) + moments of the clipped features in one pass (K3b + K5 fused)."""
    D = rows_per_stream * n_streams
    if acc is None:
        acc = torch.zeros(n_classes * D + D + n_classes + 1, dtype=torch.float64, device=out.device)
    cls = np.ascontiguousarray(clip_class, dtype=np.int32)
    if cls.size != batch.n_clips:
        raise ValueError("clip_class must have one entry per clip")
    base = acc.data_ptr()
    check(batch.lib.hpss_topdb_moments(batch.ctx.handle, batch.handle, _dev_ptr(out, torch.float32, "out"),
                                       int(rows_per_stream), int(n_streams), _dev_ptr(clip_max), float(top_db),
                                       C.c_void_p(cls.ctypes.data), int(cls.size), int(n_classes), C.c_void_p(base),
                                       C.c_void_p(base + 8 * n_classes * D), C.c_void_p(base + 8 * (n_classes * D + D)),
                                       C.c_void_p(base + 8 * (n_classes * D + D + n_classes)), _stream_ptr()))
    return acc


def featuregram_moments(batch: Batch, wave: torch.Tensor, params: Params, clip_class: Sequence[int], n_classes: int,
                        out: Optional[torch.Tensor] = None, acc: Optional[torch.Tensor] = None):
    """featuregram + moments in one library call (top_db clip and moments share one pass).
    Returns (out, acc) with ``acc`` laid out as in :func:`moments` (accumulated into when given)."""
    rows = feature_rows(params)
    D = rows
    if out is None:
        out = torch.empty(rows * batch.total_frames, dtype=torch.float32, device=wave.device)
    n = n_classes * D + D + n_classes + 1
    if acc is None:
        acc = torch.zeros(n, dtype=torch.float64, device=wave.device)
    cls = np.ascontiguousarray(clip_class, dtype=np.int32)
    if cls.size != batch.n_clips:
        raise ValueError("clip_class must have one entry per clip")
    base = acc.data_ptr()
    check(batch.lib.hpss_featuregram_moments(batch.ctx.handle, batch.handle, _dev_ptr(wave, torch.float32, "wave"),
                                             C.byref(params), _dev_ptr(out, torch.float32, "out"),
                                             C.c_void_p(cls.ctypes.data), int(cls.size), int(n_classes), C.c_void_p(base),
                                             C.c_void_p(base + 8 * n_classes * D),
                                             C.c_void_p(base + 8 * (n_classes * D + D)),
                                             C.c_void_p(base + 8 * (n_classes * D + D + n_classes)), _stream_ptr()))
    return out, acc


def stats_finalize(acc_host: np.ndarray, D: int, n_classes: int):
    acc_host = np.ascontiguousarray(acc_host, dtype=np.float64)
    mean = np.empty(D, dtype=np.float32)
    std = np.empty(D, dtype=np.float32)
    b = acc_host.ctypes.data
    check(_lib.load().hpss_stats_finalize(C.c_void_p(b), C.c_void_p(b + 8 * n_classes * D),
                                          C.c_void_p(b + 8 * (n_classes * D + D)), int(D), int(n_classes),
                                          C.c_void_p(mean.ctypes.data), C.c_void_p(std.ctypes.data)))
    counts = acc_host[n_classes * D + D:n_classes * D + D + n_classes].copy()
    return mean, std, counts, float(acc_host[-1])


def scale_data(batch: Batch, feat: torch.Tensor, D: int, mean: torch.Tensor, stdev: torch.Tensor, eps: float = 1e-10):
    out = torch.empty(feat.numel(), dtype=torch.float64, device=feat.device)
    check(batch.lib.hpss_scale_data(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                    _dev_ptr(mean, torch.float32, "mean"), _dev_ptr(stdev, torch.float32, "stdev"),
                                    float(eps), _dev_ptr(out), _stream_ptr()))
    return out


def scale_data_f32(batch: Batch, feat: torch.Tensor, D: int, mean: torch.Tensor, stdev: torch.Tensor) -> torch.Tensor:
    """numpy's float32 evaluation of (FV - mean) / stdev (lib/preprocessing.py:590-614), bit-identical."""
    out = torch.empty(feat.numel(), dtype=torch.float32, device=feat.device)
    check(batch.lib.hpss_scale_data_f32(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                        _dev_ptr(mean, torch.float32, "mean"), _dev_ptr(stdev, torch.float32, "stdev"),
                                        _dev_ptr(out), _stream_ptr()))
    return out


def row_standardize(batch: Batch, feat: torch.Tensor, D: int) -> torch.Tensor:
    check(batch.lib.hpss_row_standardize(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                         _stream_ptr()))
    return feat


def num_patches(n_frames: int, patch_size: int, patch_shift: int) -> int:
    return int(_lib.load().hpss_num_patches(int(n_frames), int(patch_size), int(patch_shift)))


def extract_patches(ctx: Context, feat: torch.Tensor, patch_size: int, patch_shift: int) -> torch.Tensor:
    """feat: CUDA float32 (D, T) -> float64 (n_patches, D, patch_size)."""
    D, T = int(feat.shape[0]), int(feat.shape[1])
    n = num_patches(T, patch_size, patch_shift)
    out = torch.empty((n, D, patch_size), dtype=torch.float64, device=feat.device)
    if n:
        check(ctx.lib.hpss_extract_patches(ctx.handle, _dev_ptr(feat, torch.float32, "feat"), D, T, int(patch_size),
                                           int(patch_shift), _dev_ptr(out), _stream_ptr()))
    return out


def num_patches_tiled(n_frames: int, patch_size: int, patch_shift: int) -> int:
    return int(_lib.load().hpss_num_patches_tiled(int(n_frames), int(patch_size), int(patch_shift)))


def patch_offsets(batch: Batch, patch_size: int, patch_shift: int) -> np.ndarray:
    out = np.zeros(batch.n_clips + 1, dtype=np.int64)
    check(batch.lib.hpss_patch_offsets(batch.handle, int(patch_size), int(patch_shift),
                                       out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


def patch_tensor(batch: Batch, feat: torch.Tensor, D: int, patch_size: int, patch_shift: int, standardize: bool = True,
                 rows: str = "all", time_major: bool = False, dtype=torch.float32) -> torch.Tensor:
    """N1 on the device: featuregrams of a batch -> the model-ready patch tensor (n_patches, n_rows, W), or
    (n_patches, W, n_rows) with ``time_major`` (the TCN layout).  ``rows``: "all" | "harm" | "perc".
    ``standardize`` applies the per-file StandardScaler IN PLACE on ``feat`` first.  The result is a torch CUDA tensor
    (hand it on with torch.utils.dlpack.to_dlpack / __dlpack__)."""
    half = D // 2
    row0, n_rows = {"all": (0, D), "harm": (0, half), "perc": (half, D - half)}[rows]
    off = patch_offsets(batch, patch_size, patch_shift)
    n = int(off[-1])
    shape = (n, patch_size, n_rows) if time_major else (n, n_rows, patch_size)
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    out = torch.empty(shape, dtype=dtype, device=feat.device)
    if feat.dtype == torch.float64:                       # frame-level-scaled float64 featuregrams: exact float64 gather
        if standardize or dtype != torch.float64:
            raise ValueError("float64 featuregrams are gathered as they are (no standardisation, float64 output)")
        check(batch.lib.hpss_patch_tensor_f64(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float64, "feat"), int(D),
                                              int(row0), int(n_rows), int(patch_size), int(patch_shift),
                                              int(bool(time_major)), _dev_ptr(out), _stream_ptr()))
        return out
    check(batch.lib.hpss_patch_tensor(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                      int(bool(standardize)), int(row0), int(n_rows), int(patch_size), int(patch_shift),
                                      int(bool(time_major)), int(dtype == torch.float64), _dev_ptr(out), _stream_ptr()))
    return out


def row_nonfinite(batch: Batch, feat: torch.Tensor, D: int) -> torch.Tensor:
    flags = torch.empty(max(1, batch.n_clips * D), dtype=torch.uint8, device=feat.device)
    check(batch.lib.hpss_row_nonfinite(batch.ctx.handle, batch.handle, _dev_ptr(feat, torch.float32, "feat"), int(D),
                                       _dev_ptr(flags), _stream_ptr()))
    return flags[:batch.n_clips * D].view(batch.n_clips, D)


STATS = {"mean": 0, "variance": 1, "skew": 2, "kurtosis": 3}


def patch_statistics(ctx: Context, patches: torch.Tensor, stat_type: str = "skew", axis: int = 0) -> torch.Tensor:
    """get_data_statistics (tools.pyx:169-211) on the device: patches (N, f, t) float64 -> (N, t) for axis=0,
    (N, f) for axis=1."""
    if patches.dim() != 3:
        raise ValueError("patches must be (N, nFeat, nFrames)")
    N, A, B = (int(x) for x in patches.shape)
    out = torch.empty((N, B if axis == 0 else A), dtype=torch.float64, device=patches.device)
    check(ctx.lib.hpss_patch_statistics(ctx.handle, _dev_ptr(patches, torch.float64, "patches"), N, A, B,
                                        STATS[stat_type], int(axis), _dev_ptr(out), _stream_ptr()))
    return out
