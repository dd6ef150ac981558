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


def median_time(batch: Batch, S: torch.Tensor, rows: int, k: int) -> torch.Tensor:
    out = torch.empty_like(S)
    check(batch.lib.hpss_median_time(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"), int(rows), int(k),
                                     _dev_ptr(out), _stream_ptr()))
    return out


def median_freq(batch: Batch, S: torch.Tensor, rows: int, k: int) -> torch.Tensor:
    out = torch.empty_like(S)
    check(batch.lib.hpss_median_freq(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"), int(rows), int(k),
                                     _dev_ptr(out), _stream_ptr()))
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


def perc_mask_mel_log(batch: Batch, S: torch.Tensor, harm: torch.Tensor, rows: int, k: int, mel_sr: int = 22050,
                      n_mels: int = 0, log_power: int = 0, amin: float = 1e-10):
    """K2p + K3 fused (frequency median + masks + mel + log); returns (out, clip_max or None)."""
    rows_out = 2 * (n_mels if n_mels > 0 else rows)
    out = torch.empty(rows_out * batch.total_frames, dtype=torch.float32, device=S.device)
    clip_max = torch.empty(2 * max(1, batch.n_clips), dtype=torch.int32, device=S.device) if log_power else None
    check(batch.lib.hpss_perc_mask_mel_log(batch.ctx.handle, batch.handle, _dev_ptr(S, torch.float32, "S"),
                                           _dev_ptr(harm, torch.float32, "harm"), int(rows), int(k), int(mel_sr),
                                           int(n_mels), int(log_power), float(amin), _dev_ptr(out), _dev_ptr(clip_max),
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


def host_alloc(n_floats: int) -> np.ndarray:
    """float32 numpy array over pinned host memory owned by the library (freed with the array)."""
    lib = _lib.load()
    p = C.c_void_p()
    check(lib.hpss_host_alloc(C.byref(p), int(n_floats) * 4))
    buf = (C.c_float * int(n_floats)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float32)
    weakref.finalize(arr, lib.hpss_host_free, C.c_void_p(p.value))    # views keep `arr` alive via .base
    return arr


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
                                 C.c_void_p(cls.ctypes.data), int(n_classes), C.c_void_p(base),
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
    base = acc.data_ptr()
    check(batch.lib.hpss_topdb_moments(batch.ctx.handle, batch.handle, _dev_ptr(out, torch.float32, "out"),
                                       int(rows_per_stream), int(n_streams), _dev_ptr(clip_max), float(top_db),
                                       C.c_void_p(cls.ctypes.data), int(n_classes), C.c_void_p(base),
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
                                             C.c_void_p(cls.ctypes.data), int(n_classes), C.c_void_p(base),
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
