"""ctypes binding of libhpss_b200.so (the C ABI declared in include/hpss_b200.h).

There is no CPU fallback: if the shared object is missing the import fails loudly, and every
compute entry point fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("HPSS_B200_LIB", _PKG / "libhpss_b200.so"))


class HpssError(RuntimeError):
    """Non-zero status from libhpss_b200 (message from hpss_last_error)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[hpss_b200 status {status}] {message}")
        self.status = status
        self.message = message


class ParameterError(ValueError):
    """Mirror of librosa.util.exceptions.ParameterError (what the reference's callers see)."""


OK, ERR_INVALID, ERR_SHORT_SIGNAL, ERR_UNSUPPORTED, ERR_CUDA, ERR_NEGATIVE, ERR_NOMEM, ERR_NONFINITE = range(8)
PCM_F32, PCM_S16 = 0, 1

FEATURES = {
    "SPEC": 0, "LOGSPEC": 1, "MELSPEC": 2, "LOGMELSPEC": 3,
    "HARMPERC": 4, "LOG_HARMPERC": 5, "MEL_HARMPERC": 6, "LOGMEL_HARMPERC": 7,
}


class Params(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("l_harm", C.c_int32), ("l_perc", C.c_int32), ("n_mels", C.c_int32),
        ("mel_sr", C.c_int32), ("feature", C.c_int32), ("amin", C.c_float), ("top_db", C.c_float),
    ]


_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes).  Every symbol include/hpss_b200.h declares is listed here; the
# CPU test-suite checks the two stay in sync.
PROTOTYPES = {
    "hpss_version": (C.c_char_p, []),
    "hpss_last_error": (C.c_char_p, []),
    "hpss_launch_count": (_u64, []),
    "hpss_ctx_create": (C.c_int, [C.c_int, _pp]),
    "hpss_ctx_destroy": (C.c_int, [_vp]),
    "hpss_ctx_device": (C.c_int, [_vp]),
    "hpss_ctx_workspace_bytes": (_u64, [_vp]),
    "hpss_host_alloc": (C.c_int, [_pp, _u64]),
    "hpss_host_free": (C.c_int, [_vp]),
    "hpss_batch_from_samples": (C.c_int, [_vp, _pi64, _i32, _i32, _i32, _pp]),
    "hpss_batch_from_frames": (C.c_int, [_vp, _pi64, _i32, _pp]),
    "hpss_batch_destroy": (C.c_int, [_vp]),
    "hpss_batch_n_clips": (_i32, [_vp]),
    "hpss_batch_total_frames": (_i64, [_vp]),
    "hpss_batch_total_samples": (_i64, [_vp]),
    "hpss_batch_frame_offsets": (C.c_int, [_vp, _pi64]),
    "hpss_batch_sample_offsets": (C.c_int, [_vp, _pi64]),
    "hpss_mel_filterbank": (C.c_int, [_i32, _i32, _i32, _vp]),
    "hpss_stft_window": (C.c_int, [_i32, _i32, _vp]),
    "hpss_stft_mag": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hpss_median_time": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "hpss_median_freq": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "hpss_mask_mel_log": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp]),
    "hpss_mask_mel_log_sr": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp]),
    "hpss_topdb_clip": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _f32, _vp]),
    "hpss_feature_rows": (_i32, [C.POINTER(Params)]),
    "hpss_featuregram": (C.c_int, [_vp, _vp, _vp, C.POINTER(Params), _vp, _vp]),
    "hpss_featuregram_from_spec": (C.c_int, [_vp, _vp, _vp, _i32, C.POINTER(Params), _vp, _vp]),
    "hpss_featuregram_host": (C.c_int, [_vp, _vp, _vp, C.POINTER(Params), _vp]),
    "hpss_moments": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "hpss_topdb_moments": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _f32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "hpss_featuregram_moments": (C.c_int, [_vp, _vp, _vp, C.POINTER(Params), _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "hpss_ctx_check": (C.c_int, [_vp, _vp]),
    "hpss_validate_audio": (C.c_int, [_vp, _vp, _i64, _vp]),
    "hpss_validate_nonneg": (C.c_int, [_vp, _vp, _i64, _vp]),
    "hpss_pipeline_create": (C.c_int, [_vp, _pi64, _i32, C.POINTER(Params), _i32, _i32, _i32, _f64, _f64, _i32, _pp]),
    "hpss_pipeline_destroy": (C.c_int, [_vp]),
    "hpss_pipeline_total_frames": (_i64, [_vp]),
    "hpss_pipeline_n_chunks": (_i32, [_vp]),
    "hpss_pipeline_frame_offsets": (C.c_int, [_vp, _pi64]),
    "hpss_pipeline_run": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "hpss_prep_out_length": (_i64, [_i64, _i32]),
    "hpss_prep_num_frames": (_i64, [_i64, _i32, _i32]),
    "hpss_prep_signals": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp, _vp, _vp, _vp]),
    "hpss_mix_signals": (C.c_int, [_vp, _vp, _pi64, _vp, _pi64, _vp, _i32, _vp, _vp]),
    "hpss_stats_finalize": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "hpss_scale_data": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _f64, _vp, _vp]),
    "hpss_scale_data_f32": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "hpss_patch_tensor_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hpss_row_standardize": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "hpss_num_patches": (_i64, [_i64, _i32, _i32]),
    "hpss_extract_patches": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "hpss_num_patches_tiled": (_i64, [_i64, _i32, _i32]),
    "hpss_patch_offsets": (C.c_int, [_vp, _i32, _i32, _pi64]),
    "hpss_patch_tensor": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hpss_row_nonfinite": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "hpss_patch_statistics": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hpss_dct_mfcc": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "hpss_dct_basis": (C.c_int, [_i32, _i32, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared object once; raise ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            f"(python -m sm_hpss_mtl_b200.build, or __graft_entry__.build()). "
            f"sm_hpss_mtl_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().hpss_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status == OK:
        return
    msg = last_error()
    if status in (ERR_SHORT_SIGNAL, ERR_NEGATIVE, ERR_NONFINITE):
        raise ParameterError(msg)          # what librosa raises at the same place
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    raise HpssError(status, msg)
