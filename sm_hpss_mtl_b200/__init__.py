"""B200-native HPSS feature front-end (drop-in for SM_HPSS_MTL's lib/preprocessing.py).

    sm_hpss_mtl_b200.preprocessing   the reference's function signatures (get_featuregram, get_feature_patches,
                                     get_data_stats, scale_data, ...) on top of the CUDA library
    sm_hpss_mtl_b200.librosa_compat  stft / hpss / melspectrogram / power_to_db with librosa's call signatures
    sm_hpss_mtl_b200.engine          thin ctypes layer over libhpss_b200.so (include/hpss_b200.h), torch tensors in/out
    sm_hpss_mtl_b200.dist            clip sharding and the one collective (moment all-reduce)
    sm_hpss_mtl_b200.build           in-tree nvcc build for sm_100a

There is no CPU fallback: importing the compute modules without the built library, or calling them without a
CUDA device, fails loudly.
"""

__version__ = "0.1"
