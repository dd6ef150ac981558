"""CPU tests of the host side: the C-ABI library loads and exports what the header declares, the
host-only entry points agree with the oracle, compute calls fail loudly without a GPU, the
reference-signature mirror's host logic (name dispatch, signal preparation) matches the golden
vectors made by the reference's own code, clip sharding and the moments all-reduce (gloo, world 2)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import preprocessing_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_glue.npz")


@pytest.fixture(scope="module")
def lib():
    from sm_hpss_mtl_b200 import _lib
    return _lib.load()


# ------------------------------------------------------------------ C ABI surface
def test_library_exports_every_header_symbol(lib):
    from sm_hpss_mtl_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hpss_b200.h")).read()
    declared = set(re.findall(r"HPSS_API\s+[\w\s\*]+?\b(hpss_\w+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hpss_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header are out of sync"
    assert b"sm_100a" in lib.hpss_version()


def test_params_struct_layout():
    from sm_hpss_mtl_b200._lib import Params
    assert C.sizeof(Params) == 8 * 4 + 2 * 4
    assert [f[0] for f in Params._fields_] == ["n_fft", "win_length", "hop_length", "l_harm", "l_perc", "n_mels",
                                               "mel_sr", "feature", "amin", "top_db"]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created and the Python API refuses to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sm_hpss_mtl_b200 import _lib, engine
    h = C.c_void_p()
    rc = lib.hpss_ctx_create(0, C.byref(h))
    assert rc == _lib.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.hpss_last_error()
    with pytest.raises(RuntimeError):
        engine.get_context()
    from sm_hpss_mtl_b200 import preprocessing as pp
    with pytest.raises(RuntimeError):
        pp.featuregram_from_signal(np.zeros(16000, np.float32), 16000, {"Tw": 25, "Ts": 10, "Model": "m",
                                   "l_harm": {"m": 21}, "l_perc": {"m": 11}}, 400, 120, "LogMelHarmPercSpec")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sm_hpss_mtl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


# ------------------------------------------------------------------ host-only entry points
@pytest.mark.parametrize("sr,n_fft,n_mels", [(22050, 400, 120), (16000, 400, 120), (22050, 512, 21), (22050, 2048, 128),
                                             (22050, 400, 40)])
def test_mel_filterbank_matches_oracle(sr, n_fft, n_mels):
    from sm_hpss_mtl_b200 import engine
    got, want = engine.mel_filterbank(sr, n_fft, n_mels), lr.mel(sr, n_fft, n_mels)
    assert got.shape == want.shape and float(np.abs(got - want).max()) <= 1e-7
    assert np.array_equal(got > 0, want > 0)


@pytest.mark.parametrize("n_fft,win", [(400, 400), (512, 400), (2048, 2048), (512, 401)])
def test_stft_window_matches_oracle(n_fft, win):
    from sm_hpss_mtl_b200 import engine
    want = lr.pad_center(lr.hann_periodic(win), n_fft).astype(np.float32)
    assert float(np.abs(engine.stft_window(n_fft, win) - want).max()) <= 6e-8


def test_num_patches_and_feature_rows(lib):
    from sm_hpss_mtl_b200 import engine
    for T in (68, 69, 98, 249, 250, 998, 5000):
        for (W, sh) in [(249, 24), (68, 68), (99, 34), (99, 1), (21, 5)]:
            assert engine.num_patches(T, W, sh) == len(range(W // 2, T - W // 2, sh))
    rows = {"SPEC": 201, "LOGSPEC": 201, "MELSPEC": 120, "LOGMELSPEC": 120, "HARMPERC": 402, "LOG_HARMPERC": 402,
            "MEL_HARMPERC": 240, "LOGMEL_HARMPERC": 240}
    for f, r in rows.items():
        assert engine.feature_rows(engine.make_params(feature=f)) == r


def test_stats_finalize_closed_form_matches_two_pass():
    from sm_hpss_mtl_b200 import engine
    rng = np.random.default_rng(0)
    D, names = 17, ["music", "speech", "speech_music"]
    groups = {n: [(rng.standard_normal((D, int(rng.integers(20, 90)))) * 6 - 35).astype(np.float32) for _ in range(3)]
              for n in names}
    acc = np.zeros(3 * D + D + 3 + 1)
    for k, n in enumerate(names):
        for fv in groups[n]:
            acc[k * D:(k + 1) * D] += fv.astype(np.float64).sum(axis=1)
            acc[3 * D:4 * D] += (fv.astype(np.float64) ** 2).sum(axis=1)
            acc[4 * D + k] += fv.shape[1]
    mean, std, counts, bad = engine.stats_finalize(acc, D, 3)
    want_mean, want_std, n0, n1, n2 = po.get_data_stats(groups, names)
    assert list(counts) == [n0, n1, n2] and bad == 0
    assert np.allclose(mean, want_mean, rtol=1e-6, atol=1e-6) and np.allclose(std, want_std, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ reference-signature mirror (host logic)
def test_feature_name_dispatch():
    from sm_hpss_mtl_b200.preprocessing import feature_family
    table = {"Spec": "SPEC", "LogSpec": "LOGSPEC", "MelSpec": "MELSPEC", "LogMelSpec": "LOGMELSPEC",
             "HarmSpec": "HARMPERC", "PercSpec": "HARMPERC", "HarmPercSpec": "HARMPERC",
             "LogHarmSpec": "LOG_HARMPERC", "LogPercSpec": "LOG_HARMPERC", "LogHarmPercSpec": "LOG_HARMPERC",
             "MelHarmSpec": "MEL_HARMPERC", "MelPercSpec": "MEL_HARMPERC", "MelHarmPercSpec": "MEL_HARMPERC",
             "LogMelHarmSpec": "LOGMEL_HARMPERC", "LogMelPercSpec": "LOGMEL_HARMPERC",
             "LogMelHarmPercSpec": "LOGMEL_HARMPERC"}
    for name, fam in table.items():
        assert feature_family(name)[0] == fam
        assert feature_family(name)[1] == (name in ("MelSpec", "LogMelSpec"))      # only these pass sr=fs
    with pytest.raises(ValueError):
        feature_family("MFCC")                                                     # not in the reference


def test_oracle_signal_preparation_matches_reference():
    """The oracle's restatement of normalise -> RMS -> silence removal (Cython semantics, incl. the tail of ones,
    "more than one stretch", the doubling below 0.1 s) -> normalise, and of SMR mixing, against what the reference's
    own functions (lib/preprocessing.py + its compiled Cython leaf) produced: signals AND the gate's markers."""
    g = np.load(GOLDEN)
    audio = {k[len("audio:"):]: g[k] for k in g.files if k.startswith("audio:")}
    assert len(audio) == 6
    for path, x in audio.items():
        got, smark, fmark, energy = po.load_and_preprocess_signal(x, 25, 10, details=True)
        assert got.shape == g["prep:" + path].shape
        assert np.array_equal(got.astype(np.float32), g["prep:" + path]), path
        assert np.array_equal(smark, g["gate:sample:" + path]) and np.array_equal(fmark, g["gate:frame:" + path])
    assert g["prep:/d/speech/short.wav"].size == 2 * 960                     # doubled once: 0.06 s -> 0.12 s
    assert (g["gate:sample:/d/music/onesil.wav"] == 0).any()                 # one stretch: marked ...
    y = po.normalize_signal(audio["/d/music/onesil.wav"])
    assert np.allclose(g["prep:/d/music/onesil.wav"], po.normalize_signal(y), atol=1e-7)   # ... but not removed
    mix = po.mix_signals(g["prep:/d/speech/sp0.wav"], g["prep:/d/music/mu0.wav"], 5)
    assert np.allclose(mix, g["mix:sp0+mu0@5"], rtol=0, atol=1e-7)
    pat = g["patch:Lemaire_et_al_MTL:LogMelHarmPercSpec:49:24"]
    for st in ("mean", "variance", "skew", "kurtosis"):
        for ax in (0, 1):
            assert np.array_equal(po.get_data_statistics(pat, st, ax), g[f"pstat:{st}:{ax}"])


def test_prep_lengths_host_side(lib):
    """hpss_prep_out_length / hpss_prep_num_frames (host-only entries) against the reference's rules."""
    from sm_hpss_mtl_b200 import engine
    for n in (2, 100, 799, 800, 960, 1599, 1600, 1601, 16000, 57_600_000):
        want = n
        while want / 16000 < 0.1:                                            # lib/preprocessing.py:345-347
            want *= 2
        assert engine.prep_out_length(n, 16000) == want
        assert engine.prep_num_frames(n, 400, 160) == len(po.frame_rms(np.zeros(n, np.float32), 400, 160)) if n < 1_000_000 \
            else engine.prep_num_frames(n, 400, 160) == 1 + n // 160
    for T, W, sh in [(98, 249, 24), (249, 249, 24), (248, 249, 24), (5, 68, 68), (300, 68, 68), (67, 68, 68), (68, 68, 1)]:
        Tt = T
        if T < W:
            while Tt <= W:                                                   # lib/preprocessing.py:139-142
                Tt += T
        assert engine.num_patches_tiled(T, W, sh) == len(range(W // 2, Tt - W // 2, sh))


def test_load_audio_wav_roundtrip(tmp_path):
    from scipy.io import wavfile
    from sm_hpss_mtl_b200 import preprocessing as pp
    x = (np.random.default_rng(0).standard_normal(4000) * 3000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, x)
    y = pp.load_audio(p)
    assert y.dtype == np.float32 and np.array_equal(y, x.astype(np.float32) / 32768.0)
    z = pp.load_pcm(p)
    assert z.dtype == np.int16 and np.array_equal(z, x)                      # 16-bit PCM stays int16 for the upload
    wavfile.write(p, 8000, x)
    assert pp.load_pcm(p).dtype == np.float32 and abs(len(pp.load_pcm(p)) - 8000) <= 1      # resampled to 16 kHz


def test_stat_jobs_and_cache_names():
    """File lists of get_data_stats and the cache-file stems, incl. the 5-class script's noise variants."""
    from sm_hpss_mtl_b200 import preprocessing as pp
    P = {"classes": {0: "music", 1: "speech", 2: "speech_music"}, "folder": "/d"}
    files = {"music": ["m0.wav"], "speech": ["s0.wav"], "speech+music": [{"speech": "s0.wav", "music": "m0.wav", "SMR": 5}]}
    jobs = pp._stat_jobs(P, files)
    assert jobs == [("music", "", "/d/music/m0.wav", -1), ("speech", "/d/speech/s0.wav", "", -1),
                    ("speech_music", "/d/speech/s0.wav", "/d/music/m0.wav", 5)]
    assert pp._feature_name_of_file("/d/speech/s0.wav", "/d/music/m0.wav", 5) == "s0_m0_5dB"
    assert pp._feature_name_of_file("/d/speech/s0.wav", "", 10, "/d/noise/n1.wav") == "s0_n1_10dB"
    assert pp._feature_name_of_file("", "", -1, "/d/noise/n1.wav") == "n1"
    assert pp._sources("speech_noise", "a", "b", "c") == ("a", "c") and pp._sources("noise", "a", "b", "c") == ("c", None)


# ------------------------------------------------------------------ sharding + the one collective
def test_shard_clips_properties():
    from sm_hpss_mtl_b200.dist import shard_clips
    rng = np.random.default_rng(1)
    for n, w in [(1086, 8), (4096, 8), (5, 8), (1, 1), (100, 3), (0, 4)]:
        lens = rng.integers(1600, 5_000_000, size=n)
        sh = shard_clips(lens, w)
        assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == n
        assert all(a <= b for a, b in sh) and all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
        if n >= 50 * w:
            tot = [int(lens[a:b].sum()) for a, b in sh]
            assert max(tot) - min(tot) <= 2 * int(lens.max())
    assert shard_clips([16000] * 4096, 8) == [(i * 512, (i + 1) * 512) for i in range(8)]


def test_stat_jobs_sharding_covers_every_file_once():
    from sm_hpss_mtl_b200.preprocessing import _shard_jobs
    for n, world in [(1086, 8), (7, 8), (0, 4), (100, 3), (5, 1)]:
        jobs = list(range(n))
        parts = [_shard_jobs(jobs, r, world) for r in range(world)]
        assert sum(parts, []) == jobs                                        # order kept, nothing lost or repeated
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= max(1, -(-n // world))


def test_stream_shard_properties():
    """Time split of one long stream (configs[3]): the owned frame ranges tile [0, T), every shard's computed range
    holds l_harm // 2 halo frames per side (clamped at the stream ends) and its sample range covers exactly them."""
    from sm_hpss_mtl_b200.dist import stream_shard
    for (L, n_fft, hop, k, world) in [(57_600_000, 2048, 512, 31, 8), (640_000, 2048, 512, 31, 3), (160_000, 400, 160, 21, 2),
                                      (16_000, 400, 160, 63, 4)]:
        T = 1 + (L - n_fft) // hop
        prev = 0
        for r in range(world):
            (t0, t1), (a, b), (s0, s1) = stream_shard(L, n_fft, hop, k, r, world)
            assert t0 == prev and t1 > t0
            prev = t1
            assert a == max(0, t0 - k // 2) and b == min(T, t1 + k // 2)
            assert s0 == a * hop and s1 == (b - 1) * hop + n_fft and s1 <= L
            assert 1 + (s1 - s0 - n_fft) // hop == b - a
        assert prev == T
    with pytest.raises(ValueError):
        stream_shard(2048, 2048, 512, 31, 0, 2)


def _gloo_worker(rank, world, port, D, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sm_hpss_mtl_b200 import dist as hd
    from sm_hpss_mtl_b200.dist import shard_clips
    rng = np.random.default_rng(7)
    fvs = [(rng.standard_normal((D, int(rng.integers(10, 60)))) * 5 - 30).astype(np.float32) for _ in range(12)]
    cls = [i % 3 for i in range(12)]
    a, b = shard_clips([fv.shape[1] for fv in fvs], world)[rank]
    acc = np.zeros(hd.moments_size(D, 3))
    for fv, k in zip(fvs[a:b], cls[a:b]):                      # what hpss_moments accumulates on a GPU
        acc[k * D:(k + 1) * D] += fv.astype(np.float64).sum(axis=1)
        acc[3 * D:4 * D] += (fv.astype(np.float64) ** 2).sum(axis=1)
        acc[4 * D + k] += fv.shape[1]
    t = torch.from_numpy(acc)
    hd.allreduce_moments(t)
    mean, std, counts = hd.finalize_stats(t.numpy(), D, 3)
    if rank == 0:
        q.put((mean, std, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_moments_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    D, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    mean, std, counts = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    fvs = [(rng.standard_normal((D, int(rng.integers(10, 60)))) * 5 - 30).astype(np.float32) for _ in range(12)]
    names = ["music", "speech", "speech_music"]
    groups = {n: [fvs[i] for i in range(12) if i % 3 == k] for k, n in enumerate(names)}
    want_mean, want_std, n0, n1, n2 = po.get_data_stats(groups, names)
    assert [int(c) for c in counts] == [n0, n1, n2]
    assert np.allclose(mean, want_mean, rtol=1e-6, atol=1e-6) and np.allclose(std, want_std, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("M,n", [(120, 20), (21, 21), (128, 13)])
def test_dct_basis_table_vs_scipy(M, n):
    """hpss_dct_basis (host-side table, no GPU needed) against scipy's orthonormal DCT-II of the identity."""
    import scipy.fft
    from sm_hpss_mtl_b200 import engine
    want = scipy.fft.dct(np.eye(M), axis=0, type=2, norm="ortho")[:n]
    got = engine.dct_basis(M, n)
    assert got.shape == (n, M) and got.dtype == np.float32
    assert np.abs(got - want).max() < 1e-7
    with pytest.raises(Exception):
        engine.dct_basis(M, M + 1)
