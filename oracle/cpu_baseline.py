"""CPU arm: the oracle (librosa's algorithm on the scipy/numpy kernels librosa itself calls)
timed on the host cores.  TEST / BENCH INFRASTRUCTURE ONLY -- used by bench.py's
``cpu_baseline`` leg and ``--impl reference``; never by the product path.

Every worker process pins BLAS to one thread (the 120x201xT sgemm is tiny; thread
oversubscription makes it ~50x slower).
"""
from __future__ import annotations

import os
import time

_WAVES = None
_CFG = None


def usable_cores() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:    # cgroup v2 quota
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()
            if q != "max":
                n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return max(1, n)


def _init(cfg, first_index, n_per_worker, counter):
    """Worker initialiser: pin threads, pre-generate this worker's clips (not timed)."""
    global _WAVES, _CFG
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from sm_hpss_mtl_b200 import synth
    with counter.get_lock():
        wid = counter.value
        counter.value += 1
    _CFG = cfg
    _WAVES = [synth.synth_clip(first_index + wid * n_per_worker + i, cfg["clip_samples"]) for i in range(n_per_worker)]
    _run_one(_WAVES[0])      # warm caches / imports


def _run_one(y):
    from oracle import preprocessing_oracle as po
    c = _CFG
    return po.featuregram(y, c["fs"], c["Tw"], c["Ts"], c["l_harm"], c["l_perc"], c["n_fft"], c["n_mels"],
                          c["featName"])


_SPLIT = 4     # tasks per worker and step: dynamic balancing at quarter-worker granularity


def _work(task):
    """Run the oracle over one quarter of this worker's clips (all clips have equal length, so
    every task costs the same whichever worker picks it up)."""
    t0 = time.perf_counter()
    q = task % _SPLIT
    n = len(_WAVES)
    for y in _WAVES[q * n // _SPLIT:(q + 1) * n // _SPLIT]:
        _run_one(y)
    return time.perf_counter() - t0


class CpuArm:
    """Pool of `cores` single-threaded workers, each owning `n_per_worker` pre-generated clips.
    ``step()`` runs the oracle once over all of them and returns the wall seconds."""

    def __init__(self, cfg, cores=None, n_per_worker=16, first_index=0):
        import multiprocessing as mp
        self.cores = cores or usable_cores()
        self.n_per_worker = n_per_worker
        self.cfg = cfg
        ctx = mp.get_context("spawn")          # never fork a process that may hold a CUDA context
        counter = ctx.Value("i", 0)
        self.pool = ctx.Pool(self.cores, initializer=_init, initargs=(cfg, first_index, n_per_worker, counter))
        self.pool.map(lambda_noop, range(self.cores))    # wait until every worker is initialised

    @property
    def clips_per_step(self):
        return self.cores * self.n_per_worker

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_work, range(self.cores * _SPLIT), chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def lambda_noop(_):
    return 0
