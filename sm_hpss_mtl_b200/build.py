"""Build libhpss_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m sm_hpss_mtl_b200.build [--force]

The shared object lands next to this file so that it travels with the repository snapshot to
the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libhpss_b200.so"
OBJ = PKG / "_build"
SOURCES = ["api.cu", "stft.cu", "stft_fast.cu", "median.cu", "median_walk.cu", "maskmel.cu", "stats.cu", "dct.cu", "prep.cu"]
GEN_HEADER = CSRC / "median_networks_gen.cuh"
GENERATOR = ROOT / "tools" / "gen_median_networks.py"
FFT_HEADER = CSRC / "fft_codelets_gen.cuh"
FFT_GENERATOR = ROOT / "tools" / "gen_fft_codelets.py"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "hpss_b200.h",
                                                                          GENERATOR, FFT_GENERATOR, Path(__file__)]):
        if p.name in (GEN_HEADER.name, FFT_HEADER.name):
            continue
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("HPSS_DEV_KS", "").encode())
    return h.hexdigest()


def generate_networks(force: bool = False) -> None:
    """Emit the selection networks.  HPSS_DEV_KS="11,21,31" restricts the generated kernel sizes
    (development builds only: every other k then takes the rank-counting fallback kernel)."""
    ks = os.environ.get("HPSS_DEV_KS", "")
    tag = GEN_HEADER.with_suffix(".ks")
    if force or not GEN_HEADER.exists() or GEN_HEADER.stat().st_mtime < GENERATOR.stat().st_mtime \
            or (tag.read_text() if tag.exists() else "") != ks:
        cmd = [sys.executable, str(GENERATOR), "--out", str(GEN_HEADER)]
        if ks:
            cmd += ["--ks", ks]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tag.write_text(ks)
    if force or not FFT_HEADER.exists() or FFT_HEADER.stat().st_mtime < FFT_GENERATOR.stat().st_mtime:
        subprocess.run([sys.executable, str(FFT_GENERATOR), "--out", str(FFT_HEADER)], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = OBJ / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    OBJ.mkdir(exist_ok=True)
    generate_networks(force)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = OBJ / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
