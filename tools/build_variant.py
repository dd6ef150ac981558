#!/usr/bin/env python3
"""Development helper: build a variant of libhpss_b200.so with extra -D flags for one source file.

    python tools/build_variant.py NAME median.cu -DHPSS_LOADER_WARPS=4 ...
    -> sm_hpss_mtl_b200/_variants/NAME.so   (use with HPSS_B200_LIB=<path>)
The other objects come from the regular build (sm_hpss_mtl_b200/_build).
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sm_hpss_mtl_b200 import build as B  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out_dir = B.PKG / "_variants"
out_dir.mkdir(exist_ok=True)
obj = out_dir / f"{name}.{src}.o"
subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-I", str(ROOT / "include"), "-c", str(B.CSRC / src), "-o", str(obj)],
               check=True)
objs = [str(obj)] + [str(B.OBJ / (s + ".o")) for s in B.SOURCES if s != src]
lib = out_dir / f"{name}.so"
subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), *objs], check=True)
print(lib)
