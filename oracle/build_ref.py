"""Compile the reference's only native leaf into oracle/_ref/ (test infrastructure only).

``lib/cython_impl/tools.pyx`` (extract_patches, scale_data, get_data_statistics, removeSilence)
is compiled from where it lies under /root/reference; generated C, objects and the extension
module all land in oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).  No
reference source is copied into the repository.  Used by tests to pin the oracle's restatement
of extract_patches / scale_data against the reference's own code.

    python -m oracle.build_ref
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF_PYX = Path("/root/reference/lib/cython_impl/tools.pyx")


def built_module_path():
    if not OUT.exists():
        return None
    for p in sorted(OUT.glob("tools*.so")):
        return p
    return None


def build(force: bool = False):
    """Returns the path of the built extension, or None when the reference tree is absent."""
    have = built_module_path()
    if have is not None and not force:
        return have
    if not REF_PYX.exists():
        return None
    import numpy
    from Cython.Build import cythonize
    from setuptools import Distribution, Extension
    from setuptools.command.build_ext import build_ext

    OUT.mkdir(exist_ok=True)
    ext = Extension("tools", [str(REF_PYX)], include_dirs=[numpy.get_include()],
                    define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
    exts = cythonize([ext], build_dir=str(OUT / "gen"), language_level=3, quiet=True)
    dist = Distribution({"name": "sm_hpss_ref_tools", "ext_modules": exts})
    cmd = build_ext(dist)
    cmd.build_lib = str(OUT)
    cmd.build_temp = str(OUT / "tmp")
    cmd.inplace = False
    cmd.ensure_finalized()
    cwd = os.getcwd()
    try:
        os.chdir(OUT)          # keep every by-product inside oracle/_ref
        cmd.run()
    finally:
        os.chdir(cwd)
    return built_module_path()


def load():
    """Import the compiled reference module (None if it has not been built)."""
    p = built_module_path()
    if p is None:
        return None
    spec = importlib.util.spec_from_file_location("tools", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
