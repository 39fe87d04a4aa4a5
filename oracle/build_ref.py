#!/usr/bin/env python
"""Build recipe for oracle/_ref -- the UNMODIFIED reference kernels, compiled here.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (tfce_mediation_b200); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may use it, as the checker or
as the CPU baseline -- never as the thing shipped.

What it does
------------
Cythonizes the reference's own hot-path extension modules from the sources
where they lie under /root/reference (read-only):

    tfce_mediation/tfce.pyx  (+ lib/fast_tfce.hpp)   -> oracle/_ref/tfce*.so
    tfce_mediation/cynumstats.pyx                     -> oracle/_ref/cynumstats*.so

The reference's own build (tfce_mediation/setup.py:4-5) needs numpy.distutils,
which no longer exists under numpy 2.x, so the two extensions are built with a
plain cythonize() call using the reference's compile flags
(tfce_mediation/setup.py:25-41: -std=c++11 -Wno-unused -g).  The .pyx files are
staged into a scratch directory under /tmp because Cython writes its generated
.cpp next to the source and /root/reference is read-only.  No reference source
is copied into the repository: only the compiled shared objects land in
oracle/_ref/ (git-ignored; they travel to the GPU box with the gpurun snapshot).

/root/reference does not exist on the GPU box: this script is a no-op there
(the prebuilt .so files are used).
"""
import glob
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TFCE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def have_ref():
    return (glob.glob(os.path.join(OUT, "tfce.*so")) and
            glob.glob(os.path.join(OUT, "cynumstats.*so")))


def build(force=False):
    src = os.path.join(REF_ROOT, "tfce_mediation")
    if not os.path.isdir(src):
        return bool(have_ref())
    if have_ref() and not force:
        return True
    os.makedirs(OUT, exist_ok=True)
    import numpy
    from setuptools import Extension, setup
    from Cython.Build import cythonize

    stage = tempfile.mkdtemp(prefix="tfce_ref_build_")
    try:
        for f in ("tfce.pyx", "cynumstats.pyx"):
            shutil.copy(os.path.join(src, f), os.path.join(stage, f))
        flags = ["-std=c++11", "-Wno-unused", "-g", "-O2"]
        exts = [
            Extension("tfce", [os.path.join(stage, "tfce.pyx")], language="c++",
                      include_dirs=[os.path.join(src, "lib"), numpy.get_include()],
                      extra_compile_args=flags),
            Extension("cynumstats", [os.path.join(stage, "cynumstats.pyx")], language="c++",
                      include_dirs=[numpy.get_include()],
                      extra_compile_args=flags),
        ]
        cwd = os.getcwd()
        os.chdir(stage)
        try:
            setup(name="tfce_ref", ext_modules=cythonize(exts, language_level=3, quiet=True),
                  script_args=["-q", "build_ext", "--build-lib", OUT,
                               "--build-temp", os.path.join(stage, "tmp")])
        finally:
            os.chdir(cwd)
    finally:
        shutil.rmtree(stage, ignore_errors=True)
    return bool(have_ref())


def load():
    """Import the compiled reference modules -> (tfce_module, cynumstats_module) or None."""
    if not have_ref():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    return importlib.import_module("tfce"), importlib.import_module("cynumstats")


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ok" if ok else "unavailable")
