"""Builds harness/libb200harness.so (the C test harness, see harness/__init__.py) against petsc_b200/lib/libpetscb200.so."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GCC = "/usr/bin/gcc"
LIBDIR = os.path.join(ROOT, "petsc_b200", "lib")
SO = os.path.join(HERE, "libb200harness.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(verbose=False, force=False):
    hsrc = sorted(glob.glob(os.path.join(HERE, "host", "*.c")))
    hdrs = glob.glob(os.path.join(HERE, "host", "*.h")) + glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _newer(SO, hsrc + hdrs + [os.path.join(LIBDIR, "libpetscb200.so")]):
        cmd = [GCC, "-O2", "-g", "-fPIC", "-std=c11", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Wno-format-truncation", "-shared", "-o", SO] + hsrc + [
            "-I", os.path.join(ROOT, "include"), "-I", HERE, "-I", os.path.join(HERE, "host"), "-L", LIBDIR, "-lpetscb200",
            "-Wl,-rpath,$ORIGIN/../petsc_b200/lib", "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
