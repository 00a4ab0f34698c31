"""Builds libpetscb200.so (CUDA kernels + C ABI) in-tree.

nvcc cross-compiles for sm_100a without a GPU.  Called by __graft_entry__.build(); also runnable as
``python -m petsc_b200.build``.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GCC = "/usr/bin/gcc"
GXX = "/usr/bin/g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-ccbin", GXX, "-I", os.path.join(ROOT, "include")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(verbose=False, force=False, ptxas_info=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    cus = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs = []
    procs = []
    for cu in cus:
        obj = os.path.join(objdir, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [cu] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", cu, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((cu, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cu, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or ptxas_info:
            sys.stdout.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % cu)
    so = os.path.join(LIBDIR, "libpetscb200.so")
    if force or _newer(so, objs):
        cmd = [NVCC, "-shared", "-o", so] + objs + ["-ccbin", GXX, "-Xcompiler", "-fPIC", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv, ptxas_info="--ptxas" in sys.argv)
