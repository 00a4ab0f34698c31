"""Door to petsc_plugin/b200_driver: the PETSc PROGRAM (the reference's own libpetsc + libpetscb200plugin.so) that runs the
BASELINE workloads on the b200 types.  bench.py, tools/ and the GPU tests call it through here.

Two ways to run it, same arguments, same JSON records back:
  * in process  (default): ctypes loads petsc_plugin/libb200driver.so -- and with it libpetsc.so, libpetscb200plugin.so and
    libpetscb200.so -- into THIS Python process and calls b200_driver_main(argc, argv);
  * subprocess: the stand-alone executable petsc_plugin/b200_driver (what `ncu` profiles).
The NCCL ranks (one process per GPU) are joined through the environment the plugin reads (petscb200_plugin.c PB_Init):
PETSCB200_NRANKS / PETSCB200_RANK / PETSCB200_NCCL_ID / PETSCB200_DEVICE.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

from . import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN_DIR = os.path.join(ROOT, "petsc_plugin")
DRIVER_EXE = os.path.join(PLUGIN_DIR, "b200_driver")
DRIVER_SO = os.path.join(PLUGIN_DIR, "libb200driver.so")
PLUGIN_SO = os.path.join(PLUGIN_DIR, "libpetscb200plugin.so")
PETSC_SO = os.path.join(ROOT, "baseline", "_ref", "petsc", "lib", "libpetsc.so")  # the host application: the unmodified reference
BLASDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"

_drv = None


def available():
    """The PETSc library is built from /root/reference in the build container (__graft_entry__.build()) into baseline/_ref and
    travels with the repository snapshot; without it there is no PETSc to host the plugin."""
    return all(os.path.exists(p) for p in (PETSC_SO, PLUGIN_SO, DRIVER_SO, DRIVER_EXE))


def unique_id_hex():
    """ncclUniqueId for PETSCB200_NCCL_ID (rank 0 calls this, the launcher hands it to every process)."""
    buf = C.create_string_buffer(128)
    _capi.check(_capi.lib().b200CommGetUniqueId(buf))
    return buf.raw.hex()


def rank_env(rank=0, size=1, uid_hex=None, device=None):
    env = {}
    if device is not None:
        env["PETSCB200_DEVICE"] = str(device)
    if size > 1:
        env.update(PETSCB200_NRANKS=str(size), PETSCB200_RANK=str(rank), PETSCB200_NCCL_ID=uid_hex)
    return env


def _load():
    global _drv
    if _drv is None:
        if not available():
            raise ImportError("petsc_plugin/libb200driver.so or the PETSc library under baseline/_ref is missing: run __graft_entry__.build() "
                              "in the build container (the reference build needs /root/reference)")
        _capi.lib()
        for dep in ("libquadmath-2284e583.so.0.0.0", "libgfortran-83c28eba.so.5.0.0", "libopenblasp-r0-59ffcd50.3.15.so"):
            p = os.path.join(BLASDIR, dep)
            if os.path.exists(p):
                C.CDLL(p, mode=C.RTLD_GLOBAL)   # libpetsc's BLAS/LAPACK (no RPATH of ours reaches them)
        C.CDLL(PETSC_SO, mode=C.RTLD_GLOBAL)
        C.CDLL(PLUGIN_SO, mode=C.RTLD_GLOBAL)
        _drv = C.CDLL(DRIVER_SO, mode=C.RTLD_GLOBAL)
        _drv.b200_driver_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        _drv.b200_driver_main.restype = C.c_int
    return _drv


def run(args, env=None, inproc=True, timeout=3600):
    """Runs the driver with `args` (list of strings) and returns the list of JSON records it emitted (rank 0's; other ranks
    return []).  Raises on a non-zero exit."""
    env = dict(env or {})
    with tempfile.NamedTemporaryFile(prefix="b200drv_", suffix=".jsonl", delete=False) as f:
        out = f.name
    argv = ["b200_driver"] + [str(a) for a in args] + ["-json_out", out, "-no_signal_handler", "-options_left", "0"]
    try:
        if inproc:
            drv = _load()
            old = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            try:
                arr = (C.c_char_p * (len(argv) + 1))(*[a.encode() for a in argv], None)
                rc = drv.b200_driver_main(len(argv), arr)
            finally:
                for k, v in old.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            if rc != 0:
                raise RuntimeError("b200_driver_main(%s) returned %d (PETSc error: see stderr)" % (" ".join(argv[1:]), rc))
        else:
            e = dict(os.environ, **env)
            e["LD_LIBRARY_PATH"] = BLASDIR + ":" + e.get("LD_LIBRARY_PATH", "")
            p = subprocess.run([DRIVER_EXE] + argv[1:], env=e, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
            if p.returncode != 0:
                raise RuntimeError("b200_driver %s failed (%d)\n%s\n%s" % (" ".join(argv[1:]), p.returncode, p.stdout[-3000:], p.stderr[-3000:]))
        recs = []
        if os.path.exists(out):
            for line in open(out):
                line = line.strip()
                if line:
                    recs.append(json.loads(line))
        return recs
    finally:
        try:
            os.unlink(out)
        except OSError:
            pass
