"""Host-side behaviour of the PETSc plugin WITHOUT a GPU (the build container): the reference's own ex2 loads
libpetscb200plugin.so with -dll_append;
  * the registrations do not disturb the reference's CPU types: -pc_type jacobi (now the plugin's sub-class of PCJACOBI) and the
    registered KSP pipecgb200 give the stock results on -mat_type aij -vec_type standard, without touching CUDA;
  * the b200 types fail loudly with PETSC_ERR_GPU -- there is no CPU fallback behind -mat_type aijb200 -vec_type b200.
Skipped where the reference library is absent or a GPU is visible (the GPU suite covers that side)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX2 = os.path.join(ROOT, "baseline", "_ref", "petsc", "bin", "ex2")
PLUGIN = os.path.join(ROOT, "petsc_plugin", "libpetscb200plugin.so")
BLASDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"


def gpu_visible():
    from petsc_b200 import _capi
    n = C.c_int(0)
    return _capi.lib().b200DeviceCount(C.byref(n)) == 0 and n.value > 0


pytestmark = pytest.mark.skipif(not (os.path.exists(EX2) and os.path.exists(PLUGIN)) or gpu_visible(), reason="needs baseline/_ref/petsc (build container) and no GPU")


def ex2(args):
    env = dict(os.environ, LD_LIBRARY_PATH=BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([EX2] + args, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def last(out):
    return [l for l in out.stdout.splitlines() if l.startswith("Norm of error")][-1]


def test_plugin_leaves_cpu_types_untouched():
    base = ["-m", "12", "-n", "11"]
    for opts in (["-pc_type", "jacobi"], ["-pc_type", "jacobi", "-ksp_type", "cg"], ["-pc_type", "jacobi", "-pc_jacobi_type", "rowmax"], ["-pc_type", "ilu"], []):
        a, b = ex2(base + opts + ["-dll_append", PLUGIN]), ex2(base + opts)
        assert a.returncode == 0 and b.returncode == 0, a.stdout + a.stderr
        assert last(a) == last(b), (opts, last(a), last(b))
    # the registered KSP runs on host vectors through the eight BLAS-1 calls: same iterates as the reference's KSPPIPECG
    a = ex2(base + ["-pc_type", "jacobi", "-ksp_type", "pipecgb200", "-dll_append", PLUGIN])
    b = ex2(base + ["-pc_type", "jacobi", "-ksp_type", "pipecg"])
    assert a.returncode == 0 and last(a) == last(b), a.stdout + a.stderr
    # the registered pipelined GMRES on host vectors (public Vec interface): the reference's KSPPGMRES history, restarts included
    for extra in ([], ["-ksp_gmres_restart", "7"]):
        a = ex2(["-m", "30", "-n", "30", "-pc_type", "jacobi", "-ksp_rtol", "1e-8", "-ksp_monitor", "-ksp_type", "pgmresb200", "-dll_append", PLUGIN] + extra)
        b = ex2(["-m", "30", "-n", "30", "-pc_type", "jacobi", "-ksp_rtol", "1e-8", "-ksp_monitor", "-ksp_type", "pgmres"] + extra)
        ha = [float(m) for m in re.findall(r"KSP Residual norm ([0-9.eE+-]+)", a.stdout)]
        hb = [float(m) for m in re.findall(r"KSP Residual norm ([0-9.eE+-]+)", b.stdout)]
        assert a.returncode == 0 and len(ha) == len(hb) and len(ha) > 20, a.stdout[-500:] + a.stderr[-500:]
        assert max(abs(x - y) for x, y in zip(ha, hb)) <= 1e-10 * hb[0]
        ea, eb = [re.search(r"Norm of error ([0-9.eE+-]+) iterations (\d+)", last(o)).groups() for o in (a, b)]
        assert ea[1] == eb[1] and abs(float(ea[0]) - float(eb[0])) <= 1e-3 * float(eb[0])
    view = ex2(base + ["-pc_type", "jacobi", "-ksp_view", "-dll_append", PLUGIN]).stdout
    assert re.search(r"type: jacobi", view) and "seqaij" in view and "b200" not in view


def test_sf_subclass_leaves_host_scatters_to_the_reference():
    """The plugin registers its PetscSF sub-class under the name "basic" as well; on host vectors every VecScatter / PetscSF operation
    of petsc_plugin/sf_driver.c must still be the reference's (no device touched, nothing staged), and -b200_keep_sfbasic must leave
    the stock type in place."""
    exe = os.path.join(ROOT, "baseline", "_ref", "petsc", "bin", "sf_driver")
    if not os.path.exists(exe):
        pytest.skip("sf_driver not built")
    env = dict(os.environ, LD_LIBRARY_PATH=BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe, "-dll_append", PLUGIN], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and "all ok" in out.stdout and "FAILED" not in out.stdout, out.stdout + out.stderr
    assert "general scatter: 0 operations on the device, 0 staged through the host" in out.stdout
    assert "ok sf_operations_ran_where_expected" in out.stdout and "ok in_place_insert" in out.stdout
    keep = subprocess.run([exe, "-dll_append", PLUGIN, "-b200_keep_sfbasic"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert "FAILED vecscatter_sf_is_the_b200_subclass" in keep.stdout      # the stock PETSCSFBASIC: no composed counter


def test_b200_types_fail_loudly_without_gpu():
    out = ex2(["-m", "5", "-n", "5", "-dll_append", PLUGIN, "-mat_type", "aijb200", "-vec_type", "b200"])
    assert out.returncode != 0
    err = out.stdout + out.stderr
    assert "GPU error" in err and "petscb200 error 97" in err       # PETSC_ERR_GPU from PB_Init: no CPU fallback
