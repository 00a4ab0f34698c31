"""The drop-in, for real: the REFERENCE's own unmodified tutorial programs (ex2.c, bench_kspsolve.c compiled from
/root/reference against the reference library, see oracle/build_ref_demo.sh) load petsc_plugin/libpetscb200plugin.so with
-dll_append and run with -mat_type aijb200 -vec_type b200.  Their output is compared with the reference's golden file and
with the same executable on the reference's CPU types (-mat_type aij -vec_type standard) on the same box."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baseline", "_ref", "petsc", "bin")
PLUGIN = os.path.join(ROOT, "petsc_plugin", "libpetscb200plugin.so")
BLASDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
B200 = ["-dll_append", PLUGIN, "-mat_type", "aijb200", "-vec_type", "b200"]

EX2_1_OUT = """  0 KSP Residual norm 3.21109
  1 KSP Residual norm 0.93268
  2 KSP Residual norm 0.103515
  3 KSP Residual norm 0.00787798
  4 KSP Residual norm 0.000387275
Norm of error 0.000392701 iterations 4
"""


def have():
    return os.path.exists(os.path.join(BIN, "ex2")) and os.path.exists(PLUGIN)


_ENV = {}   # extra environment of the PETSc program (tests/test_petsc_plugin_mock_cpu.py re-runs these bodies with the mock device preloaded)


def run(exe, args, timeout=600):
    env = dict(os.environ, LD_LIBRARY_PATH=BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), **_ENV)
    p = subprocess.run([os.path.join(BIN, exe)] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


def history(out):
    return np.array([float(m.group(1)) for m in re.finditer(r"KSP Residual norm ([0-9.eE+-]+)", out)])


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
def test_ex2_golden_through_plugin():
    """ex2_1.out reproduced (all printed digits) by the reference's own ex2 running on the b200 types."""
    out = run("ex2", ["-m", "5", "-n", "5", "-ksp_monitor", "-ksp_gmres_cgs_refinement_type", "refine_always"] + B200)
    # the harness diffs with 6 significant digits (petscdiff / -petsc_ci); the raw monitor prints 12
    got = ["%.6g" % v for v in history(out)]
    want = ["%.6g" % v for v in history(EX2_1_OUT)]
    assert got == want and len(got) == 5
    assert out.strip().splitlines()[-1] == EX2_1_OUT.strip().splitlines()[-1]  # "Norm of error 0.000392701 iterations 4"
    # (ex2 marks A symmetric, so PETSc's default PC there is ICC on the host; ask for ILU to see the device factorisation)
    view = run("ex2", ["-m", "5", "-n", "5", "-ksp_view", "-pc_type", "ilu"] + B200)
    open(os.path.join(ROOT, "gpurun_out", "ex2_ksp_view_b200.txt"), "w").write(view) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
    assert "seqaijb200" in view and re.search(r"package used to perform factorization: *b200", view), view


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
@pytest.mark.parametrize("opts", [
    ["-m", "100", "-n", "100", "-ksp_type", "gmres", "-pc_type", "jacobi"],          # BASELINE config 1
    ["-m", "60", "-n", "50", "-ksp_type", "cg", "-pc_type", "ilu"],
    ["-m", "60", "-n", "50", "-ksp_type", "gmres", "-pc_type", "none", "-ksp_gmres_restart", "10"],
    ["-m", "40", "-n", "40", "-ksp_type", "bcgs", "-pc_type", "jacobi"],               # a KSP the plugin never heard of
    ["-m", "30", "-n", "25"],                                                          # ex2's defaults: GMRES + PCICC (A is flagged symmetric) -> ICC(0) on the device
    ["-m", "30", "-n", "25", "-ksp_type", "cg", "-pc_type", "icc"],
])
def test_ex2_plugin_matches_reference_cpu_types(opts):
    a = run("ex2", opts + ["-ksp_monitor"] + B200)
    b = run("ex2", opts + ["-ksp_monitor"])
    ha, hb = history(a), history(b)
    assert abs(len(ha) - len(hb)) <= 1
    k = min(len(ha), len(hb), 31)
    # GMRES/CG: 1e-10 over the first cycle; BiCGStab's two-term recurrences amplify summation-order differences faster
    assert np.allclose(ha[:k], hb[:k], rtol=1e-10 if "bcgs" not in opts else 1e-6, atol=1e-12 * hb[0])
    m = min(len(ha), len(hb))
    assert np.allclose(ha[:m], hb[:m], rtol=1e-4, atol=1e-11 * hb[0])
    ea = float(re.search(r"Norm of error ([0-9.eE+-]+)", a).group(1)); eb = float(re.search(r"Norm of error ([0-9.eE+-]+)", b).group(1))
    assert np.isclose(ea, eb, rtol=1e-3)


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
def test_bench_kspsolve_through_plugin():
    """The reference's own benchmark driver (27-point stencil, COO assembly path falls back to the parent) on the b200 types."""
    common = ["-n", "24", "-ksp_monitor", "-print_timing", "false", "-ksp_type", "cg", "-pc_type", "ilu"]
    a = run("bench_kspsolve", common + B200)
    b = run("bench_kspsolve", common)
    ha, hb = history(a), history(b)
    assert len(ha) == len(hb) and np.allclose(ha, hb, rtol=1e-8)
    mm = run("bench_kspsolve", ["-n", "32", "-matmult", "-its", "5", "-print_timing", "false"] + B200)
    assert "Number of nonzeros = %d" % ((3 * 32 - 2) ** 3) in mm


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
def test_ex2_fused_jacobi_pc_matches_pcjacobi():
    """-pc_type jacobib200 (PCRegister'ed by the plugin; ops->applyBA = one fused SpMV+Jacobi kernel) gives the residual history
    of the reference's PCJACOBI on the same types to the last digit printed, and of the CPU types to 1e-10."""
    opts = ["-m", "100", "-n", "100", "-ksp_type", "gmres", "-ksp_monitor"]
    a = run("ex2", opts + ["-pc_type", "jacobib200"] + B200)
    b = run("ex2", opts + ["-pc_type", "jacobi"] + B200)
    c = run("ex2", opts + ["-pc_type", "jacobi"])
    ha, hb, hc = history(a), history(b), history(c)
    assert len(ha) == len(hb) and np.array_equal(ha, hb)      # same arithmetic: row sum, then one multiply
    k = min(len(ha), len(hc), 31)
    assert np.allclose(ha[:k], hc[:k], rtol=1e-10, atol=1e-12 * hc[0])
    view = run("ex2", ["-m", "8", "-n", "8", "-ksp_view", "-pc_type", "jacobib200"] + B200)
    assert "jacobib200" in view


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
def test_ex2_bicg_uses_device_transpose():
    """KSPBICG needs MatMultTranspose and PCApplyTranspose: the plugin's explicit-transpose product vs the CPU types."""
    opts = ["-m", "40", "-n", "40", "-ksp_type", "bicg", "-pc_type", "jacobi", "-ksp_monitor"]
    ha, hb = history(run("ex2", opts + B200)), history(run("ex2", opts))
    assert abs(len(ha) - len(hb)) <= 1
    k = min(len(ha), len(hb), 25)
    assert np.allclose(ha[:k], hb[:k], rtol=1e-8, atol=1e-12 * hb[0])


@pytest.mark.skipif(not (have() and os.path.exists(os.path.join(BIN, "plugin_driver"))), reason="baseline/_ref/petsc/bin/plugin_driver not built")
def test_plugin_driver_device_coo_transpose_bindtocpu():
    out = run("plugin_driver", B200 + ["-mat_b200_spmv_ordered"])  # reference-order row sums: the products are compared bit for bit
    assert "all ok" in out and "FAILED" not in out, out
    for name in ("coo_device_insert_equals_reference", "coo_device_add_equals_reference", "coo_host_values_equals_reference", "matmult_bit_exact",
                 "matmulttranspose_bit_exact", "matmulttransposeadd_bit_exact", "matmulttransposeadd_inplace_bit_exact", "matmulttranspose_after_matscale",
                 "current_memtype_is_device", "matmult_bound_to_cpu"):
        assert "ok " + name in out, out


@pytest.mark.skipif(not have(), reason="baseline/_ref/petsc not built (needs the build container)")
def test_ex2_pipecgb200_registered_ksp_matches_reference_pipecg():
    """KSPRegister("pipecgb200") (single-reduction CG with fused recurrences) inside the reference's own ex2: the residual history
    equals the reference's KSPPIPECG on its CPU types to 1e-10 over the first 30 iterations, same iteration count."""
    opts = ["-m", "40", "-n", "35", "-pc_type", "jacobi", "-ksp_monitor", "-ksp_rtol", "1e-8"]
    a = run("ex2", opts + ["-ksp_type", "pipecgb200"] + B200)
    b = run("ex2", opts + ["-ksp_type", "pipecg"])
    ha, hb = history(a), history(b)
    assert abs(len(ha) - len(hb)) <= 1
    k = min(len(ha), len(hb), 31)
    assert np.allclose(ha[:k], hb[:k], rtol=1e-10, atol=2e-12 * hb[0])
    view = run("ex2", ["-m", "8", "-n", "8", "-ksp_type", "pipecgb200", "-pc_type", "jacobi", "-ksp_view"] + B200)
    assert "pipecgb200" in view and "seqaijb200" in view


SF_CHECKS = ("vecscatter_sf_is_the_b200_subclass", "general_forward_insert", "general_forward_add", "general_forward_max", "general_reverse_add_repeated_roots",
             "general_reverse_insert_repeated_roots", "general_reverse_min_repeated_roots", "general_scatter_ran_where_expected", "mixed_device_to_host_add",
             "mixed_host_to_device_reverse_add", "mixed_scatters_were_staged", "stride_forward_insert", "stride_reverse_add", "identity_forward_insert",
             "identity_forward_add", "in_place_insert", "block_forward_add_repeated_destinations", "block_forward_insert_repeated_destinations",
             "block_reverse_add", "scatter_to_all", "petscsf_default_type_is_the_b200_subclass", "sf_operations_ran_where_expected", "sf_new_graph_replans")


@pytest.mark.skipif(not (have() and os.path.exists(os.path.join(BIN, "sf_driver"))), reason="baseline/_ref/petsc/bin/sf_driver not built")
def test_sf_driver_vecscatter_and_petscsf_on_device_vectors():
    """VecScatter / PetscSF on b200 vectors and raw device buffers inside real PETSc (petsc_plugin/sf_driver.c): every result equals
    the same operation on the reference's host vectors bit for bit; device-to-device operations ran as device kernels (counted),
    mixed and in-place ones were staged."""
    out = run("sf_driver", B200)
    assert "vec type seqb200" in out and "all ok" in out and "FAILED" not in out, out
    for name in SF_CHECKS:
        assert "ok " + name in out, (name, out)
    for op in ("replace", "sum", "max", "min", "prod"):
        assert "ok sf_bcast_%s_int_and_scalar" % op in out and "ok sf_reduce_%s_int_and_scalar" % op in out, out


@pytest.mark.skipif(not (have() and os.path.exists(os.path.join(BIN, "coherence_driver"))), reason="baseline/_ref/petsc/bin/coherence_driver not built")
def test_coherence_driver_host_device_rules():
    """One check per host/device coherence rule of the plugin (petsc_plugin/coherence_driver.c), b200 types against host types.
    Ran green on a B200 with the very last GPU seconds of round 2 (profiles/round2_last_gpu_confirmation.log)."""
    out = run("coherence_driver", B200)
    assert "mat type seqaijb200 vec type seqb200" in out and "all ok" in out and "FAILED" not in out, out
