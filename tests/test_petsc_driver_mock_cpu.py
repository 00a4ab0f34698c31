"""tests/test_petsc_driver_gpu.py once more in the build container, with the driver process bound to the host test double of the C ABI
(tests/mock/libb200mock.so, LD_PRELOAD): the reference's own KSPSolve + the plugin's types, checked against the oracle and the
reference-run fixtures.  What this covers without a GPU is the plugin's logic on the measured path -- mpiaijb200 on one rank
(100+ oracle checks), PCIe byte accounting of a device-resident solve (the mock counts every b200Memcpy), VecGetLocalVector
aliasing under PCBJACOBI, the fused PCJACOBI sub-class against the stock one, pipecgb200, ILU(0)/ICC(0) through MatSolverType b200,
residual histories at 1e-12 * r0 -- not the CUDA kernels (the mock's are sequential loops; ILU/ICC are the oracle's)."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_petsc_driver_gpu as G  # noqa: E402

MOCK = os.path.join(ROOT, "tests", "mock", "libb200mock.so")


@pytest.fixture(scope="module", autouse=True)
def mock_device():
    from petsc_b200 import _capi
    n = C.c_int(0)
    if _capi.lib().b200DeviceCount(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is visible: these bodies run on the real library there (tests -m gpu)")
    if not G.have():
        pytest.skip("baseline/_ref/petsc or petsc_plugin/b200_driver not built (needs the build container)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200mock_build", os.path.join(ROOT, "tests", "mock", "build.py"))
    mb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mb)
    mb.build()
    old = dict(G._ENV)
    G._ENV.clear()
    G._ENV["LD_PRELOAD"] = MOCK
    yield
    G._ENV.clear()
    G._ENV.update(old)


def test_mpiaijb200_one_rank_parity(oracle):
    G.test_mpiaijb200_one_rank_parity.__wrapped__(oracle) if hasattr(G.test_mpiaijb200_one_rank_parity, "__wrapped__") else G.test_mpiaijb200_one_rank_parity(oracle)


def test_device_resident_solve_moves_no_vectors_over_pcie():
    G.test_device_resident_solve_moves_no_vectors_over_pcie()


def test_bjacobi_ilu_stays_on_device():
    G.test_bjacobi_ilu_stays_on_device()


def test_fused_pcjacobi_subclass_equals_stock_pcjacobi():
    G.test_fused_pcjacobi_subclass_equals_stock_pcjacobi()


def test_e2e_host_buffers_and_ex2_config1():
    G.test_e2e_host_buffers_and_ex2_config1()


@pytest.mark.parametrize("fixture", ["ksp_lap27_10_cg_icc", "ksp_lap5_30_cg_icc", "ksp_lap7_12_gmres_icc", "ksp_lap27_10_cg_ilu", "ksp_ex2_100_gmres_jacobi",
                                     "ksp_lap5_30_pipecg_jacobi", "ksp_lap27_10_pipecg_icc"])
def test_real_petsc_ksp_history_vs_reference_fixture(oracle, fixture):
    G.test_real_petsc_ksp_history_vs_reference_fixture(oracle, fixture)


@pytest.mark.parametrize("fixture", ["ksp_lap5_30_pgmres_jacobi", "ksp_lap7_12_pgmres_ilu", "ksp_lap5_30_pgmres5_none"])
def test_pgmresb200_history_vs_reference_pgmres_fixture(oracle, fixture):
    """The pipelined GMRES registered by the plugin (reductions launched asynchronously into mapped memory, read after an event one
    iteration later) against the reference's KSPPGMRES fixtures: 1e-12 * r0, same iteration count and reason."""
    G.test_real_petsc_ksp_history_vs_reference_fixture(oracle, fixture)
