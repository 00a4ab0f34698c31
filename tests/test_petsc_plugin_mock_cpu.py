"""tests/test_petsc_plugin_gpu.py once more in the build container: the reference's unmodified ex2 / bench_kspsolve and the plugin's
driver programs with -dll_append <plugin> -mat_type aijb200 -vec_type b200, the plugin's C-ABI calls bound to the host test double
(tests/mock/libb200mock.so, LD_PRELOAD).  Covers the plugin's host logic under the reference's own programs -- golden output of
ex2_1, six KSP/PC combinations against the host types, bench_kspsolve, the registered PC / KSP / PetscSF types -- without a GPU."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_petsc_plugin_gpu as G  # noqa: E402

MOCK = os.path.join(ROOT, "tests", "mock", "libb200mock.so")


@pytest.fixture(scope="module", autouse=True)
def mock_device():
    from petsc_b200 import _capi
    n = C.c_int(0)
    if _capi.lib().b200DeviceCount(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is visible: these bodies run on the real library there (tests -m gpu)")
    if not G.have():
        pytest.skip("baseline/_ref/petsc not built (needs the build container)")
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


def test_ex2_golden_through_plugin():
    G.test_ex2_golden_through_plugin()


@pytest.mark.parametrize("opts", [
    ["-m", "100", "-n", "100", "-ksp_type", "gmres", "-pc_type", "jacobi"],
    ["-m", "60", "-n", "50", "-ksp_type", "cg", "-pc_type", "ilu"],
    ["-m", "60", "-n", "50", "-ksp_type", "gmres", "-pc_type", "none", "-ksp_gmres_restart", "10"],
    ["-m", "40", "-n", "40", "-ksp_type", "bcgs", "-pc_type", "jacobi"],
    ["-m", "30", "-n", "25"],
    ["-m", "30", "-n", "25", "-ksp_type", "cg", "-pc_type", "icc"],
])
def test_ex2_plugin_matches_reference_cpu_types(opts):
    G.test_ex2_plugin_matches_reference_cpu_types(opts)


def test_bench_kspsolve_through_plugin():
    G.test_bench_kspsolve_through_plugin()


def test_ex2_fused_jacobi_pc_matches_pcjacobi():
    G.test_ex2_fused_jacobi_pc_matches_pcjacobi()


def test_ex2_bicg_uses_device_transpose():
    G.test_ex2_bicg_uses_device_transpose()


def test_ex2_pipecgb200_registered_ksp_matches_reference_pipecg():
    G.test_ex2_pipecgb200_registered_ksp_matches_reference_pipecg()
