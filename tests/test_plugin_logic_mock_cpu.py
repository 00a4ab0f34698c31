"""The PETSc plugin's HOST LOGIC on the CPU: the reference's own device-variant test programs and KSP tutorials
(tools/ref_conformance.py, 85 cases) run inside real PETSc with the plugin loaded and the b200 types selected, while the plugin's
C-ABI calls are bound (LD_PRELOAD) to tests/mock/libb200mock.so, a host test double of libpetscb200.so with malloc'ed "device"
memory and sequential kernels.  What is exercised is everything the plugin does around the kernels: offload masks, object states and
PETSc's norm cache, lazy host arrays, mirror invalidation after MatSetValues / MatZeroEntries / COO, sub-classing of PCJACOBI and
PETSCSFBASIC, VecScatter staging ...  Each case must print what the same program prints on the host types.

This found (and now guards) the bug the reference's ex9 and mat/tests/ex254 exposed on the B200: device writes that did not bump the
vector's object state left PETSc's cached norms valid, so VecSet(x,0) was skipped and VecNorm returned stale values.
Needs the reference build of the build container (skipped elsewhere).  The mock is never used by the product or on a GPU box."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
EXP = json.load(open(os.path.join(ROOT, "tests", "ref_conformance_expected.json")))


def _setup():
    import ctypes as C
    from petsc_b200 import _capi
    n = C.c_int(0)
    if _capi.lib().b200DeviceCount(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is visible: the plugin is tested on the real library there (tests -m gpu), never on the test double")
    import ref_conformance as rc
    if not os.path.exists(rc.MANIFEST) or not os.path.exists(rc.PLUGIN):
        pytest.skip("baseline/_ref/petsc/reftests.json not built (needs the build container)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200mock_build", os.path.join(ROOT, "tests", "mock", "build.py"))
    mb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mb)
    mb.build()
    return rc, json.load(open(rc.MANIFEST))


def test_reference_programs_on_b200_types_through_the_mock_device():
    rc, manifest = _setup()
    bad, gaps_now_passing = [], []
    for c in manifest:
        cid = rc.case_id(c)
        rc_h, out_h = rc.run_case(c, False)
        rc_d, out_d = rc.run_case(c, True, mock=True)
        ok, why = rc.same_output(out_h, out_d) if (rc_h == 0 and rc_d == 0) else (False, "exit codes %d / %d: %s" % (rc_h, rc_d, rc.error_summary(out_d)))
        if cid in EXP["known_gaps"]:
            if ok:
                gaps_now_passing.append(cid)
        elif not ok:
            bad.append((cid, why))
    assert not bad, bad
    assert not gaps_now_passing, ("listed as known gaps but passing: update tests/ref_conformance_expected.json", gaps_now_passing)
    covered = {rc.case_id(c) for c in manifest}
    assert set(EXP["expected_pass"]) | set(EXP["fixed_after_last_gpu_run"]) | set(EXP["added_after_last_gpu_run"]) | set(EXP["known_gaps"]) == covered


def test_plugin_drivers_on_the_mock_device():
    """petsc_plugin/coherence_driver.c (one check per host/device coherence rule the reference's programs exposed: norm cache vs
    device writes, Vec operations on a handed-out host array, sub-vectors written on the device, MatHeaderMerge, PCJacobiGetDiagonal,
    value updates between solves), petsc_plugin/sf_driver.c (VecScatter / PetscSF on b200 vectors: the device branch of the PetscSF sub-class, staging of mixed
    and in-place scatters) and petsc_plugin/plugin_driver.c (COO from device-resident arrays, transposed products, MatBindToCPU) with
    the b200 types on the mock device: every check equals the host types bit for bit."""
    import subprocess
    rc, _ = _setup()
    env = dict(os.environ, LD_LIBRARY_PATH=rc.BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""), LD_PRELOAD=rc.MOCK)
    for exe, extra, must in (("sf_driver", [], ["vec type seqb200", "general scatter: 6 operations on the device, 0 staged through the host", "ok mixed_scatters_were_staged",
                                                  "ok sf_operations_ran_where_expected", "ok sf_new_graph_replans", "all ok"]),
                             ("plugin_driver", ["-mat_b200_spmv_ordered"], ["ok coo_device_insert_equals_reference", "ok matmulttransposeadd_inplace_bit_exact", "ok matmult_bound_to_cpu", "all ok"]),
                             ("coherence_driver", [], ["mat type seqaijb200 vec type seqb200", "ok norm_after_matmult_is_not_the_cached_one", "ok vecset_zero_after_device_write",
                                                       "ok vecset_while_host_array_is_handed_out", "ok subvector_written_on_device_then_parent_read",
                                                       "ok second_solve_after_value_update_bcgs_jacobi", "ok pcjacobigetdiagonal_after_solve", "ok inplace_lu_solve",
                                                       "ok destroy_after_headermerge", "all ok"])):
        path = os.path.join(ROOT, "baseline", "_ref", "petsc", "bin", exe)
        if not os.path.exists(path):
            pytest.skip(exe + " not built")
        p = subprocess.run([path, "-dll_append", rc.PLUGIN, "-mat_type", "aijb200", "-vec_type", "b200"] + extra, capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
        assert p.returncode == 0 and "FAILED" not in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
        for m in must:
            assert m in p.stdout, (exe, m, p.stdout[-1500:])


KSPS = ("cg groppcg pipecg pipecgrr pipelcg pipeprcg pipecg2 cgne richardson chebyshev gmres fgmres lgmres dgmres pgmres pipefgmres tcqmr fcg pipefcg bcgs "
        "qmrcgs fbcgs bcgsl pipebcgs cgs tfqmr cr pipecr lsqr bicg minres symmlq lcd gcr pipegcr cgls pipecgb200 pgmresb200").split()
PCS = ["none", "jacobi", "sor", "ilu", "icc", "bjacobi", "asm", "pbjacobi", "lu", "cholesky", "gamg", "eisenstat"]


def test_every_ksp_and_pc_of_the_reference_on_b200_types_through_the_mock_device():
    """The reference's ex2 with every KSP type (x Jacobi, ILU) and every PC type (x GMRES) on the b200 types vs the host types: the
    first dozen residual norms must agree.  This walks through every Vec operation the Krylov methods use (VecMTDot, VecDotNorm2,
    VecAXPBYPCZ, VecSwap, VecNormalize, ...), device or inherited, with their offload-mask and object-state bookkeeping.
    Not in the list: ibcgs and fbcgsr keep raw VecGetArray() pointers across Vec / Mat operations (fbcgsr.c:41-60), which only a
    host vector type can honour -- they fail the same way on the reference's own device vectors."""
    import re
    import subprocess
    rc, _ = _setup()
    ex2 = os.path.join(ROOT, "baseline", "_ref", "petsc", "bin", "ex2")
    env = dict(os.environ, LD_LIBRARY_PATH=rc.BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    combos = [(k, p) for k in KSPS for p in ("jacobi", "ilu")] + [("gmres", p) for p in PCS]
    bad, compared = [], 0
    for k, p in combos:
        args = ["-m", "9", "-n", "8", "-ksp_type", k, "-pc_type", p, "-ksp_monitor", "-ksp_max_it", "40", "-dll_append", rc.PLUGIN]
        a = subprocess.run([ex2] + args, capture_output=True, text=True, errors="replace", env=env, timeout=60)
        if a.returncode != 0:
            continue   # a combination the reference itself refuses (e.g. cholesky on the non-symmetric flag)
        b = subprocess.run([ex2] + args + ["-mat_type", "aijb200", "-vec_type", "b200"], capture_output=True, text=True, errors="replace", env=dict(env, LD_PRELOAD=rc.MOCK), timeout=60)
        ha = [float(x) for x in re.findall(r"KSP Residual norm ([0-9.eE+-]+)", a.stdout)]
        hb = [float(x) for x in re.findall(r"KSP Residual norm ([0-9.eE+-]+)", b.stdout)]
        compared += 1
        if b.returncode != 0 or not ha or abs(len(ha) - len(hb)) > 1 or any(abs(x - y) > 1e-6 * ha[0] + 1e-3 * abs(x) for x, y in zip(ha[:12], hb[:12])):
            bad.append((k, p, rc.error_summary(b.stdout + b.stderr)[:200] if b.returncode else (ha[:3], hb[:3])))
    assert not bad, bad
    assert compared >= 80, compared


def test_mock_is_not_reachable_from_the_product():
    """The test double must never be what the product loads: different file name, outside the package, not referenced by the
    package, the plugin or the build scripts."""
    import subprocess
    for path in ("petsc_b200", "petsc_plugin", "include", "bench.py", "__graft_entry__.py", "oracle/build_ref.sh", "oracle/build_ref_demo.sh"):
        out = subprocess.run(["grep", "-rIl", "b200mock", os.path.join(ROOT, path)], capture_output=True, text=True).stdout.strip()
        assert not out, out
    assert not os.path.exists(os.path.join(ROOT, "petsc_b200", "lib", "libb200mock.so"))
