"""ICC(0) on the device through the C ABI: factor layout index-exact, factor values and PCApply bit-identical to the
reference (MatCholeskyFactorNumeric_SeqAIJ with its linked-list merge order; MatSolve_SeqSBAIJ_1_NaturalOrdering) -- against the
oracle restatement and against fixtures the reference itself produced (ref_iccsolve in tests/golden/ops_*.npz)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
from conftest import golden_path

pytestmark = pytest.mark.gpu
EPS100 = 100 * 2.220446049250313e-16


@pytest.fixture(scope="module")
def H():
    from petsc_b200 import _capi
    h = _capi.Handle()
    yield h
    h.close()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def icc_plan(H, ai, aj, aa):
    from petsc_b200 import _capi
    L = _capi.lib()
    n, nnz = len(ai) - 1, len(aj)
    ai = np.ascontiguousarray(ai, np.int32); aj = np.ascontiguousarray(aj, np.int32)
    plan = C.c_void_p()
    _capi.check(L.b200Icc0Symbolic(H.h, n, _ptr(ai), _ptr(aj), C.byref(plan)))
    d_a = _capi.DeviceArray(H, nnz, np.float64).upload(np.ascontiguousarray(aa))
    bad = C.c_int(-1)
    _capi.check(L.b200Icc0Numeric(H.h, plan, d_a.ptr, C.c_double(EPS100), C.byref(bad)))
    return plan, d_a, bad.value


def matrices(oracle):
    return [("lap5", oracle.lap5(17, 13)), ("lap7", oracle.lap7(9, 7, 8)), ("lap27", oracle.lap27(9)), ("lap5_line", oracle.lap5(1, 40)), ("lap7_big", oracle.lap7(40, 33, 21))]


def test_icc0_factor_bit_exact_vs_oracle(H, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    for name, (ai, aj, aa) in matrices(oracle):
        n = len(ai) - 1
        plan, d_a, bad = icc_plan(H, ai, aj, aa)
        assert bad == 0, name
        oui, ouj, oud, oua = oracle.icc0(ai, aj, aa)
        nzu = int(oui[n])
        ui = np.zeros(n + 1, np.int32); uj = np.zeros(nzu, np.int32); ud = np.zeros(n, np.int32); ua = np.zeros(nzu)
        _capi.check(L.b200Icc0GetFactor(H.h, plan, _ptr(ui), _ptr(uj), _ptr(ud), _ptr(ua)))
        assert np.array_equal(ui, oui) and np.array_equal(uj, ouj[:nzu]) and np.array_equal(ud, oud[:n]), name   # layout: index-exact
        assert np.array_equal(ua, oua[:nzu]), (name, float(np.abs(ua - oua[:nzu]).max()))                          # factor: bit-identical
        # solve: bit-identical to MatSolve_SeqSBAIJ_1_NaturalOrdering's operation order
        rng = np.random.default_rng(3)
        for _ in range(2):
            b = rng.uniform(-1, 1, n)
            d_b = _capi.DeviceArray(H, n, np.float64).upload(b); d_x = _capi.DeviceArray(H, n, np.float64)
            _capi.check(L.b200Icc0Solve(H.h, plan, d_b.ptr, d_x.ptr))
            assert np.array_equal(d_x.download(), oracle.matsolve_icc(oui, ouj, oud, oua, b)), name
            d_b.free(); d_x.free()
        nz = C.c_int64(); lv = [C.c_int(), C.c_int(), C.c_int()]
        _capi.check(L.b200Icc0GetInfo(plan, C.byref(nz), C.byref(lv[0]), C.byref(lv[1]), C.byref(lv[2])))
        assert nz.value == nzu and lv[0].value >= 1
        _capi.check(L.b200Icc0Destroy(plan)); d_a.free()


@pytest.mark.parametrize("path", sorted(glob.glob(golden_path("ops_lap*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_icc0_pcapply_bit_exact_vs_reference_fixture(H, oracle, path):
    """ref_iccsolve = PCApply(PCICC) of the reference itself on x (oracle/gen_golden.py -icc)."""
    from petsc_b200 import _capi
    L = _capi.lib()
    g = np.load(path)
    if "ref_iccsolve" not in g.files:
        pytest.skip("fixture has no ICC record")
    ai, aj, aa = getattr(oracle, str(g["gen"]))(*[int(v) for v in g["args"]])
    n = len(ai) - 1
    plan, d_a, bad = icc_plan(H, ai, aj, aa)
    assert bad == 0
    d_b = _capi.DeviceArray(H, n, np.float64).upload(g["x"]); d_x = _capi.DeviceArray(H, n, np.float64)
    _capi.check(L.b200Icc0Solve(H.h, plan, d_b.ptr, d_x.ptr))
    assert np.array_equal(d_x.download(), g["ref_iccsolve"])
    _capi.check(L.b200Icc0Destroy(plan))
    for o in (d_a, d_b, d_x):
        o.free()


def test_icc0_indefinite_pivot_is_reported(H, oracle):
    """MatPivotCheck_pd territory: a non-positive pivot is reported (row + 1), the plan refuses to solve."""
    from petsc_b200 import _capi
    L = _capi.lib()
    ai = np.array([0, 2, 4], np.int32); aj = np.array([0, 1, 0, 1], np.int32); aa = np.array([1.0, 2.0, 2.0, 1.0])   # 1 - 4 < 0
    plan, d_a, bad = icc_plan(H, ai, aj, aa)
    assert bad == 2
    d_b = _capi.DeviceArray(H, 2, np.float64).upload(np.ones(2)); d_x = _capi.DeviceArray(H, 2, np.float64)
    assert L.b200Icc0Solve(H.h, plan, d_b.ptr, d_x.ptr) == 58   # PETSC_ERR_ORDER
    _capi.check(L.b200Icc0Destroy(plan))
    for o in (d_a, d_b, d_x):
        o.free()
    with pytest.raises(_capi.B200Error):   # missing diagonal (aijfact.c:2071)
        icc_plan(H, np.array([0, 1, 3], np.int32), np.array([1, 0, 1], np.int32), np.ones(3))
