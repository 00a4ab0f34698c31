"""GPU parity tests of the widening rows: device COO assembly and the transposed product,
through the C ABI and through the host mirror's MatSetPreallocationCOO / MatSetValuesCOO / MatMultTranspose."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import golden_path

pytestmark = pytest.mark.gpu

COO = sorted(glob.glob(golden_path("coo_*.npz")))
OPS = sorted(glob.glob(golden_path("ops_*.npz")))


@pytest.fixture(scope="module")
def H():
    from petsc_b200 import _capi
    h = _capi.Handle(0)
    yield h
    h.close()


@pytest.fixture(scope="module")
def P():
    from harness import petsc
    petsc.initialize()
    return petsc


def coo_plan(H, M, N, ci, cj):
    from petsc_b200 import _capi
    L = _capi.lib()
    d_i, d_j = H.array(ci, np.int32), H.array(cj, np.int32)
    plan = C.c_void_p()
    rc = L.b200CooPlanCreate(H.h, M, N, C.c_int64(len(ci)), d_i.ptr, d_j.ptr, C.byref(plan))
    return rc, plan


def plan_csr(H, plan, M):
    from petsc_b200 import _capi
    L = _capi.lib()
    nnz, atot, rp, cx = C.c_int64(), C.c_int64(), C.c_void_p(), C.c_void_p()
    _capi.check(L.b200CooPlanGetCsr(plan, C.byref(nnz), C.byref(atot), C.byref(rp), C.byref(cx)))
    ai = np.empty(M + 1, np.int32); aj = np.empty(max(nnz.value, 1), np.int32)
    _capi.check(L.b200MemcpyDtoH(H.h, ai.ctypes.data_as(C.c_void_p), rp, C.c_size_t(4 * (M + 1))))
    if nnz.value:
        _capi.check(L.b200MemcpyDtoH(H.h, aj.ctypes.data_as(C.c_void_p), cx, C.c_size_t(4 * nnz.value)))
    jmap = np.empty(nnz.value + 1, np.int32); perm = np.empty(max(atot.value, 1), np.int32)
    _capi.check(L.b200CooPlanGetMaps(H.h, plan, jmap.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p)))
    return ai, aj[:nnz.value], jmap, perm[:atot.value]


@pytest.mark.parametrize("path", COO, ids=[os.path.basename(p)[:-4] for p in COO])
def test_coo_plan_pattern_index_exact_and_values(H, oracle, path):
    """Pattern and repeat counts are index-exact against the reference fixture; the value pass is bit-exact for the plan's
    own (stable) order, bit-exact against the REFERENCE values when the reference's maps are adopted, and within
    rounding of the reference for the device-sorted order (identical when no pair repeats more than twice)."""
    from petsc_b200 import _capi
    L = _capi.lib()
    g = np.load(path)
    M, N, ci, cj, v1, v2 = int(g["M"]), int(g["N"]), g["coo_i"], g["coo_j"], g["v1"], g["v2"]
    rc, plan = coo_plan(H, M, N, ci, cj)
    assert rc == 0
    ai, aj, jmap, perm = plan_csr(H, plan, M)
    assert np.array_equal(ai, g["ref_ai"]) and np.array_equal(aj, g["ref_aj"])
    rAi, rAj, rjmap, rperm = oracle.coo_prealloc(M, N, ci, cj)
    assert np.array_equal(jmap, rjmap)
    # same multiset of user entries behind every nonzero; the device order is the order of the user's array
    for q in range(len(aj)):
        seg = perm[jmap[q]:jmap[q + 1]]
        assert np.array_equal(seg, np.sort(rperm[rjmap[q]:rjmap[q + 1]]))
    d_v1, d_v2, d_a = H.array(v1), H.array(v2), H.zeros(len(aj) + 1)
    _capi.check(L.b200CooSetValues(H.h, plan, d_v1.ptr, 1, d_a.ptr))
    a1 = d_a.download()[:len(aj)]
    assert np.array_equal(a1, oracle.coo_setvalues(jmap, perm, v1))
    _capi.check(L.b200CooSetValues(H.h, plan, d_v2.ptr, 0, d_a.ptr))
    a2 = d_a.download()[:len(aj)]
    assert np.array_equal(a2, oracle.coo_setvalues(jmap, perm, v2, Aa=a1))
    reps = np.diff(jmap)
    few = reps <= 2
    assert np.array_equal(a1[few], g["ref_aa1"][few])
    assert np.allclose(a1, g["ref_aa1"], rtol=0, atol=4e-16 * reps.max())
    assert np.allclose(a2, g["ref_aa2"], rtol=0, atol=8e-16 * reps.max())
    L.b200CooPlanDestroy(plan)
    # the reference's own maps -> the reference's values, bit for bit
    plan2 = C.c_void_p()
    jm64, pm64 = np.ascontiguousarray(rjmap, np.int64), np.ascontiguousarray(rperm, np.int64)
    _capi.check(L.b200CooPlanCreateFromMaps(H.h, C.c_int64(len(rAj)), C.c_int64(len(rperm)), jm64.ctypes.data_as(C.c_void_p), pm64.ctypes.data_as(C.c_void_p), C.byref(plan2)))
    d_b = H.zeros(len(aj) + 1)
    _capi.check(L.b200CooSetValues(H.h, plan2, d_v1.ptr, 1, d_b.ptr))
    assert np.array_equal(d_b.download()[:len(aj)], g["ref_aa1"])
    _capi.check(L.b200CooSetValues(H.h, plan2, d_v2.ptr, 0, d_b.ptr))
    assert np.array_equal(d_b.download()[:len(aj)], g["ref_aa2"])
    L.b200CooPlanDestroy(plan2)


def test_coo_edge_cases(H, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    # out-of-range row / column: the reference's PETSC_ERR_ARG_OUTOFRANGE (aij.c:4561, 4631)
    rc, _ = coo_plan(H, 4, 4, [0, 4], [0, 0])
    assert rc == 63
    rc, _ = coo_plan(H, 4, 4, [0, 1], [0, 4])
    assert rc == 63
    # empty input, all-negative input
    for ci, cj in (([], []), ([-1, -1, 2], [0, 1, -3])):
        rc, plan = coo_plan(H, 5, 5, np.array(ci, np.int32), np.array(cj, np.int32))
        assert rc == 0
        ai, aj, jmap, perm = plan_csr(H, plan, 5)
        assert np.array_equal(ai, np.zeros(6, np.int32)) and len(aj) == 0 and len(perm) == 0 and jmap[0] == 0
        _capi.check(L.b200CooSetValues(H.h, plan, None, 1, None))
        L.b200CooPlanDestroy(plan)
    # a larger random case against the oracle (many rows empty, heavy repeats on a few entries)
    rng = np.random.default_rng(5)
    M, N, n = 3000, 2500, 200000
    ci = rng.integers(-1, M, n).astype(np.int32); cj = rng.integers(-1, N, n).astype(np.int32)
    ci[:5000] = 17; cj[:5000] = rng.integers(0, 3, 5000)
    rc, plan = coo_plan(H, M, N, ci, cj)
    assert rc == 0
    ai, aj, jmap, perm = plan_csr(H, plan, M)
    rAi, rAj, rjmap, rperm = oracle.coo_prealloc(M, N, ci, cj)
    assert np.array_equal(ai, rAi) and np.array_equal(aj, rAj) and np.array_equal(jmap, rjmap)
    ok = np.flatnonzero((ci >= 0) & (cj >= 0))
    assert np.array_equal(np.sort(perm), ok)
    # stable: inside every nonzero the user's positions are increasing
    assert all(np.all(np.diff(perm[jmap[q]:jmap[q + 1]]) > 0) for q in np.flatnonzero(np.diff(jmap) > 1)[:2000])
    v = rng.uniform(-1, 1, n)
    d_v, d_a = H.array(v), H.zeros(len(aj) + 1)
    _capi.check(L.b200CooSetValues(H.h, plan, d_v.ptr, 1, d_a.ptr))
    assert np.array_equal(d_a.download()[:len(aj)], oracle.coo_setvalues(jmap, perm, v))
    L.b200CooPlanDestroy(plan)


def parity_mode(L, _capi, T):
    """one lane per row on the transposed plan: strict increasing-row accumulation, FMA-free (bit-exact mode)"""
    tp = C.c_void_p()
    _capi.check(L.b200CsrTransposeGetPlan(T, C.byref(tp)))
    _capi.check(L.b200CsrPlanSetLayout(tp, 1, 0, 0, 0))


def load_matrix(O, g):
    gen = str(g["gen"])
    if gen == "stored":
        return g["ai"], g["aj"], g["aa"]
    return getattr(O, gen)(*[int(a) for a in g["args"]])


@pytest.mark.parametrize("path", OPS, ids=[os.path.basename(p)[:-4] for p in OPS])
def test_transpose_bit_exact_vs_reference(H, oracle, path):
    from petsc_b200 import _capi
    L = _capi.lib()
    g = np.load(path)
    ai, aj, aa = load_matrix(oracle, g)
    m = len(ai) - 1
    d_ai, d_aj, d_aa = H.array(ai, np.int32), H.array(aj, np.int32), H.array(aa)
    T = C.c_void_p()
    _capi.check(L.b200CsrTransposeCreate(H.h, m, m, C.c_int64(len(aj)), d_ai.ptr, d_aj.ptr, C.byref(T)))
    _capi.check(L.b200CsrTransposeSetValues(H.h, T, d_aa.ptr))
    d_x, d_z, d_y = H.array(g["x"]), H.array(g["y"]), H.empty(m)
    # automatic layout (several lanes on long rows): north_star tolerance
    _capi.check(L.b200CsrTransposeSpMV(H.h, T, d_x.ptr, None, d_y.ptr))
    assert np.allclose(d_y.download(), g["ref_multtr"], rtol=1e-12, atol=1e-13)
    parity_mode(L, _capi, T)
    _capi.check(L.b200CsrTransposeSpMV(H.h, T, d_x.ptr, None, d_y.ptr))
    assert np.array_equal(d_y.download(), g["ref_multtr"])
    _capi.check(L.b200CsrTransposeSpMV(H.h, T, d_x.ptr, d_z.ptr, d_y.ptr))
    assert np.array_equal(d_y.download(), g["ref_multtradd"])
    # the transposed pattern is the CSR of A^T (index-exact against scipy-free construction)
    tptr, trow, tperm = np.empty(m + 1, np.int32), np.empty(len(aj), np.int32), np.empty(len(aj), np.int32)
    _capi.check(L.b200CsrTransposeGet(H.h, T, tptr.ctypes.data_as(C.c_void_p), trow.ctypes.data_as(C.c_void_p), tperm.ctypes.data_as(C.c_void_p)))
    rows = np.repeat(np.arange(m), np.diff(ai))
    order = np.lexsort((rows, aj))
    assert np.array_equal(tperm, order.astype(np.int32)) and np.array_equal(trow, rows[order])
    assert np.array_equal(tptr, np.concatenate([[0], np.cumsum(np.bincount(aj, minlength=m))]))
    L.b200CsrTransposeDestroy(T)


def test_transpose_rectangular_and_ragged(H, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(3)
    for m, n, dens in ((700, 1900, 9), (2500, 300, 40), (64, 64, 0), (1, 5, 3)):
        lens = rng.integers(0, dens + 1, m)
        lens[m // 2] = min(n, 4 * dens)
        ai = np.zeros(m + 1, np.int32); ai[1:] = np.cumsum(lens)
        aj = (np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]) if ai[-1] else np.zeros(0)).astype(np.int32)
        aa = rng.uniform(-1, 1, ai[-1])
        x, z = rng.uniform(-1, 1, m), rng.uniform(-1, 1, n)
        d_ai, d_aj, d_aa = H.array(ai, np.int32), H.array(aj if len(aj) else np.zeros(1, np.int32), np.int32), H.array(aa if len(aa) else np.zeros(1))
        T = C.c_void_p()
        _capi.check(L.b200CsrTransposeCreate(H.h, m, n, C.c_int64(len(aj)), d_ai.ptr, d_aj.ptr, C.byref(T)))
        _capi.check(L.b200CsrTransposeSetValues(H.h, T, d_aa.ptr))
        parity_mode(L, _capi, T)
        d_x, d_z, d_y = H.array(x), H.array(z), H.empty(n)
        _capi.check(L.b200CsrTransposeSpMV(H.h, T, d_x.ptr, None, d_y.ptr))
        assert np.array_equal(d_y.download(), oracle.matmulttranspose(ai, aj, aa, x, n=n)), (m, n)
        _capi.check(L.b200CsrTransposeSpMV(H.h, T, d_x.ptr, d_z.ptr, d_z.ptr))  # in place
        assert np.array_equal(d_z.download(), oracle.matmulttranspose(ai, aj, aa, x, n=n, z=z)), (m, n)
        L.b200CsrTransposeDestroy(T)


def test_host_mirror_coo_and_transpose(P, oracle):
    """MatSetPreallocationCOO / MatSetValuesCOO / MatMultTranspose of the host mirror (PETSc names and argument meaning):
    a Q1 finite-element style assembly, then a solve, then the transposed product after the values changed."""
    petsc = P
    g = np.load(golden_path("coo_fem_q1_7x6.npz"))
    M = int(g["M"])
    A = petsc.Mat.create(m=M, n=M, M=M, N=M, comm=petsc.COMM_SELF, mtype="seqaijb200")
    A.set_spmv_layout(lanes=1)  # bit-exact products
    A.set_preallocation_coo(g["coo_i"], g["coo_j"])
    A.set_values_coo(g["v1"])
    ai, aj, aa = A.csr_host()
    assert np.array_equal(ai, g["ref_ai"]) and np.array_equal(aj, g["ref_aj"])
    assert np.allclose(aa, g["ref_aa1"], rtol=0, atol=2e-15)
    A.set_values_coo(g["v2"], add=True)
    ai, aj, aa2 = A.csr_host()
    assert np.allclose(aa2, g["ref_aa2"], rtol=0, atol=4e-15)
    x, y = A.create_vecs()
    rng = np.random.default_rng(2)
    xv = rng.uniform(-1, 1, M)
    x.set_array(xv)
    A.mult(x, y)
    assert np.array_equal(y.array(), oracle.matmult(ai, aj, aa2, xv))
    A.mult_transpose(x, y)
    assert np.array_equal(y.array(), oracle.matmulttranspose(ai, aj, aa2, xv))
    # values change -> the transposed copy is refreshed
    A.set_values_coo(g["v1"])
    _, _, aa3 = A.csr_host()
    A.mult_transpose(x, y)
    assert np.array_equal(y.array(), oracle.matmulttranspose(ai, aj, aa3, xv))
    z = y.duplicate()
    z.set_array(xv[::-1].copy())
    A.mult_transpose_add(x, z, z)
    assert np.array_equal(z.array(), oracle.matmulttranspose(ai, aj, aa3, xv, z=xv[::-1].copy()))
    # device-resident index and value arrays are accepted as they are (PetscGetMemType dispatch)
    from petsc_b200 import _capi
    Hh = type("Hh", (), {"h": petsc.handle()})
    d_i, d_j, d_v = _capi.DeviceArray(Hh, len(g["coo_i"]), np.int32), _capi.DeviceArray(Hh, len(g["coo_i"]), np.int32), _capi.DeviceArray(Hh, len(g["coo_i"]), np.float64)
    d_i.upload(g["coo_i"]); d_j.upload(g["coo_j"]); d_v.upload(g["v1"])
    B = petsc.Mat.create(m=M, n=M, M=M, N=M, comm=petsc.COMM_SELF, mtype="seqaijb200")
    B.set_preallocation_coo(d_i.ptr, d_j.ptr, n=len(g["coo_i"]))
    B.set_values_coo(d_v.ptr)
    bi, bj, ba = B.csr_host()
    assert np.array_equal(bi, ai) and np.array_equal(bj, aj) and np.array_equal(ba, aa3)
    for o in (x, y, z):
        o.destroy()
    A.destroy(); B.destroy()
