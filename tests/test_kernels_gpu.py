"""GPU parity tests of the sm_100a kernels, called through the C ABI (include/petscb200.h), checked against the CPU
oracle and the reference fixtures.  Bit-exact where the reference arithmetic is in-tree and deterministic (MatMult,
MatMultAdd, GetDiagonal, VecMAXPY, PointwiseMult); 1e-12 relative for reductions (north_star tolerance)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import golden_path

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # BASELINE.json north_star: fp64 within 1e-12 relative


@pytest.fixture(scope="module")
def H():
    from petsc_b200 import _capi
    h = _capi.Handle(0)
    yield h
    h.close()


def upload_csr(H, ai, aj, aa):
    d_ai, d_aj, d_aa = H.array(ai, np.int32), H.array(aj, np.int32), H.array(aa, np.float64)
    plan = H.csr_plan(len(ai) - 1, len(ai) - 1, len(aj), d_ai, d_aj)
    return d_ai, d_aj, d_aa, plan


def matrices(O):
    rng = np.random.default_rng(1)
    yield "lap5_37x29", O.lap5(37, 29)
    yield "lap7_23x17x11", O.lap7(23, 17, 11)
    yield "lap27_13", O.lap27(13)
    yield "rand_5000x5", O.random_csr(5000, 5, 11)
    yield "rand_3000x32", O.random_csr(3000, 32, 12)
    yield "rand_700x128", O.random_csr(700, 128, 13)
    yield "rand_300x512", O.random_csr(300, 512, 14, ncols=4000)
    # ragged: empty rows, one very long row (exercises the in-kernel fallback tile)
    n = 4000
    lens = rng.integers(0, 9, n)
    lens[5] = 0; lens[n - 1] = 0; lens[1234] = 3900
    ai = np.zeros(n + 1, np.int32); ai[1:] = np.cumsum(lens)
    aj = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.int32)
    aa = rng.uniform(-1, 1, ai[-1])
    yield "ragged_longrow", (ai, aj, aa)
    # single row / tiny
    yield "tiny", (np.array([0, 2, 3], np.int32), np.array([0, 1, 1], np.int32), np.array([2.0, -1.0, 3.0]))


def test_spmv_bit_exact_parity_mode(H, oracle):
    """lanes_per_row = 1 reproduces MatMult_SeqAIJ / MatMultAdd_SeqAIJ bit for bit on every matrix class."""
    rng = np.random.default_rng(7)
    for name, (ai, aj, aa) in matrices(oracle):
        m = len(ai) - 1
        ncol = int(aj.max()) + 1 if len(aj) else m
        ncol = max(ncol, m)
        x = rng.uniform(-1, 1, ncol); y0 = rng.uniform(-1, 1, m)
        d_ai, d_aj, d_aa, plan = upload_csr(H, ai, aj, aa)
        H.csr_plan_set_layout(plan, lanes=1)
        d_x, d_y, d_z = H.array(x), H.empty(m), H.array(y0)
        H.spmv(plan, d_aa, d_x, d_y)
        ref = oracle.matmult(ai, aj, aa, x)
        assert np.array_equal(d_y.download(), ref), name
        H.spmv_add(plan, d_aa, d_x, d_z, d_z)  # in place z = z + A x
        assert np.array_equal(d_z.download(), oracle.matmultadd(ai, aj, aa, x, y0)), name


def test_spmv_every_layout_is_bit_exact_tree_mode_within_tolerance(H, oracle):
    """b200CsrPlanSetSummation(plan, 0): the reference's left-to-right FMA-free row sums for EVERY lanes-per-row (automatic
    layout included), so MatMult / MatMultAdd / the fused Jacobi epilogue are bit-identical to the CPU path on every matrix
    class.  The default (1) is the FMA + shuffle-tree variant for rows with several lanes: north_star tolerance."""
    from petsc_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(8)
    for name, (ai, aj, aa) in matrices(oracle):
        m = len(ai) - 1
        ncol = max(int(aj.max()) + 1 if len(aj) else m, m)
        x = rng.uniform(-1, 1, ncol); y0 = rng.uniform(-1, 1, m); dinv = rng.uniform(0.5, 2.0, m)
        d_ai, d_aj, d_aa, plan = upload_csr(H, ai, aj, aa)
        lay = H.csr_plan_layout(plan)
        d_x, d_y, d_y0, d_dinv, d_w = H.array(x), H.empty(m), H.array(y0), H.array(dinv), H.empty(m)
        ref = oracle.matmult(ai, aj, aa, x)
        refadd = oracle.matmultadd(ai, aj, aa, x, y0)
        scale = np.abs(ref).max() if m else 1.0
        _capi.check(L.b200CsrPlanSetSummation(plan, 0))
        for lanes in (0, 2, 4, 8, 16, 32):
            H.csr_plan_set_layout(plan, lanes=lanes)
            H.spmv(plan, d_aa, d_x, d_y)
            assert np.array_equal(d_y.download(), ref), (name, lanes, lay)
            H.spmv_add(plan, d_aa, d_x, d_y0, d_y)
            assert np.array_equal(d_y.download(), refadd), (name, lanes)
            H.spmv_jacobi(plan, d_aa, d_x, d_dinv, d_w, d_y)
            assert np.array_equal(d_y.download(), ref) and np.array_equal(d_w.download(), ref * dinv), (name, lanes)
        _capi.check(L.b200CsrPlanSetSummation(plan, 1))
        for lanes in (0, 2, 4, 8, 16, 32):
            H.csr_plan_set_layout(plan, lanes=lanes)
            H.spmv(plan, d_aa, d_x, d_y)
            assert np.abs(d_y.download() - ref).max() <= RTOL * max(scale, 1e-300), (name, lanes)
        for rows, stages, ctas in ((64, 3, 1), (256, 4, 2), (1024, 1, 1), (8, 2, 4)):
            H.csr_plan_set_layout(plan, lanes=1, rows=rows, stages=stages, ctas=ctas)
            H.spmv(plan, d_aa, d_x, d_y)
            assert np.array_equal(d_y.download(), ref), (name, rows, stages, ctas)


def test_spmv_column_blocks_keep_the_reference_order(H, oracle):
    """b200CsrPlanSetColumnBlocks: y = A x as nb passes over column ranges (x stays L2-resident per pass).  The passes continue
    each other's row sums in column order, so with the exact summation every result is still bit-identical to MatMult_SeqAIJ /
    MatMultAdd_SeqAIJ / the fused Jacobi form, for every lane count; the tree summation stays within tolerance."""
    from petsc_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(18)
    for name, (ai, aj, aa) in matrices(oracle):
        m = len(ai) - 1
        ncol = max(int(aj.max()) + 1 if len(aj) else m, m)
        x = rng.uniform(-1, 1, ncol); y0 = rng.uniform(-1, 1, m); dinv = rng.uniform(0.5, 2.0, m)
        d_ai, d_aj, d_aa = H.array(ai, np.int32), H.array(aj, np.int32), H.array(aa, np.float64)
        plan = H.csr_plan(m, ncol, len(aj), d_ai, d_aj)
        d_x, d_y, d_y0, d_z, d_dinv, d_w = H.array(x), H.empty(m), H.array(y0), H.empty(m), H.array(dinv), H.empty(m)
        ref, refadd = oracle.matmult(ai, aj, aa, x), oracle.matmultadd(ai, aj, aa, x, y0)
        scale = max(np.abs(ref).max() if m else 1.0, 1e-300)
        for nb in (2, 3, 7):
            _capi.check(L.b200CsrPlanSetColumnBlocks(H.h, plan, nb))
            assert L.b200CsrSpMV(H.h, plan, d_aa.ptr, d_x.ptr, d_y.ptr) == 58, name   # PETSC_ERR_ORDER: values not packed yet
            _capi.check(L.b200CsrPlanPackValues(H.h, plan, d_aa.ptr))
            _capi.check(L.b200CsrPlanSetSummation(plan, 0))
            for lanes in (0, 1, 4, 32):
                H.csr_plan_set_layout(plan, lanes=lanes)
                H.spmv(plan, d_aa, d_x, d_y)
                assert np.array_equal(d_y.download(), ref), (name, nb, lanes)
                H.spmv_add(plan, d_aa, d_x, d_y0, d_z)                      # z = y0 + A x, out of place
                assert np.array_equal(d_z.download(), refadd), (name, nb, lanes)
                d_z.upload(y0)
                H.spmv_add(plan, d_aa, d_x, d_z, d_z)                       # in place
                assert np.array_equal(d_z.download(), refadd), (name, nb, lanes)
                H.spmv_jacobi(plan, d_aa, d_x, d_dinv, d_w, d_y)            # w = dinv .* (A x), y = A x
                assert np.array_equal(d_y.download(), ref) and np.array_equal(d_w.download(), ref * dinv), (name, nb, lanes)
                H.spmv_jacobi(plan, d_aa, d_x, d_dinv, d_w)                 # without the plain product
                assert np.array_equal(d_w.download(), ref * dinv), (name, nb, lanes)
            _capi.check(L.b200CsrPlanSetSummation(plan, 1))
            H.csr_plan_set_layout(plan, lanes=8)
            H.spmv(plan, d_aa, d_x, d_y)
            assert np.abs(d_y.download() - ref).max() <= RTOL * scale, (name, nb)
        # new values: pack again; then drop the blocks
        aa2 = aa * 0.5
        d_aa.upload(aa2)
        _capi.check(L.b200CsrPlanPackValues(H.h, plan, d_aa.ptr))
        _capi.check(L.b200CsrPlanSetSummation(plan, 0))
        H.spmv(plan, d_aa, d_x, d_y)
        assert np.array_equal(d_y.download(), oracle.matmult(ai, aj, aa2, x)), name
        _capi.check(L.b200CsrPlanSetColumnBlocks(H.h, plan, 0))
        H.csr_plan_set_layout(plan, lanes=1)
        H.spmv(plan, d_aa, d_x, d_y)
        assert np.array_equal(d_y.download(), oracle.matmult(ai, aj, aa2, x)), name
        L.b200CsrPlanDestroy(plan)


def test_spmv_fixture_vs_reference(H, oracle):
    """Committed outputs of the reference's own MatMult_SeqAIJ / MatMultAdd / GetDiagonal / PCApply_Jacobi."""
    from petsc_b200 import _capi
    L = _capi.lib()
    for path in sorted(glob.glob(golden_path("ops_*.npz"))):
        g = np.load(path)
        if str(g["gen"]) == "stored":
            ai, aj, aa = g["ai"], g["aj"], g["aa"]
        else:
            ai, aj, aa = getattr(oracle, str(g["gen"]))(*[int(a) for a in g["args"]])
        m = len(ai) - 1
        d_ai, d_aj, d_aa, plan = upload_csr(H, ai, aj, aa)
        H.csr_plan_set_layout(plan, lanes=1)
        d_x, d_y, d_z = H.array(g["x"]), H.array(g["y"]), H.empty(m)
        H.spmv(plan, d_aa, d_x, d_z)
        assert np.array_equal(d_z.download(), g["ref_mult"]), path
        H.spmv_add(plan, d_aa, d_x, d_y, d_z)
        assert np.array_equal(d_z.download(), g["ref_multadd"]), path
        d_diag, d_pos, d_dinv, d_w = H.empty(m), H.empty(m, np.int32), H.empty(m), H.empty(m)
        _capi.check(L.b200CsrGetDiagonal(H.h, m, d_ai.ptr, d_aj.ptr, d_aa.ptr, d_diag.ptr, d_pos.ptr))
        assert np.array_equal(d_diag.download(), g["ref_diag"])
        assert np.array_equal(aj[d_pos.download()], np.arange(m))
        nz = C.c_int(-1)
        _capi.check(L.b200JacobiInvertDiagonal(H.h, C.c_int64(m), d_diag.ptr, d_dinv.ptr, C.byref(nz)))
        assert nz.value == 0
        # PCApply_Jacobi = VecPointwiseMult(x, dinv)
        _capi.check(L.b200VecPointwiseMult(H.h, C.c_int64(m), d_x.ptr, d_dinv.ptr, d_w.ptr))
        assert np.array_equal(d_w.download(), g["ref_jacobi"])
        # fused SpMV + Jacobi == PCApply_Jacobi(MatMult(x)) bit for bit
        H.spmv_jacobi(plan, d_aa, d_x, d_dinv, d_w, d_z)
        assert np.array_equal(d_z.download(), g["ref_mult"])
        assert np.array_equal(d_w.download(), g["ref_mult"] * (1.0 / g["ref_diag"]))


def test_blas1_fixture_vs_reference(H, oracle):
    for path in sorted(glob.glob(golden_path("ops_*.npz"))):
        g = np.load(path)
        x, y, V, alpha = g["x"], g["y"], g["V"], g["alpha"]
        n, nv = len(x), V.shape[0]
        d_x, d_y = H.array(x), H.array(y)
        d_V = [H.array(V[j]) for j in range(nv)]
        # VecMAXPY_Seq: bit-exact (elementwise, same association, no FMA)
        nrm = H.maxpy(n, alpha, d_V, d_y, want_norm=True)
        assert np.array_equal(d_y.download(), g["ref_maxpy"]), path
        assert np.isclose(nrm, np.linalg.norm(g["ref_maxpy"]), rtol=RTOL)
        # VecMDot: tolerance (summation order differs from both the in-tree loop and dgemv)
        got = H.mdot(n, d_x, d_V)
        scale = np.linalg.norm(x) * np.linalg.norm(V, axis=1)
        assert np.all(np.abs(got - g["ref_mdot"]) <= RTOL * scale), path
        assert np.all(np.abs(got - g["ref_mdot_gemv"]) <= RTOL * scale), path
        dn = g["ref_dotnorm"]
        d_y2 = H.array(y)
        assert abs(H.dot(n, d_x, d_y2) - dn[0]) <= RTOL * np.linalg.norm(x) * np.linalg.norm(y)
        assert np.isclose(H.norm2(n, d_x), dn[1], rtol=RTOL)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 100003, 1 << 20])
def test_blas1_elementwise_and_reductions(H, oracle, n):
    from petsc_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(n + 5)
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    a, b = 0.37, -1.25
    N = C.c_int64(n)
    d_x, d_y, d_w = H.array(x), H.array(y), H.empty(n)
    dd = C.c_double

    def run(f, *args):
        _capi.check(f(H.h, N, *args))

    run(L.b200VecWAXPY, dd(a), d_x.ptr, d_y.ptr, d_w.ptr); assert np.allclose(d_w.download(), a * x + y, rtol=1e-14, atol=4e-16)
    run(L.b200VecPointwiseMult, d_x.ptr, d_y.ptr, d_w.ptr); assert np.array_equal(d_w.download(), x * y)
    run(L.b200VecAXPY, dd(a), d_x.ptr, d_y.ptr); y1 = y + a * x; assert np.allclose(d_y.download(), y1, rtol=1e-14, atol=4e-16)
    run(L.b200VecAYPX, dd(b), d_x.ptr, d_y.ptr); y2 = x + b * y1; assert np.allclose(d_y.download(), y2, rtol=1e-14, atol=4e-16)
    run(L.b200VecAXPBY, dd(a), dd(b), d_x.ptr, d_y.ptr); y3 = a * x + b * y2; assert np.allclose(d_y.download(), y3, rtol=1e-14, atol=1e-15)
    y3 = d_y.download(); run(L.b200VecScale, dd(b), d_y.ptr); assert np.array_equal(d_y.download(), b * y3)
    run(L.b200VecCopy, d_x.ptr, d_w.ptr); assert np.array_equal(d_w.download(), x)
    run(L.b200VecSet, dd(2.5), d_w.ptr); assert np.all(d_w.download() == 2.5)
    run(L.b200VecReciprocal, d_w.ptr); assert np.all(d_w.download() == 0.4)
    y4 = b * y3
    if n:
        assert abs(H.dot(n, d_x, d_y) - float(np.dot(x, y4))) <= RTOL * np.linalg.norm(x) * np.linalg.norm(y4) + 1e-300
        assert np.isclose(H.norm2(n, d_x), np.linalg.norm(x), rtol=RTOL)
        r = dd()
        _capi.check(L.b200VecNorm(H.h, N, d_x.ptr, 0, C.byref(r))); assert np.isclose(r.value, np.abs(x).sum(), rtol=RTOL)
        _capi.check(L.b200VecNorm(H.h, N, d_x.ptr, 3, C.byref(r))); assert r.value == np.abs(x).max()
        _capi.check(L.b200VecSum(H.h, N, d_x.ptr, C.byref(r))); assert abs(r.value - x.sum()) <= RTOL * np.abs(x).sum()
        idx = C.c_int64()
        _capi.check(L.b200VecMax(H.h, N, d_x.ptr, C.byref(idx), C.byref(r))); assert r.value == x.max() and idx.value == int(np.argmax(x))
        _capi.check(L.b200VecMin(H.h, N, d_x.ptr, C.byref(idx), C.byref(r))); assert r.value == x.min() and idx.value == int(np.argmin(x))
        # fused AXPY+dot
        d_y5 = H.array(y)
        _capi.check(L.b200VecAXPYDot(H.h, N, dd(a), d_x.ptr, d_y5.ptr, d_y5.ptr, C.byref(r)))
        assert np.isclose(r.value, float(np.dot(y + a * x, y + a * x)), rtol=1e-12)
    else:
        assert H.dot(0, d_x, d_y) == 0.0 and H.norm2(0, d_x) == 0.0


@pytest.mark.parametrize("nv", [1, 2, 3, 4, 5, 7, 8, 13, 30, 31, 32, 33, 47, 70])
def test_mdot_maxpy_all_nv(H, oracle, nv):
    rng = np.random.default_rng(nv)
    for n in (1031, 65536 + 3):
        x = rng.uniform(-1, 1, n); V = rng.uniform(-1, 1, (nv, n)); alpha = rng.uniform(-1, 1, nv)
        d_x = H.array(x); d_V = [H.array(V[j]) for j in range(nv)]
        got = H.mdot(n, d_x, d_V)
        ref = V @ x
        assert np.all(np.abs(got - ref) <= RTOL * np.linalg.norm(x) * np.linalg.norm(V, axis=1))
        ys = [np.ascontiguousarray(V[j]) for j in range(nv)]
        ref_x = oracle.vecmaxpy(x.copy(), alpha, ys)
        nrm = H.maxpy(n, alpha, d_V, d_x, want_norm=True)
        assert np.array_equal(d_x.download(), ref_x), (nv, n)  # bit-exact vs the VecMAXPY_Seq restatement
        assert np.isclose(nrm, np.linalg.norm(ref_x), rtol=RTOL)


def test_unaligned_subvectors(H, oracle):
    """Local-vector aliases (VecGetLocalVector in PCApply_BJacobi) may start at any 8-byte offset."""
    from petsc_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(3)
    n = 10007
    x, y = rng.uniform(-1, 1, n + 3), rng.uniform(-1, 1, n + 3)
    d_x, d_y = H.array(x), H.array(y)
    px, py = d_x.offset(1), d_y.offset(3)
    r = C.c_double()
    _capi.check(L.b200VecDot(H.h, C.c_int64(n), px, py, C.byref(r)))
    assert abs(r.value - np.dot(x[1:n + 1], y[3:n + 3])) <= RTOL * n
    _capi.check(L.b200VecAXPY(H.h, C.c_int64(n), C.c_double(0.5), px, py))
    ref = y.copy(); ref[3:n + 3] += 0.5 * x[1:n + 1]
    assert np.allclose(d_y.download(), ref, rtol=1e-14, atol=4e-16)


def test_device_laplace7_generator_matches_oracle(H, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    for (nx, ny, nz, r0, r1) in ((7, 5, 4, 0, 140), (7, 5, 4, 33, 97), (16, 16, 16, 1000, 4096), (1, 1, 1, 0, 1), (3, 1, 2, 0, 6)):
        ai, aj, aa = oracle.lap7(nx, ny, nz)
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace7Nnz(nx, ny, nz, C.c_int64(r0), C.c_int64(r1), C.byref(nnz)))
        assert nnz.value == ai[r1] - ai[r0]
        d_ai, d_aj, d_aa = H.empty(r1 - r0 + 1, np.int32), H.empty(nnz.value, np.int32), H.empty(nnz.value)
        _capi.check(L.b200GenLaplace7(H.h, nx, ny, nz, C.c_int64(r0), C.c_int64(r1), d_ai.ptr, d_aj.ptr, d_aa.ptr))
        assert np.array_equal(d_ai.download(), ai[r0:r1 + 1] - ai[r0])
        assert np.array_equal(d_aj.download(), aj[ai[r0]:ai[r1]])
        assert np.array_equal(d_aa.download(), aa[ai[r0]:ai[r1]])


def test_error_paths(H):
    from petsc_b200 import _capi
    L = _capi.lib()
    assert L.b200VecAXPY(H.h, C.c_int64(-1), C.c_double(1), None, None) == 63  # PETSC_ERR_ARG_OUTOFRANGE
    assert L.b200VecAXPY(H.h, C.c_int64(4), C.c_double(1), None, None) == 85   # PETSC_ERR_ARG_NULL
    assert b"null" in L.b200GetLastErrorString()
    p = C.c_void_p()
    assert L.b200CsrPlanCreate(H.h, 1, 1, C.c_int64(1 << 40), None, None, C.byref(p)) == 56  # PETSC_ERR_SUP (32-bit PetscInt)


def test_launch_counter(H):
    from petsc_b200 import _capi
    a = _capi.launch_count()
    d = H.zeros(1000)
    _capi.check(_capi.lib().b200VecScale(H.h, C.c_int64(1000), C.c_double(2.0), d.ptr))
    assert _capi.launch_count() == a + 1


def test_device_laplace27_and_random_generators(H, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    for n in (2, 5, 9):
        ai, aj, aa = oracle.lap27(n)
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
        assert nnz.value == len(aj)
        d_i, d_j, d_a = H.empty(n ** 3 + 1, np.int32), H.empty(nnz.value, np.int32), H.empty(nnz.value)
        _capi.check(L.b200GenLaplace27(H.h, n, d_i.ptr, d_j.ptr, d_a.ptr))
        assert np.array_equal(d_i.download(), ai) and np.array_equal(d_j.download(), aj) and np.array_equal(d_a.download(), aa)
    n, d = 5000, 32
    d_i, d_j, d_a = H.empty(n + 1, np.int32), H.empty(n * d, np.int32), H.empty(n * d)
    _capi.check(L.b200GenRandomCsr(H.h, n, n, d, C.c_uint64(7), d_i.ptr, d_j.ptr, d_a.ptr))
    ai, aj, aa = d_i.download(), d_j.download().reshape(n, d), d_a.download()
    assert np.array_equal(ai, np.arange(n + 1) * d) and np.all(np.diff(aj, axis=1) > 0) and aj.min() >= 0 and aj.max() < n
    assert np.all(np.abs(aa) < 1.0) and abs(aa.mean()) < 0.01
    x = np.random.default_rng(0).uniform(-1, 1, n)
    plan = H.csr_plan(n, n, n * d, d_i, d_j)
    d_x, d_y = H.array(x), H.empty(n)
    H.csr_plan_set_layout(plan, lanes=1)
    H.spmv(plan, d_a, d_x, d_y)
    assert np.array_equal(d_y.download(), oracle.matmult(ai, aj.reshape(-1), aa, x))


# ---------------------------------------------------------------------------------------------- streaming CSR-vector kernel
def _spmv_all(H, ai, aj, aa, x, y, dinv):
    """y0 = A x, z = y + A x, w = dinv .* (A x) through one plan (whatever kernel the plan chose)."""
    from petsc_b200 import _capi
    L = _capi.lib()
    m, nnz = len(ai) - 1, len(aj)
    n = len(x)
    d_i, d_j, d_a = H.array(ai, np.int32), H.array(aj, np.int32), H.array(aa, np.float64)
    plan = C.c_void_p()
    _capi.check(L.b200CsrPlanCreate(H.h, m, n, C.c_int64(nnz), d_i.ptr, d_j.ptr, C.byref(plan)))
    d_x, d_y, d_d = H.array(x), H.array(y), H.array(dinv)
    o1, o2, o3 = H.empty(m), H.empty(m), H.empty(m)
    _capi.check(L.b200CsrSpMV(H.h, plan, d_a.ptr, d_x.ptr, o1.ptr))
    _capi.check(L.b200CsrSpMVAdd(H.h, plan, d_a.ptr, d_x.ptr, d_y.ptr, o2.ptr))
    _capi.check(L.b200CsrSpMVJacobi(H.h, plan, d_a.ptr, d_x.ptr, d_d.ptr, o3.ptr, None))
    lay = [C.c_int() for _ in range(6)]
    _capi.check(L.b200CsrPlanGetLayout(plan, *[C.byref(v) for v in lay]))
    res = o1.download(), o2.download(), o3.download()
    L.b200CsrPlanDestroy(plan)
    for o in (d_i, d_j, d_a, d_x, d_y, d_d, o1, o2, o3):
        o.free()
    return res


@pytest.mark.parametrize("lanes", [2, 8, 32])
def test_spmv_vector_kernel_forced(H, oracle, lanes, monkeypatch):
    """The streaming CSR-vector kernel (auto-selected for scattered / very long rows) forced onto every test matrix: MatMult,
    MatMultAdd and the Jacobi epilogue within 1e-12 of MatMult_SeqAIJ (tree-ordered row sums, like every multi-lane layout)."""
    monkeypatch.setenv("PETSCB200_SPMV_VECTOR", str(lanes))
    rng = np.random.default_rng(21)
    for name, (ai, aj, aa) in matrices(oracle):
        m = len(ai) - 1
        n = int(max(aj.max() + 1 if len(aj) else 1, m))
        x, y, dinv = rng.uniform(-1, 1, n), rng.uniform(-1, 1, m), rng.uniform(0.5, 2, m)
        ref = oracle.matmult(ai, aj, aa, x) if n == m else None
        if ref is None:
            ref = np.array([np.dot(aa[ai[r]:ai[r + 1]], x[aj[ai[r]:ai[r + 1]]]) for r in range(m)])
        scale = np.array([np.abs(aa[ai[r]:ai[r + 1]] * x[aj[ai[r]:ai[r + 1]]]).sum() for r in range(m)]) + 1e-300
        y0, z, w = _spmv_all(H, ai, aj, aa, x, y, dinv)
        assert np.all(np.abs(y0 - ref) <= 1e-12 * scale), name
        assert np.all(np.abs(z - (y + ref)) <= 1e-12 * (scale + np.abs(y))), name
        assert np.all(np.abs(w - dinv * ref) <= 1e-12 * scale * dinv), name


def test_spmv_power_law_row_bins(H, oracle):
    """A power-law matrix: 2.5 M rows of ~4 scattered entries plus a tail of very long rows (up to 600 k entries).  The plan must
    pick the streaming kernel by itself (scattered columns), give the tail its own one-CTA-per-row bin, and agree with
    MatMult_SeqAIJ to 1e-12."""
    from petsc_b200 import _capi
    L = _capi.lib()
    n = 2_500_000
    rng = np.random.default_rng(33)
    ai0, aj0, aa0 = oracle.random_csr(n, 4, 5)
    lens = np.full(n, 4, np.int64)
    longrows = {7: 600_000, 1000: 120_000, 123_456: 40_000, 2_000_001: 9_000, n - 1: 20_000}
    for r, l in longrows.items():
        lens[r] = l
    ai = np.zeros(n + 1, np.int64); ai[1:] = np.cumsum(lens)
    aj = np.empty(ai[-1], np.int32); aa = np.empty(ai[-1])
    short = np.ones(n, bool); short[list(longrows)] = False
    # short rows keep their 4 entries
    idx_new = (ai[:-1][short][:, None] + np.arange(4)[None, :]).ravel()
    idx_old = (ai0[:-1].astype(np.int64)[short][:, None] + np.arange(4)[None, :]).ravel()
    aj[idx_new] = aj0[idx_old]; aa[idx_new] = aa0[idx_old]
    for r, l in longrows.items():
        cols = np.sort(rng.choice(n, l, replace=False)).astype(np.int32)
        aj[ai[r]:ai[r + 1]] = cols; aa[ai[r]:ai[r + 1]] = rng.uniform(-1, 1, l)
    ai = ai.astype(np.int32)
    x = rng.uniform(-1, 1, n)
    ref = oracle.matmult(ai, aj, aa, x)
    d_i, d_j, d_a, d_x, d_y = H.array(ai, np.int32), H.array(aj, np.int32), H.array(aa), H.array(x), H.empty(n)
    plan = C.c_void_p()
    _capi.check(L.b200CsrPlanCreate(H.h, n, n, C.c_int64(len(aj)), d_i.ptr, d_j.ptr, C.byref(plan)))
    before = _capi.launch_count()
    _capi.check(L.b200CsrSpMV(H.h, plan, d_a.ptr, d_x.ptr, d_y.ptr))
    assert _capi.launch_count() - before == 3            # streaming kernel + the long-row bin (segment partials, per-row finish)
    y = d_y.download()
    scale = np.abs(ref) + 1e-9 * np.sqrt(lens)
    assert np.all(np.abs(y - ref) <= 1e-12 * np.maximum(scale, np.sqrt(lens))), float(np.abs(y - ref).max())
    for r in longrows:
        assert abs(y[r] - ref[r]) <= 1e-12 * np.sqrt(lens[r]) * 10
    L.b200CsrPlanDestroy(plan)
    for o in (d_i, d_j, d_a, d_x, d_y):
        o.free()
