"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run these sizes in
seconds): 7-point 512^3 (configs[1], 134 M rows / 938 M nonzeros) and 27-point 256^3 (configs[2]).

Exact properties (bit for bit, no tolerance):
  * checksum of A*1: the row sums of the 7-point operator are the integers 6 - #neighbours, so sum(A*1) = 7N - nnz exactly;
  * scaling by a power of two commutes with every kernel exactly: A(2x) = 2(Ax), M^-1(2u) = 2 M^-1 u (ILU(0) sweeps);
  * the operator is symmetric with sorted rows, so MatMultTranspose(A, x) = MatMult(A, x) bit for bit.
Tolerance properties (north_star: 1e-12 relative): linearity, symmetry of the bilinear form, VecMDot vs VecDot, VecMAXPY
undone by VecAXPYs, fused norm vs recomputed norm, GMRES residual estimate vs the true preconditioned residual."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def P():
    from harness import petsc
    petsc.initialize()  # idempotent; the device was chosen by whichever module initialised first (cuda:0 by default)
    return petsc


def device_matrix(petsc, gen, n):
    """The benchmark operator generated on the device and adopted by a seqaijb200 matrix."""
    from petsc_b200 import _capi
    L = _capi.lib()
    H = petsc.handle()
    Hh = type("Hh", (), {"h": H})
    N = n ** 3
    nnz = C.c_int64()
    if gen == 7:
        _capi.check(L.b200GenLaplace7Nnz(n, n, n, C.c_int64(0), C.c_int64(N), C.byref(nnz)))
    else:
        _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
    nnz = nnz.value
    d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
    if gen == 7:
        _capi.check(L.b200GenLaplace7(H, n, n, n, C.c_int64(0), C.c_int64(N), d_i.ptr, d_j.ptr, d_a.ptr))
    else:
        _capi.check(L.b200GenLaplace27(H, n, d_i.ptr, d_j.ptr, d_a.ptr))
    A = petsc.Mat.create(m=N, n=N, M=N, N=N, comm=petsc.COMM_SELF, mtype="seqaijb200")
    A.set_csr_device(d_i.ptr, d_j.ptr, d_a.ptr)
    for o in (d_i, d_j, d_a):
        o.free()
    return A, N, nnz


@pytest.fixture(scope="module")
def lap7_512(P):
    free, total = C.c_size_t(), C.c_size_t()
    from petsc_b200 import _capi
    _capi.check(_capi.lib().b200MemGetInfo(C.byref(free), C.byref(total)))
    if free.value < 120 * (1 << 30):
        pytest.skip("needs ~100 GB of free device memory (B200: 180 GB)")
    A, N, nnz = device_matrix(P, 7, 512)
    yield A, N, nnz
    A.destroy()


def random_vec(petsc, like, seed):
    n = like.local_size()
    v = like.duplicate()
    v.set_array(np.random.default_rng(seed).uniform(-1.0, 1.0, n))
    return v


def test_fullsize_spmv_checksum_and_exact_scaling(P, lap7_512):
    A, N, nnz = lap7_512
    n = 512
    assert N == n ** 3 and nnz == 7 * n ** 3 - 6 * n ** 2          # 937 951 232
    x, b = A.create_vecs()
    x.set(1.0)
    A.mult(x, b)
    assert b.sum() == float(7 * N - nnz)                            # checksum of checksums: exact integer arithmetic
    assert b.max()[1] == 3.0 and b.min()[1] == 0.0                  # corners have 3 neighbours, interior rows sum to 0
    r = random_vec(P, x, 1)
    y1, y2 = x.duplicate(), x.duplicate()
    A.mult(r, y1)
    r.scale(2.0)
    A.mult(r, y2)                                                   # A(2x)
    y1.scale(2.0)                                                   # 2(Ax)
    y2.axpy(-1.0, y1)
    assert y2.norm(3) == 0.0                                        # NORM_INFINITY: bit-identical
    for o in (x, b, r, y1, y2):
        o.destroy()


def test_fullsize_spmv_linearity_symmetry_transpose(P, lap7_512):
    A, N, nnz = lap7_512
    x, _ = A.create_vecs()
    y = random_vec(P, x, 3)
    x.destroy()
    x = random_vec(P, y, 2)
    ax, ay, w, z = x.duplicate(), x.duplicate(), x.duplicate(), x.duplicate()
    A.mult(x, ax)
    A.mult(y, ay)
    w.waxpy(-1.5, y, x)                                             # w = x - 1.5 y
    A.mult(w, z)                                                    # A(x - 1.5 y)
    w.waxpy(-1.5, ay, ax)                                           # Ax - 1.5 Ay
    scale = w.norm()
    z.axpy(-1.0, w)
    assert z.norm() <= RTOL * scale
    # symmetric bilinear form: y.(Ax) = x.(Ay)
    assert abs(y.dot(ax) - x.dot(ay)) <= RTOL * x.norm() * ay.norm()
    # A is symmetric with sorted rows: the explicit-transpose product repeats MatMult bit for bit (full-size sort + gather)
    A.mult_transpose(x, z)
    z.axpy(-1.0, ax)
    assert z.norm(3) == 0.0
    for o in (x, y, ax, ay, w, z):
        o.destroy()


def test_fullsize_orthogonalisation_ops(P, lap7_512):
    """VecMDot / VecMAXPY(+fused norm) at n = 512^3 with nv = 30 (the last GMRES(30) iteration's shapes)."""
    A, N, nnz = lap7_512
    petsc = P
    x, _ = A.create_vecs()
    w = random_vec(petsc, x, 5)
    V, arr = petsc.duplicate_vecs(x, 30)
    V[0].set_array(np.random.default_rng(6).uniform(-1.0, 1.0, N))
    V[0].normalize()
    for j in range(29):                                             # a Krylov-like basis made on the device
        A.mult(V[j], V[j + 1])
        V[j + 1].normalize()
    wn = w.norm()
    md = w.mdot(V)
    single = np.array([w.dot(v) for v in V])
    assert np.all(np.abs(md - single) <= RTOL * wn)                 # unit vectors: |w.v| <= |w|
    alpha = np.linspace(-1.0, 1.0, 30) + 0.013
    w2 = w.duplicate()
    w.copy_to(w2)
    w2.maxpy(alpha, V)
    fused = w2.norm()                                               # served by the sum of squares the MAXPY kernel left behind
    chk = w2.duplicate()
    w2.copy_to(chk)
    assert abs(chk.norm() - fused) <= RTOL * fused                  # an independent norm kernel on a copy
    for j in range(30):
        w2.axpy(-alpha[j], V[j])
    w2.axpy(-1.0, w)
    assert w2.norm() <= 64 * 2.3e-16 * (wn + np.abs(alpha).sum())   # 60 roundings of O(1) quantities per entry
    petsc.destroy_vecs(30, arr)
    for o in (x, w, w2, chk):
        o.destroy()


def test_fullsize_gmres_cycle_residual_is_true_residual(P, lap7_512):
    """BASELINE configs[1] itself: one GMRES(30)+Jacobi restart cycle; the norms from the Givens recurrence are monotone and the last
    one is the true preconditioned residual ||D^-1 (b - A x)|| of the iterate that BuildSoln produced."""
    A, N, nnz = lap7_512
    petsc = P
    x, b = A.create_vecs()
    u = x.duplicate(); u.set(1.0)
    A.mult(u, b)
    petsc.options_clear()
    petsc.options_insert("-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-300 -ksp_max_it 30")
    ksp = petsc.KSP.create(petsc.COMM_SELF)
    ksp.set_operators(A); ksp.set_residual_history(); ksp.set_from_options()
    ksp.solve(b, x)
    h = ksp.history()
    assert ksp.its() == 30 and ksp.reason() == -3 and len(h) == 31
    assert np.all(np.diff(h) <= 0.0)
    r, d = x.duplicate(), x.duplicate()
    A.mult(x, r)
    r.aypx(-1.0, b)                                                 # r = b - A x
    A.get_diagonal(d)
    d.reciprocal()
    r.pointwise_mult(r, d)
    assert abs(r.norm() - ksp.rnorm()) <= 1e-9 * h[0]
    # the same cycle again from the same state: deterministic kernels -> the identical history, bit for bit
    x.set(0.0)
    ksp.solve(b, x)
    assert np.array_equal(ksp.history(), h)
    ksp.destroy()
    for o in (x, b, u, r, d):
        o.destroy()
    petsc.options_clear()


def test_fullsize_cg_ilu0_27pt_256(P):
    """BASELINE configs[2]: 27-point 256^3, CG + ILU(0) (factor and sweeps on the device)."""
    petsc = P
    from petsc_b200 import _capi
    free, total = C.c_size_t(), C.c_size_t()
    _capi.check(_capi.lib().b200MemGetInfo(C.byref(free), C.byref(total)))
    if free.value < 60 * (1 << 30):
        pytest.skip("needs ~40 GB of free device memory")
    A, N, nnz = device_matrix(petsc, 27, 256)
    assert nnz == 449455096                                          # SURVEY 8 / bench_kspsolve
    x, b = A.create_vecs()
    u = x.duplicate(); u.set(1.0)
    A.mult(u, b)
    petsc.options_clear()
    petsc.options_insert("-ksp_type cg -pc_type ilu -ksp_rtol 1e-8")
    ksp = petsc.KSP.create(petsc.COMM_SELF)
    ksp.set_operators(A); ksp.set_residual_history(); ksp.set_from_options()
    ksp.solve(b, x)
    assert ksp.reason() == 2 and 130 <= ksp.its() <= 165             # 147 measured; the oracle cannot run this size
    x.axpy(-1.0, u)
    assert x.norm(3) < 1e-5                                          # exact solution = ones (bench_kspsolve convention)
    # the ILU(0) application is linear: a power-of-two scaling passes through both sweeps exactly
    pc = ksp.get_pc()
    r = random_vec(petsc, x, 9)
    z1, z2 = x.duplicate(), x.duplicate()
    pc.apply(r, z1)
    r.scale(4.0)
    pc.apply(r, z2)
    z1.scale(4.0)
    z2.axpy(-1.0, z1)
    assert z2.norm(3) == 0.0
    ksp.destroy()
    for o in (x, b, u, r, z1, z2):
        o.destroy()
    A.destroy()
    petsc.options_clear()


def test_fullsize_spmv_sampled_rows_bit_exact_vs_reference_order(P, lap7_512):
    """The 512^3 product meets the reference arithmetic directly on 20 000 sampled rows: each sampled row of the 7-point operator
    is rebuilt from its definition (diag 6, off-diagonals -1, columns ascending, Dirichlet by truncation) and summed strictly left
    to right starting from 0.0 -- PetscSparseDensePlusDot (aij.h:609-614) -- with IEEE doubles; the kernel (one lane per row) must
    give the same bits."""
    A, N, nnz = lap7_512
    n = 512
    x, y = A.create_vecs()
    xs = np.random.default_rng(77).uniform(-1.0, 1.0, N)
    x.set_array(xs)
    A.mult(x, y)
    ya = y.array()
    rng = np.random.default_rng(78)
    rows = np.unique(np.concatenate([rng.integers(0, N, 20000), [0, 1, n - 1, n, n * n - 1, n * n, N - 1, N - n, N - n * n]]))
    ix, iy, iz = rows % n, (rows // n) % n, rows // (n * n)
    ref = np.zeros(len(rows))
    # ascending columns: -n*n, -n, -1, diag, +1, +n, +n*n  (present iff inside the grid); strict left-to-right accumulation
    for off, ok, val in ((-n * n, iz > 0, -1.0), (-n, iy > 0, -1.0), (-1, ix > 0, -1.0), (0, np.ones(len(rows), bool), 6.0), (1, ix < n - 1, -1.0), (n, iy < n - 1, -1.0), (n * n, iz < n - 1, -1.0)):
        term = val * xs[np.clip(rows + off, 0, N - 1)]
        ref = np.where(ok, ref + term, ref)
    assert np.array_equal(ya[rows], ref)
    x.destroy(); y.destroy()


def test_fullsize_ilu0_27pt_subproblem_meets_oracle(P, oracle):
    """The 27-point generator + ILU(0) factor + both sweeps at 64^3 (262 144 rows: the largest the OpenMP-free oracle does in
    seconds) are bit-identical to MatLUFactorNumeric_SeqAIJ + MatSolve_SeqAIJ_NaturalOrdering; the 256^3 run of the same code
    is covered by the exact-scaling test above."""
    from petsc_b200 import _capi
    L = _capi.lib()
    H = P.handle()
    Hh = type("Hh", (), {"h": H})
    n = 64
    N = n ** 3
    ai, aj, aa = oracle.lap27(n)
    nnz = C.c_int64()
    _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
    assert nnz.value == len(aj)
    d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, len(aj), np.int32), _capi.DeviceArray(Hh, len(aj), np.float64)
    _capi.check(L.b200GenLaplace27(H, n, d_i.ptr, d_j.ptr, d_a.ptr))
    assert np.array_equal(d_i.download(), ai) and np.array_equal(d_j.download(), aj) and np.array_equal(d_a.download(), aa)   # generator == bench_kspsolve.c:115-303 restatement
    plan = C.c_void_p()
    _capi.check(L.b200Ilu0Symbolic(H, N, ai.ctypes.data_as(C.c_void_p), aj.ctypes.data_as(C.c_void_p), C.byref(plan)))
    ns = C.c_int(-1)
    eps100 = 100 * 2.220446049250313e-16
    _capi.check(L.b200Ilu0Numeric(H, plan, d_a.ptr, C.c_double(eps100), C.c_double(eps100), C.byref(ns)))
    bi, bj, bd, ba = oracle.ilu0(ai, aj, aa)
    b = np.random.default_rng(9).uniform(-1, 1, N)
    d_b = _capi.DeviceArray(Hh, N, np.float64).upload(b); d_x = _capi.DeviceArray(Hh, N, np.float64)
    _capi.check(L.b200Ilu0Solve(H, plan, d_b.ptr, d_x.ptr))
    assert np.array_equal(d_x.download(), oracle.matsolve(bi, bj, bd, ba, b))
    _capi.check(L.b200Ilu0Destroy(plan))
    for o in (d_i, d_j, d_a, d_b, d_x):
        o.free()
