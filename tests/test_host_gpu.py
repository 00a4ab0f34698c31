"""GPU parity tests through the C host mirror (include/petscb200_host.h): -mat_type aijb200 -vec_type b200 with
KSPGMRES / KSPCG and PCJACOBI / PCILU / PCBJACOBI, against the reference fixtures, the reference's golden outputs and
the CPU oracle."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
from conftest import assert_history_1e12

from conftest import golden_path

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def P():
    from harness import petsc
    petsc.initialize()
    yield petsc
    petsc.options_clear()


def sig6(v):
    return float("%.6g" % v)


def load_matrix(O, g):
    if str(g["gen"]) == "stored":
        return g["ai"], g["aj"], g["aa"]
    return getattr(O, str(g["gen"]))(*[int(a) for a in g["args"]])


def solve(P, ai, aj, aa, opts, b=None):
    P.options_clear()
    P.options_insert(opts)
    A = P.Mat.from_csr(ai, aj, aa)
    n = len(ai) - 1
    x, bv = A.create_vecs()
    if b is None:
        u = x.duplicate(); u.set(1.0); A.mult(u, bv); u.destroy()
    else:
        bv.set_array(b)
    ksp = P.KSP.create(P.COMM_SELF)
    ksp.set_operators(A)
    ksp.set_residual_history()
    ksp.set_from_options()
    ksp.solve(bv, x)
    out = dict(x=x.array(), its=ksp.its(), reason=ksp.reason(), rnorm=ksp.rnorm(), hist=ksp.history())
    ksp.destroy(); x.destroy(); bv.destroy(); A.destroy()
    P.options_clear()
    return out


def test_types_and_registry(P):
    v = P.Vec.create(n=10, N=10, comm=P.COMM_SELF)
    assert v.get_type() == "seqb200"
    A = P.Mat.create(m=4, n=4, M=4, N=4, comm=P.COMM_SELF)
    assert A.get_type() == "seqaijb200"
    with pytest.raises(P.PetscError) as e:
        P.Vec.create(n=3, N=3, comm=P.COMM_SELF, vtype="nosuchtype")
    assert e.value.code == 86 and "Unknown vector type" in str(e.value)  # PETSC_ERR_ARG_UNKNOWN_TYPE, vecreg.c wording
    v.destroy(); A.destroy()


def test_ex2_assembly_by_matsetvalues(P, oracle):
    """ex2.c:70-92 builds the 5-point operator with MatSetValues(ADD_VALUES) row by row; the assembled CSR must equal the
    generator's (MatAssemblyEnd_SeqAIJ compaction: sorted columns, duplicates summed)."""
    m, n = 6, 5
    A = P.Mat.create(m=m * n, n=m * n, M=m * n, N=m * n, comm=P.COMM_SELF)
    for Ii in range(m * n):
        i, j = divmod(Ii, n)
        if i > 0: A.set_values([Ii], [Ii - n], [-1.0])
        if i < m - 1: A.set_values([Ii], [Ii + n], [-1.0])
        if j > 0: A.set_values([Ii], [Ii - 1], [-1.0])
        if j < n - 1: A.set_values([Ii], [Ii + 1], [-1.0])
        A.set_values([Ii], [Ii], [2.0]); A.set_values([Ii], [Ii], [2.0])  # duplicates are added
    A.assemble()
    ai, aj, aa = A.csr_host()
    rai, raj, raa = oracle.lap5(m, n)
    assert np.array_equal(ai, rai) and np.array_equal(aj, raj) and np.array_equal(aa, raa)
    x, y = A.create_vecs()
    x.set(1.0); A.mult(x, y)
    assert np.array_equal(y.array(), oracle.matmult(rai, raj, raa, np.ones(m * n)))
    with pytest.raises(P.PetscError) as e:
        A.mult(x, x)
    assert e.value.code == 61  # PETSC_ERR_ARG_IDN
    w = P.Vec.create(n=3, N=3, comm=P.COMM_SELF)
    with pytest.raises(P.PetscError) as e:
        A.mult(w, y)
    assert e.value.code == 60  # PETSC_ERR_ARG_SIZ
    for o in (x, y, w, A): o.destroy()


def test_vec_ops_offload_and_norm_cache(P, oracle):
    rng = np.random.default_rng(0)
    n = 5003
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x, y = P.Vec.from_array(a), P.Vec.from_array(b)
    y.axpy(0.5, x); ref = b + 0.5 * a
    assert np.allclose(y.array(), ref, rtol=1e-14, atol=1e-15)
    # host write access invalidates the device copy and the cached norm
    n1 = y.norm()
    assert np.isclose(n1, np.linalg.norm(ref), rtol=RTOL)
    assert y.norm() == n1
    ref2 = y.array(); ref2[7] = 100.0; y.set_array(ref2)
    assert np.isclose(y.norm(), np.linalg.norm(ref2), rtol=RTOL)
    y.scale(-2.0)
    assert np.isclose(y.norm(), 2 * np.linalg.norm(ref2), rtol=RTOL)
    # VecDuplicateVecs slab + MDot/MAXPY (+ fused norm served by VecNorm)
    vs, arr = P.duplicate_vecs(x, 7)
    V = rng.uniform(-1, 1, (7, n))
    for v, row in zip(vs, V): v.set_array(row)
    d = x.mdot(vs)
    assert np.all(np.abs(d - V @ a) <= RTOL * np.linalg.norm(a) * np.linalg.norm(V, axis=1))
    al = rng.uniform(-1, 1, 7)
    x.maxpy(al, vs)
    refx = oracle.vecmaxpy(a.copy(), al, [np.ascontiguousarray(r) for r in V])
    nrm = x.norm()
    assert np.array_equal(x.array(), refx)
    assert np.isclose(nrm, np.linalg.norm(refx), rtol=RTOL)
    assert np.isclose(x.normalize(), np.linalg.norm(refx), rtol=RTOL)
    assert np.isclose(x.norm(), 1.0, rtol=1e-14)  # cached norm scaled by VecScale (rvector.c:1010)
    assert np.isclose(x.dot(x), 1.0, rtol=1e-12)
    i, mx = y.max(); assert mx == y.array().max() and i == int(np.argmax(y.array()))
    assert abs(y.sum() - y.array().sum()) < 1e-9
    P.destroy_vecs(7, arr)
    x.destroy(); y.destroy()


OPS = sorted(glob.glob(golden_path("ops_*.npz")))
def _device_path_exists(path):
    """ICC(0) and pipelined-GMRES fixtures pin the ORACLE only so far (tests/test_oracle_golden.py; their device paths are the
    next round's): they are not run through the host mirror.  Whitelist, so that a new oracle-only fixture can never turn
    into a failing device test by accident."""
    opts = [str(s) for s in np.load(path)["opts"]]
    ksp = opts[opts.index("-ksp_type") + 1] if "-ksp_type" in opts else "gmres"
    pc = opts[opts.index("-pc_type") + 1] if "-pc_type" in opts else "ilu"
    return ksp in ("gmres", "cg", "pipecg", "preonly") and pc in ("none", "jacobi", "ilu", "bjacobi")


KSPF = [p for p in sorted(glob.glob(golden_path("ksp_*.npz"))) if _device_path_exists(p)]


@pytest.mark.parametrize("path", OPS, ids=[os.path.basename(p)[:-4] for p in OPS])
def test_pc_apply_bit_exact_vs_reference(P, oracle, path):
    g = np.load(path)
    ai, aj, aa = load_matrix(oracle, g)
    A = P.Mat.from_csr(ai, aj, aa)
    x = P.Vec.from_array(g["x"]); y = x.duplicate()
    for pctype, key in (("jacobi", "ref_jacobi"), ("ilu", "ref_ilusolve"), ("bjacobi", "ref_ilusolve")):
        pc = P.PC.create(P.COMM_SELF, pctype)
        pc.set_operators(A)
        pc.setup()
        pc.apply(x, y)
        assert np.array_equal(y.array(), g[key]), (pctype, path)  # PCApply_Jacobi / PCApply_ILU: bit-identical
        pc.apply(x, y)
        assert np.array_equal(y.array(), g[key])  # repeated solves reuse the schedule (epoch flags)
        pc.destroy()
    for o in (x, y, A): o.destroy()


def test_ilu0_factor_bit_exact_vs_oracle(P, oracle):
    from petsc_b200 import _capi
    L = _capi.lib()
    H = P.handle()
    for ai, aj, aa in (oracle.lap5(17, 13), oracle.lap7(9, 7, 8), oracle.lap27(9)):
        n = len(ai) - 1
        nnz = len(aj)
        plan = C.c_void_p()
        _capi.check(L.b200Ilu0Symbolic(H, n, ai.ctypes.data_as(C.c_void_p), aj.ctypes.data_as(C.c_void_p), C.byref(plan)))
        d_a = C.c_void_p()
        _capi.check(L.b200Malloc(H, C.byref(d_a), C.c_size_t(8 * nnz)))
        _capi.check(L.b200MemcpyHtoD(H, d_a, aa.ctypes.data_as(C.c_void_p), C.c_size_t(8 * nnz)))
        ns = C.c_int(-1)
        eps100 = 100 * 2.220446049250313e-16
        _capi.check(L.b200Ilu0Numeric(H, plan, d_a, C.c_double(eps100), C.c_double(eps100), C.byref(ns)))
        assert ns.value == 0
        bi = np.zeros(n + 1, np.int32); bd = np.zeros(n + 1, np.int32); bj = np.zeros(nnz, np.int32); ba = np.zeros(nnz)
        _capi.check(L.b200Ilu0GetFactor(H, plan, bi.ctypes.data_as(C.c_void_p), bj.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p), ba.ctypes.data_as(C.c_void_p)))
        obi, obj, obd, oba = oracle.ilu0(ai, aj, aa)
        assert np.array_equal(bi, obi) and np.array_equal(bj, obj) and np.array_equal(bd, obd)  # layout: index-exact
        assert np.array_equal(ba, oba)  # MatLUFactorNumeric_SeqAIJ: bit-identical factor
        lv = [C.c_int(), C.c_int()]
        _capi.check(L.b200Ilu0GetInfo(plan, C.byref(lv[0]), C.byref(lv[1]), None))
        assert lv[0].value >= 1 and lv[0].value == lv[1].value  # symmetric pattern: same depth both ways
        _capi.check(L.b200Ilu0Destroy(plan)); _capi.check(L.b200Free(H, d_a))


def test_ilu0_zero_pivot_shift(P, oracle):
    """MatPivotCheck_nz: a structurally present but numerically zero pivot is repaired by the NONZERO shift."""
    ai = np.array([0, 2, 4], np.int32); aj = np.array([0, 1, 0, 1], np.int32); aa = np.array([1.0, 1.0, 1.0, 1.0])
    A = P.Mat.from_csr(ai, aj, aa)
    pc = P.PC.create(P.COMM_SELF, "ilu"); pc.set_operators(A); pc.setup()
    x = P.Vec.from_array([1.0, 2.0]); y = x.duplicate()
    pc.apply(x, y)
    bi, bj, bd, ba = oracle.ilu0(ai, aj, aa)
    assert np.array_equal(y.array(), oracle.matsolve(bi, bj, bd, ba, np.array([1.0, 2.0])))
    for o in (pc, x, y, A): o.destroy()
    ai = np.array([0, 1, 3], np.int32); aj = np.array([1, 0, 1], np.int32)
    A = P.Mat.from_csr(ai, aj, np.ones(3))
    pc = P.PC.create(P.COMM_SELF, "ilu"); pc.set_operators(A)
    with pytest.raises(P.PetscError) as e:
        pc.setup()
    assert "missing diagonal" in str(e.value).lower()
    pc.destroy(); A.destroy()


def test_ex2_1_golden(P, oracle):
    """src/ksp/ksp/tutorials/output/ex2_1.out: -m 5 -n 5 -ksp_monitor -ksp_gmres_cgs_refinement_type refine_always (default PC = ILU)."""
    ai, aj, aa = oracle.lap5(5, 5)
    r = solve(P, ai, aj, aa, "-ksp_gmres_cgs_refinement_type refine_always -ksp_rtol %r" % (1e-2 / 36))
    assert [sig6(v) for v in r["hist"]] == [3.21109, 0.93268, 0.103515, 0.00787798, 0.000387275]
    assert r["its"] == 4 and r["reason"] == 2
    assert sig6(np.linalg.norm(r["x"] - 1.0)) == 0.000392701


@pytest.mark.parametrize("path", KSPF, ids=[os.path.basename(p)[:-4] for p in KSPF])
def test_ksp_history_vs_reference(P, oracle, path):
    g = np.load(path)
    ai, aj, aa = load_matrix(oracle, g)
    r = solve(P, ai, aj, aa, " ".join(str(s) for s in g["opts"]))
    ref = g["ref_hist"]
    assert r["reason"] == int(g["ref_reason"])
    assert abs(r["its"] - int(g["ref_its"])) <= 1
    k = min(31, len(ref), len(r["hist"]))
    # first restart cycle: 1e-12 relative to the initial residual (north_star tolerance); later cycles amplify rounding
    assert_history_1e12(r["hist"], ref, k, os.path.basename(path))
    m = min(len(ref), len(r["hist"]))
    assert np.allclose(r["hist"][:m], ref[:m], rtol=1e-5, atol=1e-12 * ref[0])
    assert np.allclose(r["x"], g["ref_sol"], rtol=0, atol=1e-6 * max(1.0, float(np.abs(g["ref_sol"]).max())))


def test_jacobi_fused_equals_unfused(P, oracle):
    ai, aj, aa = oracle.lap7(20, 18, 16)
    a = solve(P, ai, aj, aa, "-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-9 -pc_jacobi_b200_fuse 1 -vec_b200_fuse_maxpy_norm 1")
    b = solve(P, ai, aj, aa, "-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-9 -pc_jacobi_b200_fuse 0 -vec_b200_fuse_maxpy_norm 1")
    assert a["its"] == b["its"] and np.array_equal(a["hist"], b["hist"]) and np.array_equal(a["x"], b["x"])
    c = solve(P, ai, aj, aa, "-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-9 -vec_b200_fuse_maxpy_norm 0")
    assert abs(a["its"] - c["its"]) <= 1 and np.allclose(a["hist"][:30], c["hist"][:30], rtol=1e-9)
    o_x, o = oracle.ksp_solve("gmres", ai, aj, aa, oracle.matmult(ai, aj, aa, np.ones(len(ai) - 1)), pc="jacobi", rtol=1e-9)
    assert abs(a["its"] - o["its"]) <= 1 and np.allclose(a["hist"][:30], o["hist"][:30], rtol=1e-9)


def test_cg_ilu_27pt_config3_shape(P, oracle):
    """BASELINE config 3 at test size: 27-point operator, KSPCG + PCILU(0)."""
    ai, aj, aa = oracle.lap27(14)
    b = oracle.matmult(ai, aj, aa, np.ones(len(ai) - 1))
    r = solve(P, ai, aj, aa, "-ksp_type cg -pc_type ilu -ksp_rtol 1e-10")
    ox, o = oracle.ksp_solve("cg", ai, aj, aa, b, pc="ilu", rtol=1e-10)
    assert r["its"] == o["its"] and r["reason"] == o["reason"]
    assert np.allclose(r["hist"], o["hist"], rtol=1e-8, atol=1e-13 * o["hist"][0])
    assert np.allclose(r["x"], 1.0, atol=1e-8)


def test_gmres_restart_and_maxit(P, oracle):
    ai, aj, aa = oracle.lap5(40, 40)
    r = solve(P, ai, aj, aa, "-ksp_type gmres -pc_type none -ksp_gmres_restart 7 -ksp_max_it 23 -ksp_rtol 1e-14")
    assert r["its"] == 23 and r["reason"] == -3  # KSP_DIVERGED_ITS
    ox, o = oracle.ksp_solve("gmres", ai, aj, aa, oracle.matmult(ai, aj, aa, np.ones(1600)), pc="none", restart=7, max_it=23, rtol=1e-14)
    assert o["its"] == 23 and o["reason"] == -3
    assert np.allclose(r["hist"], o["hist"], rtol=1e-9)
    assert len(r["hist"]) == len(o["hist"])


def test_vec_user_array_is_the_host_storage(P):
    """VecCreateSeqWithArray / VecPlaceArray / VecResetArray (bvec2.c, rvector.c:2593): the caller's host array is the vector's
    host storage -- device results land in it on the next host access, host edits reach the device on the next device op."""
    petsc = P
    a = np.arange(12.0)
    v = petsc.Vec.with_array(a, 12)
    w = v.duplicate(); w.set(2.0)
    v.axpy(1.0, w)                                   # device op: uploads a, computes a + 2
    assert np.array_equal(a, np.arange(12.0))        # nothing came back yet
    assert np.array_equal(v.array(), np.arange(12.0) + 2.0)
    assert np.array_equal(a, np.arange(12.0) + 2.0)  # the host access copied into the user's array itself
    a[:] = 5.0; v.touch_host()                       # the user changes the array and says so (VecGetArrayWrite/Restore)
    assert v.norm() == pytest.approx(5.0 * np.sqrt(12.0), rel=1e-14)
    # place / reset on a vector that owns its storage
    u = w.duplicate(); u.set(1.0)
    b = np.full(12, 7.0)
    u.place_array(b)
    assert np.array_equal(u.array(), b)
    u.scale(2.0)
    assert np.array_equal(u.array(), np.full(12, 14.0)) and np.array_equal(b, np.full(12, 14.0))
    with pytest.raises(petsc.PetscError):
        u.place_array(b)                             # second VecPlaceArray without VecResetArray (PETSC_ERR_ARG_WRONGSTATE)
    u.reset_array()
    u.set(3.0)
    assert np.array_equal(u.array(), np.full(12, 3.0)) and np.array_equal(b, np.full(12, 14.0))
    for o in (u, v, w):
        o.destroy()


def test_pipecg_single_reduction_matches_cg(P, oracle):
    """KSPPIPECG (one fused reduction + one host synchronisation per iteration) against KSPCG and the oracle's restatement of
    pipecg.c on a 3-D problem: same iteration count to +-2, same solution, history equal to the oracle's over the first 30
    iterations."""
    ai, aj, aa = oracle.lap7(24, 20, 16)
    n = len(ai) - 1
    cg = solve(P, ai, aj, aa, "-ksp_type cg -pc_type jacobi -ksp_rtol 1e-9")
    pc = solve(P, ai, aj, aa, "-ksp_type pipecg -pc_type jacobi -ksp_rtol 1e-9")
    assert pc["reason"] == 2 and abs(pc["its"] - cg["its"]) <= 2
    assert np.allclose(pc["x"], cg["x"], rtol=0, atol=1e-7)
    _, o = oracle.ksp_solve("pipecg", ai, aj, aa, oracle.matmult(ai, aj, aa, np.ones(n)), pc="jacobi", rtol=1e-9)
    k = min(30, len(o["hist"]), len(pc["hist"]))
    assert abs(o["its"] - pc["its"]) <= 1 and np.allclose(pc["hist"][:k], o["hist"][:k], rtol=1e-9, atol=1e-12 * o["hist"][0])
    # the one-kernel form of the eight vector recurrences (default) is the same arithmetic as the eight separate kernels
    un = solve(P, ai, aj, aa, "-ksp_type pipecg -pc_type jacobi -ksp_rtol 1e-9 -ksp_pipecg_b200_fuse_update 0")
    assert un["its"] == pc["its"] and np.array_equal(un["hist"], pc["hist"]) and np.array_equal(un["x"], pc["x"])
    # an odd length exercises the scalar tail of the vectorised kernel
    ai2, aj2, aa2 = oracle.lap5(13, 7)
    a = solve(P, ai2, aj2, aa2, "-ksp_type pipecg -pc_type jacobi -ksp_rtol 1e-10")
    b = solve(P, ai2, aj2, aa2, "-ksp_type pipecg -pc_type jacobi -ksp_rtol 1e-10 -ksp_pipecg_b200_fuse_update 0")
    assert a["reason"] == 2 and a["its"] == b["its"] and np.array_equal(a["hist"], b["hist"]) and np.array_equal(a["x"], b["x"])
