"""CPU-side checks of the C-ABI libraries: they load, export every symbol the headers declare, fail loudly without a GPU,
and the pure host logic (options database, MPIAIJ column split, halo plan) matches the oracle.  No compute calls."""
import ctypes as C
import os

import numpy as np
import pytest

from petsc_b200 import _capi
from harness import petsc


def test_kernel_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = _capi.exported_symbols()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.b200Version()
    # and the other direction: the shared object exports no b200* entry point that include/petscb200.h does not declare
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("b200")}
    assert exported == set(names), sorted(exported ^ set(names))


def test_host_library_exports_every_declared_symbol():
    L = petsc.lib()
    names = petsc.host_symbols()
    assert len(names) > 100
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    n = C.c_int(0)
    rc = _capi.lib().b200DeviceCount(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(petsc.PetscError) as e:
        petsc.initialize()
    assert e.value.code in (97, 96)  # PETSC_ERR_GPU
    h = C.c_void_p()
    assert _capi.lib().b200Create(C.byref(h), 0) == 97
    assert b"cuda error" in _capi.lib().b200GetLastErrorString()


def test_package_does_not_import_oracle():
    import subprocess
    import sys
    out = subprocess.check_output([sys.executable, "-c", "import sys; import petsc_b200, petsc_b200._capi, harness.petsc; print([m for m in sys.modules if 'oracle' in m])"],
                                  cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.strip() == b"[]"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert "oracle" not in open(os.path.join(root, "petsc_plugin", "petscb200_plugin.c")).read().lower()   # the PETSc binding of the product
    for dirpath, _, files in list(os.walk(os.path.join(root, "petsc_b200"))) + list(os.walk(os.path.join(root, "include"))):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f), errors="replace").read().lower(), f


def test_options_database():
    L = petsc.lib()
    petsc.options_clear()
    petsc.options_insert("-ksp_type gmres -ksp_rtol 1e-7 -ksp_monitor -shift -1.5 -sub_pc_type ilu")
    v = C.c_double(0); s = C.c_int(0); buf = C.create_string_buffer(64); b = C.c_int(0)
    petsc.chk(L.PetscOptionsGetReal(None, None, b"-ksp_rtol", C.byref(v), C.byref(s))); assert s.value and v.value == 1e-7
    petsc.chk(L.PetscOptionsGetReal(None, None, b"-shift", C.byref(v), C.byref(s))); assert s.value and v.value == -1.5
    petsc.chk(L.PetscOptionsGetString(None, None, b"-ksp_type", buf, 64, C.byref(s))); assert buf.value == b"gmres"
    petsc.chk(L.PetscOptionsGetString(None, b"sub_", b"-pc_type", buf, 64, C.byref(s))); assert buf.value == b"ilu"
    petsc.chk(L.PetscOptionsGetBool(None, None, b"-ksp_monitor", C.byref(b), C.byref(s))); assert s.value and b.value
    petsc.chk(L.PetscOptionsGetBool(None, None, b"-nope", C.byref(b), C.byref(s))); assert not s.value
    petsc.options_set("-ksp_type", "cg")
    petsc.chk(L.PetscOptionsGetString(None, None, b"-ksp_type", buf, 64, C.byref(s))); assert buf.value == b"cg"
    assert L.PetscOptionsSetValue(None, b"bad", b"1") == 62  # PETSC_ERR_ARG_WRONG
    assert b"must start with" in L.PetscB200GetLastErrorMessage()
    petsc.options_clear()


def _split(L, ai, ajg, aa, cstart, cend):
    m = len(ai) - 1
    P = C.POINTER
    Ai, Aj, Bi, Bj, g = P(C.c_int)(), P(C.c_int)(), P(C.c_int)(), P(C.c_int)(), P(C.c_int)()
    Aa, Ba = P(C.c_double)(), P(C.c_double)()
    ec = C.c_int()
    petsc.chk(L.PetscB200MPIAIJSplit(m, int(cstart), int(cend), ai.ctypes.data_as(C.c_void_p), ajg.ctypes.data_as(C.c_void_p), aa.ctypes.data_as(C.c_void_p),
                                     C.byref(Ai), C.byref(Aj), C.byref(Aa), C.byref(Bi), C.byref(Bj), C.byref(Ba), C.byref(g), C.byref(ec)))
    as_np = np.ctypeslib.as_array
    ai_ = as_np(Ai, (m + 1,)).copy(); bi_ = as_np(Bi, (m + 1,)).copy()
    A = (ai_, as_np(Aj, (max(ai_[-1], 1),))[:ai_[-1]].copy(), as_np(Aa, (max(ai_[-1], 1),))[:ai_[-1]].copy())
    B = (bi_, as_np(Bj, (max(bi_[-1], 1),))[:bi_[-1]].copy(), as_np(Ba, (max(bi_[-1], 1),))[:bi_[-1]].copy())
    return A, B, as_np(g, (max(ec.value, 1),))[:ec.value].copy()


@pytest.mark.parametrize("size", [2, 3, 5])
def test_mpiaij_split_and_halo_plan_match_oracle(oracle, size):
    """MatSetUpMultiply_MPIAIJ restatement in the host library (index work: bit-exact) vs the oracle's, rank by rank."""
    L = petsc.lib()
    for gen in (lambda: oracle.lap5(9, 7), lambda: oracle.lap7(6, 5, 7), lambda: oracle.random_csr(211, 6, 5)):
        ai, aj, aa = gen()
        n = len(ai) - 1
        rs = oracle.split_ownership(n, size)
        ranges = np.ascontiguousarray(rs, dtype=np.int64)
        for r in range(size):
            r0, r1 = int(rs[r]), int(rs[r + 1])
            lai = (ai[r0:r1 + 1] - ai[r0]).astype(np.int32)
            laj = np.ascontiguousarray(aj[ai[r0]:ai[r1]]); laa = np.ascontiguousarray(aa[ai[r0]:ai[r1]])
            A, B, g = _split(L, lai, laj, laa, r0, r1)
            oA, oB, og = oracle.mpiaij_split(lai, laj.astype(np.int64), laa, r0, r1)
            for x, y in zip(A + B, oA + oB):
                assert np.array_equal(x, y)
            assert np.array_equal(g, og) and np.all(np.diff(g) > 0)
            rc = np.zeros(size, np.int32); ro = np.zeros(size, np.int32)
            petsc.chk(L.PetscB200HaloPlanRecv(len(g), g.astype(np.int32).ctypes.data_as(C.c_void_p), size, ranges.ctypes.data_as(C.c_void_p),
                                              rc.ctypes.data_as(C.c_void_p), ro.ctypes.data_as(C.c_void_p)))
            assert rc[r] == 0 and rc.sum() == len(g)
            for p in range(size):
                seg = g[ro[p]:ro[p] + rc[p]]
                assert np.all((seg >= rs[p]) & (seg < rs[p + 1]))


def test_mpiaij_setup_helpers_index_exact_on_cpu(oracle):
    """The host half of MatSetUpMultiply_MPIAIJ that the PETSc plugin and the harness share (b200MpiaijSplitHost,
    b200MpiaijBuildGarray: plain C inside libpetscb200.so, no device needed) against the oracle's restatement of mmaij.c:25-61:
    diagonal / off-diagonal blocks, garray and the renumbered off-diagonal columns, for 1, 2, 3 and 5 ranks."""
    L = _capi.lib()
    for name, (ai, aj, aa) in (("lap5", oracle.lap5(13, 9)), ("lap7", oracle.lap7(7, 6, 9)), ("rand", oracle.random_csr(203, 7, 3))):
        n = len(ai) - 1
        for size in (1, 2, 3, 5):
            rs = oracle.split_ownership(n, size)
            for rank in range(size):
                r0, r1 = int(rs[rank]), int(rs[rank + 1])
                m = r1 - r0
                lai = np.ascontiguousarray(ai[r0:r1 + 1] - ai[r0], np.int32)
                laj = np.ascontiguousarray(aj[ai[r0]:ai[r1]], np.int32); laa = np.ascontiguousarray(aa[ai[r0]:ai[r1]])
                nzA, nzB = C.c_int64(), C.c_int64()
                vp = C.c_void_p
                p = lambda a: a.ctypes.data_as(vp)   # noqa: E731
                _capi.check(L.b200MpiaijSplitHost(m, r0, r1, p(lai), p(laj), p(laa), C.byref(nzA), C.byref(nzB), None, None, None, None, None, None))
                Ai, Bi = np.zeros(m + 1, np.int32), np.zeros(m + 1, np.int32)
                Aj, Bj = np.zeros(nzA.value + 1, np.int32), np.zeros(nzB.value + 1, np.int32)
                Aa, Ba = np.zeros(nzA.value + 1), np.zeros(nzB.value + 1)
                _capi.check(L.b200MpiaijSplitHost(m, r0, r1, p(lai), p(laj), p(laa), C.byref(nzA), C.byref(nzB), p(Ai), p(Aj), p(Aa), p(Bi), p(Bj), p(Ba)))
                g, ec = vp(), C.c_int()
                _capi.check(L.b200MpiaijBuildGarray(C.c_int64(nzB.value), p(Bj), C.byref(g), C.byref(ec)))
                garray = np.ctypeslib.as_array(C.cast(g, C.POINTER(C.c_int)), shape=(max(ec.value, 1),))[:ec.value].copy()
                _capi.check(L.b200HostFree(g))
                oA, oB, og = oracle.mpiaij_split(lai, laj.astype(np.int64), laa, r0, r1)
                assert np.array_equal(garray, og), (name, size, rank)
                assert np.array_equal(Ai, oA[0]) and np.array_equal(Aj[:nzA.value], oA[1]) and np.array_equal(Aa[:nzA.value], oA[2]), (name, size, rank)
                assert np.array_equal(Bi, oB[0]) and np.array_equal(Bj[:nzB.value], oB[1]) and np.array_equal(Ba[:nzB.value], oB[2]), (name, size, rank)


def test_indexed_plan_host_grouping_is_stable(oracle):
    """The host half of b200IndexedPlanCreate (b200IndexedGroupHost, plain C, no device): destinations grouped in increasing order,
    the entries of a group in ENTRY order -- replaying the groups reproduces the oracle's sequential PetscSFLinkScatterLocal."""
    L = _capi.lib()
    vp = C.c_void_p
    p = lambda a: None if a is None else a.ctypes.data_as(vp)   # noqa: E731
    rng = np.random.default_rng(5)

    def group(sidx, didx, n):
        sc, dc, grouped = C.c_int(), C.c_int(), C.c_int()
        se, de, ng = C.c_int64(), C.c_int64(), C.c_int64()
        gd, go, gs = vp(), vp(), vp()
        _capi.check(L.b200IndexedGroupHost(C.c_int64(n), p(sidx), 0, p(didx), 0, C.byref(sc), C.byref(dc), C.byref(se), C.byref(de), C.byref(grouped), C.byref(ng),
                                           C.byref(gd), C.byref(go), C.byref(gs)))
        out = dict(sc=sc.value, dc=dc.value, se=se.value, de=de.value, grouped=grouped.value, ng=ng.value)
        if grouped.value:
            arr = lambda q, k: np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_int)), shape=(k,)).copy()   # noqa: E731
            out.update(gdst=arr(gd, ng.value), goff=arr(go, ng.value + 1), gsrc=arr(gs, n))
            for q in (gd, go, gs):
                _capi.check(L.b200HostFree(q))
        else:
            assert not gd.value and not go.value and not gs.value
        return out

    n = 5000
    sidx = rng.integers(0, 700, n).astype(np.int32); didx = rng.integers(0, 300, n).astype(np.int32)
    g = group(sidx, didx, n)
    assert g["grouped"] and g["ng"] == len(np.unique(didx)) and g["se"] == sidx.max() + 1 and g["de"] == didx.max() + 1 and not g["sc"] and not g["dc"]
    order = np.argsort(didx, kind="stable")
    assert np.array_equal(g["gdst"], np.unique(didx)) and np.array_equal(g["gsrc"], sidx[order]) and g["goff"][-1] == n
    assert np.array_equal(np.diff(g["goff"]), np.bincount(didx)[np.unique(didx)])
    # replay group by group == the sequential reference loop, bit for bit (sums are order dependent)
    src, dst = rng.uniform(-1, 1, 700), rng.uniform(-1, 1, 300)
    want = oracle.sf_scatter(sidx, didx, src, dst, "sum")
    got = dst.copy()
    for k in range(g["ng"]):
        v = got[g["gdst"][k]]
        for j in range(g["goff"][k], g["goff"][k + 1]):
            v = v + src[g["gsrc"][j]]
        got[g["gdst"][k]] = v
    assert np.array_equal(got, want)
    # distinct destinations: no grouping; contiguous sides detected; NULL index arrays = contiguous
    perm = rng.permutation(400).astype(np.int32)
    g = group(perm, np.arange(7, 407, dtype=np.int32), 400)
    assert not g["grouped"] and g["dc"] and not g["sc"] and g["de"] == 407
    g = group(None, perm, 400)
    assert not g["grouped"] and g["sc"] and not g["dc"] and g["ng"] == 400
    g = group(None, None, 0)
    assert not g["grouped"] and g["se"] == 0 and g["de"] == 0
    sc = C.c_int()
    assert L.b200IndexedGroupHost(C.c_int64(2), p(np.array([0, -3], np.int32)), 0, None, 0, C.byref(sc), C.byref(sc), C.byref(C.c_int64()), C.byref(C.c_int64()), C.byref(sc),
                                  C.byref(C.c_int64()), C.byref(vp()), C.byref(vp()), C.byref(vp())) == 63
