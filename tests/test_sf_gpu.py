"""Indexed scatter-with-op on device data (b200IndexedOp = the local part of PetscSFBcast / PetscSFReduce) through the
C ABI: bit-identical to the reference's own results (tests/golden/sf_*.npz, made by oracle/ref_driver.c -sf through PetscSF) and to
the oracle restatement of PetscSFLinkScatterLocal on larger seeded cases -- every op, float64 and int32, block sizes, contiguous
sides, repeated destinations (grouped kernel, order-dependent sums)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
from conftest import SF_OPS, golden_path, sf_graph_order

pytestmark = pytest.mark.gpu
OPS = {"replace": 0, "sum": 1, "prod": 2, "max": 3, "min": 4}


@pytest.fixture(scope="module")
def H():
    from petsc_b200 import _capi
    h = _capi.Handle()
    yield h
    h.close()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def indexed_op(H, sidx, didx, src, dst, op, bs=1, n=None):
    """dst after dst[didx[i]] op= src[sidx[i]] (i in order) on the device; None index = contiguous from 0."""
    from petsc_b200 import _capi
    L = _capi.lib()
    n = len(sidx) if sidx is not None else (len(didx) if didx is not None else n)
    si = None if sidx is None else np.ascontiguousarray(sidx, np.int32)
    di = None if didx is None else np.ascontiguousarray(didx, np.int32)
    plan = C.c_void_p()
    _capi.check(L.b200IndexedPlanCreate(H.h, C.c_int64(n), _ptr(si), 0, _ptr(di), 0, C.byref(plan)))
    isint = np.asarray(src).dtype == np.int32
    dt = np.int32 if isint else np.float64
    d_s = _capi.DeviceArray(H, max(len(src), 1), dt); d_d = _capi.DeviceArray(H, max(len(dst), 1), dt)
    if len(src):
        d_s.upload(np.concatenate([src, np.zeros(max(len(src), 1) - len(src), dt)]))
    if len(dst):
        d_d.upload(np.concatenate([dst, np.zeros(max(len(dst), 1) - len(dst), dt)]))
    _capi.check(L.b200IndexedOp(H.h, plan, 1 if isint else 0, int(bs), OPS[op], d_s.ptr, d_d.ptr))
    out = d_d.download()[:len(dst)]
    info = [C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int(), C.c_int64(), C.c_int64()]
    _capi.check(L.b200IndexedPlanGetInfo(plan, *[C.byref(v) for v in info]))
    _capi.check(L.b200IndexedPlanDestroy(H.h, plan))
    d_s.free(); d_d.free()
    return out, dict(n=info[0].value, ngroups=info[1].value, grouped=info[2].value, src_contig=info[3].value, dst_contig=info[4].value)


SF = sorted(glob.glob(golden_path("sf_*.npz")))


@pytest.mark.parametrize("path", SF, ids=[os.path.basename(p)[:-4] for p in SF])
def test_indexed_op_bit_exact_vs_reference_petscsf(H, path):
    g = np.load(path)
    leafloc, rootidx = sf_graph_order(g)
    bs = int(g["bs"])
    for op in SF_OPS:
        out, info = indexed_op(H, rootidx, leafloc, g["root"], g["leaf"], op, bs)          # broadcast: leaf op= root
        assert np.array_equal(out, g["ref_bcast_" + op]), ("bcast", op)
        assert not info["grouped"]                                                         # leaves are distinct
        out, info = indexed_op(H, leafloc, rootidx, g["leaf"], g["root"], op, bs)          # reduction: root op= leaf, roots repeat
        assert np.array_equal(out, g["ref_reduce_" + op]), ("reduce", op)
    for op in ("sum", "max"):
        out, _ = indexed_op(H, rootidx, leafloc, g["rooti"], g["leafi"], op)
        assert np.array_equal(out, g["ref_bcast_%s_i32" % op]), op
        out, _ = indexed_op(H, leafloc, rootidx, g["leafi"], g["rooti"], op)
        assert np.array_equal(out, g["ref_reduce_%s_i32" % op]), op


def test_indexed_op_bit_exact_vs_oracle_large(H, oracle):
    rng = np.random.default_rng(77)
    nsrc, ndst, n = 300_000, 120_000, 500_000          # ~4 entries per destination: long ordered sums
    for bs in (1, 2, 5):
        sidx = rng.integers(0, nsrc, n).astype(np.int32); didx = rng.integers(0, ndst, n).astype(np.int32)
        src = rng.uniform(-2, 2, nsrc * bs); dst = rng.uniform(-2, 2, ndst * bs)
        for op in SF_OPS:
            out, info = indexed_op(H, sidx, didx, src, dst, op, bs)
            assert info["grouped"] and info["ngroups"] == len(np.unique(didx))
            assert np.array_equal(out, oracle.sf_scatter(sidx, didx, src, dst, op, bs)), (bs, op)
    # distinct destinations (a permutation), contiguous source / contiguous destination / both
    perm = rng.permutation(ndst).astype(np.int32)
    src = rng.uniform(-1, 1, ndst * 3); dst = rng.uniform(-1, 1, ndst * 3)
    for sidx, didx in ((perm, None), (None, perm), (None, None), (perm[::-1].copy(), perm)):
        for op in ("replace", "sum", "min"):
            out, info = indexed_op(H, sidx, didx, src, dst, op, 3, n=ndst)
            assert not info["grouped"] and info["src_contig"] == (sidx is None) and info["dst_contig"] == (didx is None)
            assert np.array_equal(out, oracle.sf_scatter(sidx if sidx is not None else np.arange(ndst, dtype=np.int32), didx, src, dst, op, 3)), op
    # int32, wrap-around products included
    si = rng.integers(0, 5000, 40000).astype(np.int32); di = rng.integers(0, 900, 40000).astype(np.int32)
    s32 = rng.integers(-2**31, 2**31 - 1, 5000).astype(np.int32); d32 = rng.integers(-2**31, 2**31 - 1, 900).astype(np.int32)
    for op in SF_OPS:
        out, _ = indexed_op(H, si, di, s32, d32, op)
        assert np.array_equal(out, oracle.sf_scatter(si, di, s32, d32, op)), op


def test_indexed_op_errors_and_empty(H):
    from petsc_b200 import _capi
    L = _capi.lib()
    out, info = indexed_op(H, np.zeros(0, np.int32), np.zeros(0, np.int32), np.ones(3), np.ones(4), "sum")
    assert np.array_equal(out, np.ones(4)) and info["n"] == 0
    plan = C.c_void_p()
    bad = np.array([0, -1], np.int32)
    assert L.b200IndexedPlanCreate(H.h, C.c_int64(2), _ptr(bad), 0, None, 0, C.byref(plan)) == 63      # PETSC_ERR_ARG_OUTOFRANGE
    idx = np.array([1, 0], np.int32)
    _capi.check(L.b200IndexedPlanCreate(H.h, C.c_int64(2), _ptr(idx), 0, None, 0, C.byref(plan)))
    d = _capi.DeviceArray(H, 2, np.float64).upload(np.ones(2))
    assert L.b200IndexedOp(H.h, plan, 0, 1, 0, d.ptr, d.ptr) == 56                                        # in place: PETSC_ERR_SUP
    assert L.b200IndexedOp(H.h, plan, 0, 1, 9, d.ptr, d.ptr) != 0
    _capi.check(L.b200IndexedPlanDestroy(H.h, plan)); d.free()
