"""The three ILU(0) sweep implementations behind b200Ilu0Solve -- level-scheduled pipe kernels (default for narrow levels), packed
slot-space sweeps (default for wide levels, i.e. the 7-point operator at benchmark sizes) and the segment-marching kernel --
must all reproduce MatSolve_SeqAIJ_NaturalOrdering bit for bit.  The default choice depends on the level width, which small test
matrices never reach, so the variants are forced through their environment switches (read when the plan is made / solved)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS100 = 100 * 2.220446049250313e-16


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def H():
    from petsc_b200 import _capi
    h = _capi.Handle()
    yield h
    h.close()


VARIANTS = {"pipe": {"PETSCB200_ILU_PACKED": "0", "PETSCB200_ILU_MARCH": "0"},
            "packed": {"PETSCB200_ILU_PACKED": "1", "PETSCB200_ILU_PACKED_MIN_WIDTH": "0", "PETSCB200_ILU_MARCH": "0"}}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_ilu0_sweep_variants_bit_exact(H, oracle, variant):
    from petsc_b200 import _capi
    L = _capi.lib()
    old = {k: os.environ.get(k) for k in ("PETSCB200_ILU_PACKED", "PETSCB200_ILU_PACKED_MIN_WIDTH", "PETSCB200_ILU_MARCH")}
    os.environ.update(VARIANTS[variant])
    try:
        rng = np.random.default_rng(11)
        for name, (ai, aj, aa) in (("lap5", oracle.lap5(23, 19)), ("lap7", oracle.lap7(17, 13, 11)), ("lap27", oracle.lap27(11)), ("lap7_big", oracle.lap7(48, 40, 36))):
            n, nnz = len(ai) - 1, len(aj)
            plan = C.c_void_p()
            _capi.check(L.b200Ilu0Symbolic(H.h, n, _ptr(ai), _ptr(aj), C.byref(plan)))
            d_a = _capi.DeviceArray(H, nnz, np.float64).upload(aa)
            ns = C.c_int(-1)
            _capi.check(L.b200Ilu0Numeric(H.h, plan, d_a.ptr, C.c_double(EPS100), C.c_double(EPS100), C.byref(ns)))
            bi, bj, bd, ba = oracle.ilu0(ai, aj, aa)
            for _ in range(2):
                b = rng.uniform(-1, 1, n)
                d_b = _capi.DeviceArray(H, n, np.float64).upload(b); d_x = _capi.DeviceArray(H, n, np.float64)
                _capi.check(L.b200Ilu0Solve(H.h, plan, d_b.ptr, d_x.ptr))
                assert np.array_equal(d_x.download(), oracle.matsolve(bi, bj, bd, ba, b)), (variant, name)
                d_b.free(); d_x.free()
            _capi.check(L.b200Ilu0Destroy(plan)); d_a.free()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_packed_equals_pipe_at_benchmark_width(H):
    """7-point 320^3 (32.8 M rows, 34 k rows per dependency level): wide enough that the plan picks the packed slot-space sweeps by
    itself.  The same factor solved by the level-scheduled pipe kernels must give the same bits; and M^-1 (4b) = 4 M^-1 b exactly
    (power-of-two scaling commutes with every FMA-free operation)."""
    from petsc_b200 import _capi
    L = _capi.lib()
    n1 = 320
    N = n1 ** 3
    nnz = C.c_int64()
    _capi.check(L.b200GenLaplace7Nnz(n1, n1, n1, C.c_int64(0), C.c_int64(N), C.byref(nnz)))
    nnz = nnz.value
    d_i, d_j, d_a = _capi.DeviceArray(H, N + 1, np.int32), _capi.DeviceArray(H, nnz, np.int32), _capi.DeviceArray(H, nnz, np.float64)
    _capi.check(L.b200GenLaplace7(H.h, n1, n1, n1, C.c_int64(0), C.c_int64(N), d_i.ptr, d_j.ptr, d_a.ptr))
    ai, aj = d_i.download(), d_j.download()
    b = np.random.default_rng(5).uniform(-1, 1, N)
    d_b = _capi.DeviceArray(H, N, np.float64).upload(b); d_b4 = _capi.DeviceArray(H, N, np.float64).upload(4.0 * b)
    res = {}
    old = os.environ.get("PETSCB200_ILU_PACKED_MIN_WIDTH")
    try:
        for variant, minw in (("packed", None), ("pipe", "1e30")):
            if minw is None:
                os.environ.pop("PETSCB200_ILU_PACKED_MIN_WIDTH", None)
            else:
                os.environ["PETSCB200_ILU_PACKED_MIN_WIDTH"] = minw
            plan = C.c_void_p()
            _capi.check(L.b200Ilu0Symbolic(H.h, N, _ptr(ai), _ptr(aj), C.byref(plan)))
            ns = C.c_int(-1)
            _capi.check(L.b200Ilu0Numeric(H.h, plan, d_a.ptr, C.c_double(EPS100), C.c_double(EPS100), C.byref(ns)))
            d_x = _capi.DeviceArray(H, N, np.float64); d_x4 = _capi.DeviceArray(H, N, np.float64)
            _capi.check(L.b200Ilu0Solve(H.h, plan, d_b.ptr, d_x.ptr))
            _capi.check(L.b200Ilu0Solve(H.h, plan, d_b4.ptr, d_x4.ptr))
            res[variant] = d_x.download()
            assert np.array_equal(d_x4.download(), 4.0 * res[variant]), variant
            _capi.check(L.b200Ilu0Destroy(plan)); d_x.free(); d_x4.free()
    finally:
        if old is None:
            os.environ.pop("PETSCB200_ILU_PACKED_MIN_WIDTH", None)
        else:
            os.environ["PETSCB200_ILU_PACKED_MIN_WIDTH"] = old
    assert np.array_equal(res["packed"], res["pipe"])
    assert np.isfinite(res["packed"]).all()
    for o in (d_i, d_j, d_a, d_b, d_b4):
        o.free()
