"""ctypes binding of liboracle.so (the CPU restatement of the reference's Krylov path).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product package ``petsc_b200`` never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.ora_vecdot.restype = C.c_double
        _LIB.ora_vecdot_omp.restype = C.c_double
        _LIB.ora_vecnorm2.restype = C.c_double
        for f in ("ora_lap5_nnz", "ora_lap7_nnz", "ora_lap27_nnz", "ora_lap7_rows_nnz"):
            getattr(_LIB, f).restype = C.c_int64
    return _LIB


class KspOpts(C.Structure):
    _fields_ = [("pc_type", C.c_int), ("restart", C.c_int), ("cgs_refine", C.c_int), ("max_it", C.c_int),
                ("rtol", C.c_double), ("abstol", C.c_double), ("dtol", C.c_double), ("nblocks", C.c_int),
                ("use_omp", C.c_int)]


class KspResult(C.Structure):
    _fields_ = [("its", C.c_int), ("reason", C.c_int), ("rnorm", C.c_double), ("nhist", C.c_int)]


PC = {"none": 0, "jacobi": 1, "ilu": 2, "bjacobi": 3, "icc": 4}
REFINE = {"never": 0, "ifneeded": 1, "always": 2}


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- generators
def lap5(m, n):
    L = lib()
    nnz = L.ora_lap5_nnz(m, n)
    ai = np.empty(m * n + 1, np.int32); aj = np.empty(nnz, np.int32); aa = np.empty(nnz, np.float64)
    L.ora_lap5(m, n, _p(ai), _p(aj), _p(aa))
    return ai, aj, aa


def lap7(nx, ny=None, nz=None, omp=False):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    L = lib()
    nnz = L.ora_lap7_nnz(nx, ny, nz)
    ai = np.empty(nx * ny * nz + 1, np.int32); aj = np.empty(nnz, np.int32); aa = np.empty(nnz, np.float64)
    (L.ora_lap7_omp if omp else L.ora_lap7)(nx, ny, nz, _p(ai), _p(aj), _p(aa))
    return ai, aj, aa


def lap27(n):
    L = lib()
    nnz = L.ora_lap27_nnz(n)
    ai = np.empty(n ** 3 + 1, np.int32); aj = np.empty(nnz, np.int32); aa = np.empty(nnz, np.float64)
    L.ora_lap27(n, _p(ai), _p(aj), _p(aa))
    return ai, aj, aa


def random_csr(n, d, seed, ncols=None):
    """SURVEY 8(d) 'R': fixed row length d, distinct uniform columns sorted, diagonal forced present, values U(-1,1)."""
    ncols = n if ncols is None else ncols
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, ncols, size=(n, d), dtype=np.int64)
    if n == ncols:
        cols[:, 0] = np.arange(n)
    cols.sort(axis=1)
    # make columns distinct within a row by bumping duplicates (keeps sortedness)
    for _ in range(d):
        dup = cols[:, 1:] <= cols[:, :-1]
        if not dup.any():
            break
        cols[:, 1:] = np.where(dup, cols[:, :-1] + 1, cols[:, 1:])
    cols = np.minimum(cols, ncols - 1)
    # after clamping there can be duplicates at the far right edge; fix by rebuilding those few rows
    bad = np.where((cols[:, 1:] <= cols[:, :-1]).any(axis=1))[0]
    for r in bad:
        c = set([r] if n == ncols else [])
        while len(c) < d:
            c.add(int(rng.integers(0, ncols)))
        cols[r] = np.sort(np.fromiter(c, dtype=np.int64))
    ai = (np.arange(n + 1, dtype=np.int64) * d).astype(np.int32)
    aj = cols.reshape(-1).astype(np.int32)
    aa = rng.uniform(-1.0, 1.0, size=n * d)
    return ai, aj, aa


# ---------------------------------------------------------------- Mat
def matmult(ai, aj, aa, x, omp=False):
    m = len(ai) - 1
    y = np.empty(m, np.float64)
    f = lib().ora_matmult_seqaij_omp if omp else lib().ora_matmult_seqaij
    f(m, _p(ai), _p(aj), _p(aa), _p(x), _p(y))
    return y


def matmultadd(ai, aj, aa, x, y):
    m = len(ai) - 1
    z = np.empty(m, np.float64)
    lib().ora_matmultadd_seqaij(m, _p(ai), _p(aj), _p(aa), _p(x), _p(y), _p(z))
    return z


def getdiagonal(ai, aj, aa):
    m = len(ai) - 1
    d = np.empty(m, np.float64); pos = np.empty(m, np.int32)
    lib().ora_getdiagonal_seqaij(m, _p(ai), _p(aj), _p(aa), _p(d), _p(pos))
    return d, pos


# ---------------------------------------------------------------- Vec
def vecdot(x, y):
    return lib().ora_vecdot(C.c_int64(len(x)), _p(x), _p(y))


def vecnorm2(x):
    return lib().ora_vecnorm2(C.c_int64(len(x)), _p(x))


def _ptrs(ys):
    arr = (C.c_void_p * len(ys))(*[y.ctypes.data for y in ys])
    return arr


def vecmdot(x, ys, omp=False):
    z = np.empty(len(ys), np.float64)
    f = lib().ora_vecmdot_omp if omp else lib().ora_vecmdot
    f(C.c_int64(len(x)), len(ys), _p(x), _ptrs(ys), _p(z))
    return z


def vecmaxpy(x, alpha, ys, omp=False):
    """in place: x += sum alpha_j ys_j"""
    alpha = _f64(alpha)
    f = lib().ora_vecmaxpy_omp if omp else lib().ora_vecmaxpy
    f(C.c_int64(len(x)), len(ys), _p(alpha), _ptrs(ys), _p(x))
    return x


# ---------------------------------------------------------------- ILU(0)
def ilu0(ai, aj, aa, zeropivot=100 * 2.220446049250313e-16, shiftamount=100 * 2.220446049250313e-16):
    n = len(ai) - 1
    nnz = int(ai[n])
    bi = np.empty(n + 1, np.int32); bdiag = np.empty(n + 1, np.int32); bj = np.empty(nnz, np.int32)
    ba = np.zeros(nnz, np.float64)
    rc = lib().ora_ilu0_symbolic(n, _p(ai), _p(aj), _p(bi), _p(bj), _p(bdiag))
    if rc:
        raise ValueError("missing diagonal in row %d" % (-rc - 1))
    ns = lib().ora_lu_numeric(n, _p(ai), _p(aj), _p(aa), _p(bi), _p(bj), _p(bdiag), _p(ba), C.c_double(zeropivot),
                              C.c_double(shiftamount))
    if ns < 0:
        raise ValueError("factorisation failed")
    return bi, bj, bdiag, ba


def matsolve(bi, bj, bdiag, ba, b):
    n = len(bi) - 1
    x = np.empty(n, np.float64)
    lib().ora_matsolve_natural(n, _p(bi), _p(bj), _p(bdiag), _p(ba), _p(b), _p(x))
    return x


# ---------------------------------------------------------------- MPIAIJ
def split_ownership(N, size):
    r = np.empty(size + 1, np.int64)
    lib().ora_split_ownership(C.c_int64(N), size, _p(r))
    return r


def mpiaij_split(ai, aj_global, aa, cstart, cend):
    """ai local row pointer of this rank's rows, aj_global int64 global columns."""
    m = len(ai) - 1
    nnz = int(ai[m])
    ajg = np.ascontiguousarray(aj_global, dtype=np.int64)
    Ai = np.empty(m + 1, np.int32); Bi = np.empty(m + 1, np.int32)
    Aj = np.empty(nnz, np.int32); Bj = np.empty(nnz, np.int32)
    Aa = np.empty(nnz, np.float64); Ba = np.empty(nnz, np.float64)
    g = np.empty(max(nnz, 1), np.int64)
    ec = lib().ora_mpiaij_split(m, C.c_int64(cstart), C.c_int64(cend), _p(ai), _p(ajg), _p(aa), _p(Ai), _p(Aj), _p(Aa),
                                _p(Bi), _p(Bj), _p(Ba), _p(g))
    return (Ai, Aj[:Ai[m]].copy(), Aa[:Ai[m]].copy()), (Bi, Bj[:Bi[m]].copy(), Ba[:Bi[m]].copy()), g[:ec].copy()


# ---------------------------------------------------------------- KSP
def ksp_solve(ksp_type, ai, aj, aa, b, pc="ilu", restart=30, refine="never", max_it=10000, rtol=1e-5, abstol=1e-50,
              dtol=1e4, nblocks=1, omp=False):
    n = len(ai) - 1
    o = KspOpts(PC[pc], restart, REFINE[refine], max_it, rtol, abstol, dtol, nblocks, int(omp))
    res = KspResult()
    cap = max_it + max_it // max(restart, 1) + 8
    hist = np.zeros(cap, np.float64)
    x = np.zeros(n, np.float64)
    f = {"gmres": lib().ora_ksp_gmres, "cg": lib().ora_ksp_cg, "pipecg": lib().ora_ksp_pipecg, "pgmres": lib().ora_ksp_pgmres}[ksp_type]
    rc = f(n, _p(ai), _p(aj), _p(aa), _p(_f64(b)), _p(x), C.byref(o), C.byref(res), _p(hist), cap)
    if rc:
        raise RuntimeError("oracle KSP setup failed rc=%d" % rc)
    return x, dict(its=res.its, reason=res.reason, rnorm=res.rnorm, hist=hist[:min(res.nhist, cap)].copy())


def icc0(ai, aj, aa, zeropivot=100 * 2.220446049250313e-16):
    """MatICCFactorSymbolic_SeqAIJ (levels 0, natural ordering) + MatCholeskyFactorNumeric_SeqAIJ -> (ui, uj, udiag, ua)."""
    ai, aj, aa = _i32(ai), _i32(aj), _f64(aa)
    n = len(ai) - 1
    ui = np.empty(n + 1, np.int32); udiag = np.empty(n, np.int32); uj = np.empty(len(aj) + 1, np.int32); ua = np.zeros(len(aj) + 1)
    rc = lib().ora_icc0_symbolic(n, _p(ai), _p(aj), _p(ui), _p(uj), _p(udiag))
    if rc:
        raise ValueError("matrix is missing the diagonal entry of row %d" % (-rc - 1))
    rc = lib().ora_icc0_numeric(n, _p(ai), _p(aj), _p(aa), _p(ui), _p(uj), _p(udiag), _p(ua), C.c_double(zeropivot))
    if rc:
        raise ArithmeticError("non-positive pivot in row %d (the reference would shift)" % (-rc - 1))
    nz = int(ui[-1])
    return ui, uj[:nz].copy(), udiag, ua[:nz].copy()


def matsolve_icc(ui, uj, udiag, ua, b):
    n = len(ui) - 1
    x = np.empty(n)
    lib().ora_matsolve_icc(n, _p(_i32(ui)), _p(_i32(uj)), _p(_i32(udiag)), _p(_f64(ua)), _p(_f64(b)), _p(x))
    return x


def matmulttranspose(ai, aj, aa, x, n=None, z=None):
    """MatMultTranspose[Add]_SeqAIJ: y = (z or 0) + A^T x."""
    m = len(ai) - 1
    n = m if n is None else n
    x = _f64(x); y = np.empty(n)
    zz = None if z is None else _f64(z)
    lib().ora_matmulttranspose_seqaij(m, n, _p(_i32(ai)), _p(_i32(aj)), _p(_f64(aa)), _p(x), None if zz is None else _p(zz), _p(y))
    return y


def coo_prealloc(M, N, coo_i, coo_j):
    """MatSetPreallocationCOO_SeqAIJ -> (Ai, Aj, jmap, perm) in the reference's own order."""
    ci, cj = _i32(coo_i), _i32(coo_j)
    n = len(ci)
    Ai = np.empty(M + 1, np.int32); Aj = np.empty(max(n, 1), np.int32)
    jmap = np.empty(n + 1, np.int64); perm = np.empty(max(n, 1), np.int64)
    nnz = C.c_int64(); atot = C.c_int64()
    rc = lib().ora_coo_prealloc(M, N, C.c_int64(n), _p(ci), _p(cj), _p(Ai), _p(Aj), _p(jmap), _p(perm), C.byref(nnz), C.byref(atot))
    if rc:
        raise ValueError("COO %s index out of range" % ("row" if rc == 1 else "column"))
    return Ai, Aj[:nnz.value].copy(), jmap[:nnz.value + 1].copy(), perm[:atot.value].copy()


def coo_setvalues(jmap, perm, v, Aa=None):
    """MatSetValuesCOO_SeqAIJ: INSERT_VALUES when Aa is None, else ADD_VALUES onto a copy of Aa."""
    nnz = len(jmap) - 1
    out = np.zeros(nnz) if Aa is None else _f64(Aa).copy()
    jm = np.ascontiguousarray(jmap, dtype=np.int64); pm = np.ascontiguousarray(perm, dtype=np.int64)
    lib().ora_coo_setvalues(C.c_int64(nnz), _p(jm), _p(pm), _p(_f64(v)), 1 if Aa is None else 0, _p(out))
    return out


SF_OPS = {"replace": 0, "sum": 1, "prod": 2, "max": 3, "min": 4}


def sf_scatter(sidx, didx, src, dst, op="replace", bs=1):
    """PetscSFLinkScatterLocal / ScatterAnd<Op> (sfpack.c:190-222, 1082): returns dst after dst[didx[i]] op= src[sidx[i]] in order.
    float64 or int32 data (by src.dtype); sidx / didx None = contiguous from 0."""
    n = len(sidx) if sidx is not None else len(didx)
    si = None if sidx is None else _i32(sidx); di = None if didx is None else _i32(didx)
    if np.asarray(src).dtype == np.int32:
        s = _i32(src); d = _i32(dst).copy()
        lib().ora_sf_scatter_i32(C.c_int64(n), int(bs), SF_OPS[op], None if si is None else _p(si), None if di is None else _p(di), _p(s), _p(d))
    else:
        s = _f64(src); d = _f64(dst).copy()
        lib().ora_sf_scatter_f64(C.c_int64(n), int(bs), SF_OPS[op], None if si is None else _p(si), None if di is None else _p(di), _p(s), _p(d))
    return d


def set_num_threads(n):
    lib().ora_set_num_threads(int(n))


def max_threads():
    return lib().ora_max_threads()
