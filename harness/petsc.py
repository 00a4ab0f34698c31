"""TEST HARNESS: ctypes door to harness/libb200harness.so, the C mini-PETSc of harness/host (harness/petscb200_host.h).

Thin by design: every method is one call into the C library (which calls the sm_100a kernels).  Names follow petsc4py
loosely so tests read like the reference's examples.  No CPU fallback: without the CUDA libraries this module raises.
"""
import ctypes as C
import os

import numpy as np

from petsc_b200 import _capi

_HARNESS_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200harness.so")

_host = None
vp, dbl, i32 = C.c_void_p, C.c_double, C.c_int

COMM_WORLD, COMM_SELF = 1, 2
DECIDE = -1


class PetscError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("PETSc(b200) error %d\n%s" % (code, msg))
        self.code = code


def lib():
    global _host
    if _host is None:
        _capi.lib()  # kernel library first (RTLD_GLOBAL)
        if not os.path.exists(_HARNESS_LIB):
            raise ImportError("harness/libb200harness.so is not built; run __graft_entry__.build()")
        _host = C.CDLL(_HARNESS_LIB, mode=C.RTLD_GLOBAL)
        _host.PetscB200GetLastErrorMessage.restype = C.c_char_p
    return _host


def chk(rc):
    if rc:
        raise PetscError(rc, lib().PetscB200GetLastErrorMessage().decode(errors="replace"))


def host_symbols():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "petscb200_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(?:PetscErrorCode|const char\s*\*)\s*(\w+)\s*\(", hdr)))


def initialize(options="", device=None):
    L = lib()
    if device is not None:
        chk(L.PetscB200SetDevice(int(device)))
    chk(L.PetscInitializeNoArguments())
    if options:
        chk(L.PetscOptionsInsertString(None, options.encode()))


def finalize():
    chk(lib().PetscFinalize())


def options_set(name, value=None):
    chk(lib().PetscOptionsSetValue(None, name.encode(), None if value is None else str(value).encode()))


def options_insert(s):
    chk(lib().PetscOptionsInsertString(None, s.encode()))


def options_clear():
    chk(lib().PetscOptionsClear(None))


def comm_unique_id():
    buf = (C.c_char * 128)()
    chk(lib().PetscB200CommGetUniqueId(buf))
    return bytes(buf)


def comm_init(rank, size, uid):
    chk(lib().PetscB200CommInit(int(rank), int(size), C.c_char_p(uid)))


def handle():
    h = vp()
    chk(lib().PetscB200GetHandle(C.byref(h)))
    return h


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Vec:
    def __init__(self, ptr=None, own=True):
        self.p = ptr if ptr is not None else vp()
        self.own = own

    @classmethod
    def create(cls, n=DECIDE, N=DECIDE, comm=COMM_WORLD, vtype=None):
        v = cls()
        chk(lib().VecCreate(comm, C.byref(v.p)))
        chk(lib().VecSetSizes(v.p, int(n), int(N)))
        if vtype:
            chk(lib().VecSetType(v.p, vtype.encode()))
        else:
            chk(lib().VecSetFromOptions(v.p))
        return v

    @classmethod
    def from_array(cls, a, comm=COMM_SELF):
        a = _f64(a)
        v = cls.create(n=len(a), N=len(a) if comm == COMM_SELF else DECIDE, comm=comm)
        v.set_array(a)
        return v

    @classmethod
    def with_array(cls, host_ptr, n, N=None, comm=COMM_SELF):
        """VecCreateSeqWithArray / VecCreateMPIWithArray: the caller's HOST buffer (a raw pointer, e.g. pinned memory from
        b200MallocHost, or a numpy array that outlives the Vec) is the vector's host storage; nothing is copied here."""
        v = cls()
        if isinstance(host_ptr, np.ndarray):
            v._keep = host_ptr
            host_ptr = host_ptr.ctypes.data_as(vp)
        if N is None:
            chk(lib().VecCreateSeqWithArray(comm, 1, int(n), host_ptr, C.byref(v.p)))
        else:
            chk(lib().VecCreateMPIWithArray(comm, 1, int(n), int(N), host_ptr, C.byref(v.p)))
        return v

    def place_array(self, host_ptr):
        if isinstance(host_ptr, np.ndarray):
            self._keep = host_ptr
            host_ptr = host_ptr.ctypes.data_as(vp)
        chk(lib().VecPlaceArray(self.p, host_ptr))

    def reset_array(self):
        chk(lib().VecResetArray(self.p))

    def duplicate(self):
        w = Vec()
        chk(lib().VecDuplicate(self.p, C.byref(w.p)))
        return w

    def destroy(self):
        if self.own and self.p:
            chk(lib().VecDestroy(C.byref(self.p)))
        self.p = vp()

    def get_type(self):
        t = C.c_char_p()
        chk(lib().VecGetType(self.p, C.byref(t)))
        return t.value.decode()

    def local_size(self):
        n = i32()
        chk(lib().VecGetLocalSize(self.p, C.byref(n)))
        return n.value

    def size(self):
        n = i32()
        chk(lib().VecGetSize(self.p, C.byref(n)))
        return n.value

    def ownership_range(self):
        a, b = i32(), i32()
        chk(lib().VecGetOwnershipRange(self.p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_array(self, a):
        a = _f64(a)
        p = C.POINTER(dbl)()
        chk(lib().VecGetArrayWrite(self.p, C.byref(p)))
        n = self.local_size()
        assert len(a) == n
        C.memmove(p, a.ctypes.data, 8 * n)
        chk(lib().VecRestoreArrayWrite(self.p, C.byref(p)))

    def array(self):
        p = C.POINTER(dbl)()
        n = self.local_size()
        chk(lib().VecGetArrayRead(self.p, C.byref(p)))
        out = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0)
        chk(lib().VecRestoreArrayRead(self.p, C.byref(p)))
        return out

    def touch_host(self):
        """Declare the pinned host mirror the valid copy (VecGetArrayWrite/Restore): the next device op uploads it."""
        p = C.POINTER(dbl)()
        chk(lib().VecGetArrayWrite(self.p, C.byref(p)))
        chk(lib().VecRestoreArrayWrite(self.p, C.byref(p)))

    def host_read(self, k=4):
        """Bring the vector to the pinned host mirror (VecGetArrayRead: device->host copy) and return its first k entries."""
        p = C.POINTER(dbl)()
        chk(lib().VecGetArrayRead(self.p, C.byref(p)))
        out = [p[i] for i in range(min(k, self.local_size()))]
        chk(lib().VecRestoreArrayRead(self.p, C.byref(p)))
        return out

    def device_ptr(self):
        p = vp()
        mt = i32()
        chk(lib().VecGetArrayReadAndMemType(self.p, C.byref(p), C.byref(mt)))
        chk(lib().VecRestoreArrayReadAndMemType(self.p, C.byref(p)))
        return p.value

    def set(self, alpha):
        chk(lib().VecSet(self.p, dbl(alpha)))

    def copy_to(self, y):
        chk(lib().VecCopy(self.p, y.p))

    def scale(self, a):
        chk(lib().VecScale(self.p, dbl(a)))

    def axpy(self, a, x):
        chk(lib().VecAXPY(self.p, dbl(a), x.p))

    def aypx(self, b, x):
        chk(lib().VecAYPX(self.p, dbl(b), x.p))

    def axpby(self, a, b, x):
        chk(lib().VecAXPBY(self.p, dbl(a), dbl(b), x.p))

    def waxpy(self, a, x, y):
        chk(lib().VecWAXPY(self.p, dbl(a), x.p, y.p))

    def maxpy(self, alpha, xs):
        nv = len(xs)
        arr = (vp * nv)(*[x.p for x in xs])
        al = (dbl * nv)(*[float(a) for a in alpha])
        chk(lib().VecMAXPY(self.p, nv, al, arr))

    def dot(self, y):
        r = dbl()
        chk(lib().VecDot(self.p, y.p, C.byref(r)))
        return r.value

    def mdot(self, ys):
        nv = len(ys)
        arr = (vp * nv)(*[y.p for y in ys])
        out = (dbl * nv)()
        chk(lib().VecMDot(self.p, nv, arr, out))
        return np.array(out[:])

    def norm(self, ntype=1):
        r = dbl()
        chk(lib().VecNorm(self.p, int(ntype), C.byref(r)))
        return r.value

    def normalize(self):
        r = dbl()
        chk(lib().VecNormalize(self.p, C.byref(r)))
        return r.value

    def pointwise_mult(self, x, y):
        chk(lib().VecPointwiseMult(self.p, x.p, y.p))

    def reciprocal(self):
        chk(lib().VecReciprocal(self.p))

    def sum(self):
        r = dbl()
        chk(lib().VecSum(self.p, C.byref(r)))
        return r.value

    def max(self):
        r, i = dbl(), i32()
        chk(lib().VecMax(self.p, C.byref(i), C.byref(r)))
        return i.value, r.value

    def min(self):
        r, i = dbl(), i32()
        chk(lib().VecMin(self.p, C.byref(i), C.byref(r)))
        return i.value, r.value


def duplicate_vecs(v, m):
    arr = C.POINTER(vp)()
    chk(lib().VecDuplicateVecs(v.p, int(m), C.byref(arr)))
    vs = [Vec(vp(arr[i]), own=False) for i in range(m)]
    return vs, arr


def destroy_vecs(m, arr):
    chk(lib().VecDestroyVecs(int(m), C.byref(arr)))


class Mat:
    def __init__(self, ptr=None, own=True):
        self.p = ptr if ptr is not None else vp()
        self.own = own

    @classmethod
    def create(cls, m=DECIDE, n=DECIDE, M=DECIDE, N=DECIDE, comm=COMM_WORLD, mtype=None):
        A = cls()
        chk(lib().MatCreate(comm, C.byref(A.p)))
        chk(lib().MatSetSizes(A.p, int(m), int(n), int(M), int(N)))
        if mtype:
            chk(lib().MatSetType(A.p, mtype.encode()))
        else:
            chk(lib().MatSetFromOptions(A.p))
        return A

    @classmethod
    def from_csr(cls, ai, aj, aa, ncols=None, comm=COMM_SELF):
        """Sequential matrix from host CSR (MatCreateSeqAIJWithArrays)."""
        ai, aj, aa = _i32(ai), _i32(aj), _f64(aa)
        m = len(ai) - 1
        A = cls()
        chk(lib().MatCreateSeqAIJWithArrays(comm, m, m if ncols is None else ncols, ai.ctypes.data_as(vp), aj.ctypes.data_as(vp),
                                            aa.ctypes.data_as(vp), C.byref(A.p)))
        return A

    @classmethod
    def from_local_csr(cls, ai, aj_global, aa, N, comm=COMM_WORLD):
        """Row-partitioned matrix from this rank's rows with global column indices (MatMPIAIJSetPreallocationCSR)."""
        ai, aj, aa = _i32(ai), _i32(aj_global), _f64(aa)
        m = len(ai) - 1
        A = cls.create(m=m, n=m, M=DECIDE, N=DECIDE, comm=comm)
        chk(lib().MatMPIAIJSetPreallocationCSR(A.p, ai.ctypes.data_as(vp), aj.ctypes.data_as(vp), aa.ctypes.data_as(vp)))
        return A

    def set_csr_device(self, d_i, d_j, d_a):
        chk(lib().MatB200SetCSRDevice(self.p, d_i, d_j, d_a))

    def set_values(self, rows, cols, vals, add=True):
        rows, cols, vals = _i32(rows), _i32(cols), _f64(vals)
        chk(lib().MatSetValues(self.p, len(rows), rows.ctypes.data_as(vp), len(cols), cols.ctypes.data_as(vp), vals.ctypes.data_as(vp), 2 if add else 1))

    def assemble(self):
        chk(lib().MatAssemblyBegin(self.p, 0))
        chk(lib().MatAssemblyEnd(self.p, 0))

    def get_type(self):
        t = C.c_char_p()
        chk(lib().MatGetType(self.p, C.byref(t)))
        return t.value.decode()

    def sizes(self):
        m, n, M, N = i32(), i32(), i32(), i32()
        chk(lib().MatGetLocalSize(self.p, C.byref(m), C.byref(n)))
        chk(lib().MatGetSize(self.p, C.byref(M), C.byref(N)))
        return (m.value, n.value), (M.value, N.value)

    def ownership_range(self):
        a, b = i32(), i32()
        chk(lib().MatGetOwnershipRange(self.p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def create_vecs(self):
        r, l = Vec(), Vec()
        chk(lib().MatCreateVecs(self.p, C.byref(r.p), C.byref(l.p)))
        return r, l

    def mult(self, x, y):
        chk(lib().MatMult(self.p, x.p, y.p))

    def mult_add(self, x, y, z):
        chk(lib().MatMultAdd(self.p, x.p, y.p, z.p))

    def mult_transpose(self, x, y):
        chk(lib().MatMultTranspose(self.p, x.p, y.p))

    def mult_transpose_add(self, x, y, z):
        chk(lib().MatMultTransposeAdd(self.p, x.p, y.p, z.p))

    def set_preallocation_coo(self, coo_i, coo_j, n=None):
        """MatSetPreallocationCOO: numpy int arrays (host) or raw device pointers (then pass n)."""
        if n is None:
            ci, cj = _i32(coo_i), _i32(coo_j)
            self._coo_keep = (ci, cj)
            chk(lib().MatSetPreallocationCOO(self.p, C.c_int64(len(ci)), ci.ctypes.data_as(vp), cj.ctypes.data_as(vp)))
        else:
            chk(lib().MatSetPreallocationCOO(self.p, C.c_int64(n), coo_i, coo_j))

    def set_values_coo(self, v, add=False):
        """MatSetValuesCOO(INSERT_VALUES | ADD_VALUES): numpy array (host) or a raw device pointer."""
        if isinstance(v, np.ndarray) or isinstance(v, (list, tuple)):
            v = _f64(v)
            chk(lib().MatSetValuesCOO(self.p, v.ctypes.data_as(vp), 2 if add else 1))
        else:
            chk(lib().MatSetValuesCOO(self.p, v, 2 if add else 1))

    def get_diagonal(self, v):
        chk(lib().MatGetDiagonal(self.p, v.p))

    def csr_nnz(self):
        info = (dbl * 10)()
        chk(lib().MatGetInfo(self.p, 1, info))
        return int(info[2])

    def set_spmv_layout(self, lanes=0, rows=0, stages=0, ctas=0):
        chk(lib().MatB200SetSpMVLayout(self.p, lanes, rows, stages, ctas))

    def set_spmv_ordered(self, ordered=True):
        chk(lib().MatB200SetSpMVOrdered(self.p, 1 if ordered else 0))

    def mpiaij_blocks(self):
        Ad, Ao, cm = vp(), vp(), C.POINTER(i32)()
        chk(lib().MatMPIAIJGetSeqAIJ(self.p, C.byref(Ad), C.byref(Ao), C.byref(cm)))
        A, B = Mat(Ad, own=False), Mat(Ao, own=False)
        (_, ec), _ = B.sizes()
        garray = np.ctypeslib.as_array(cm, shape=(ec,)).copy() if ec else np.zeros(0, np.int32)
        return A, B, garray

    def csr_host(self):
        m, pi, pj, pa = i32(), C.POINTER(i32)(), C.POINTER(i32)(), C.POINTER(dbl)()
        chk(lib().MatSeqAIJGetCSRHost(self.p, C.byref(m), C.byref(pi), C.byref(pj), C.byref(pa)))
        ai = np.ctypeslib.as_array(pi, shape=(m.value + 1,)).copy()
        nz = int(ai[-1])
        aj = np.ctypeslib.as_array(pj, shape=(nz,)).copy() if nz else np.zeros(0, np.int32)
        aa = np.ctypeslib.as_array(pa, shape=(nz,)).copy() if nz else np.zeros(0)
        return ai, aj, aa

    def destroy(self):
        if self.own and self.p:
            chk(lib().MatDestroy(C.byref(self.p)))
        self.p = vp()


class PC:
    def __init__(self, ptr=None, own=True):
        self.p = ptr if ptr is not None else vp()
        self.own = own

    @classmethod
    def create(cls, comm=COMM_WORLD, pctype=None):
        pc = cls()
        chk(lib().PCCreate(comm, C.byref(pc.p)))
        if pctype:
            chk(lib().PCSetType(pc.p, pctype.encode()))
        return pc

    def set_type(self, t):
        chk(lib().PCSetType(self.p, t.encode()))

    def get_type(self):
        t = C.c_char_p()
        chk(lib().PCGetType(self.p, C.byref(t)))
        return t.value.decode()

    def set_from_options(self):
        chk(lib().PCSetFromOptions(self.p))

    def set_operators(self, A, P=None):
        chk(lib().PCSetOperators(self.p, A.p, (P or A).p))

    def setup(self):
        chk(lib().PCSetUp(self.p))

    def apply(self, x, y):
        chk(lib().PCApply(self.p, x.p, y.p))

    def destroy(self):
        if self.own and self.p:
            chk(lib().PCDestroy(C.byref(self.p)))
        self.p = vp()


class KSP:
    def __init__(self):
        self.p = vp()
        self._hist = None

    @classmethod
    def create(cls, comm=COMM_WORLD):
        k = cls()
        chk(lib().KSPCreate(comm, C.byref(k.p)))
        return k

    def set_operators(self, A, P=None):
        chk(lib().KSPSetOperators(self.p, A.p, (P or A).p))

    def set_type(self, t):
        chk(lib().KSPSetType(self.p, t.encode()))

    def get_pc(self):
        p = vp()
        chk(lib().KSPGetPC(self.p, C.byref(p)))
        return PC(p, own=False)

    def set_tolerances(self, rtol=-3.0, atol=-3.0, divtol=-3.0, max_it=-3):
        chk(lib().KSPSetTolerances(self.p, dbl(rtol), dbl(atol), dbl(divtol), int(max_it)))

    def set_from_options(self):
        chk(lib().KSPSetFromOptions(self.p))

    def set_residual_history(self, n=100000):
        self._hist = np.zeros(n, np.float64)
        chk(lib().KSPSetResidualHistory(self.p, self._hist.ctypes.data_as(vp), int(n), 1))

    def solve(self, b, x):
        chk(lib().KSPSolve(self.p, b.p, x.p))

    def its(self):
        n = i32()
        chk(lib().KSPGetIterationNumber(self.p, C.byref(n)))
        return n.value

    def reason(self):
        n = i32()
        chk(lib().KSPGetConvergedReason(self.p, C.byref(n)))
        return n.value

    def rnorm(self):
        r = dbl()
        chk(lib().KSPGetResidualNorm(self.p, C.byref(r)))
        return r.value

    def history(self):
        n = i32()
        p = vp()
        chk(lib().KSPGetResidualHistory(self.p, C.byref(p), C.byref(n)))
        return self._hist[:n.value].copy()

    def destroy(self):
        if self.p:
            chk(lib().KSPDestroy(C.byref(self.p)))
        self.p = vp()
