"""ctypes binding of include/petscb200.h (libpetscb200.so).  Thin: every call goes straight through the C ABI.

The CUDA extension is the product: if the shared library is missing or a call fails this module raises -- there is no
CPU fallback anywhere in the package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpetscb200.so")

_lib = None


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("petscb200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpetscb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `python -m petsc_b200.build`. There is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.b200GetLastErrorString.restype = C.c_char_p
        _lib.b200Version.restype = C.c_char_p
        _lib.b200KernelLaunchCount.restype = C.c_longlong
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200GetLastErrorString().decode())


def exported_symbols():
    """Names declared in include/petscb200.h (parsed from the header)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "petscb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200[A-Z]\w*)\s*\(", hdr)))


i64 = C.c_int64
dbl = C.c_double
vp = C.c_void_p


class DeviceArray:
    """A device allocation made with b200Malloc, typed for convenience."""

    def __init__(self, handle, n, dtype):
        self.h = handle
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.ptr = vp()
        check(lib().b200Malloc(handle.h, C.byref(self.ptr), C.c_size_t(self.n * self.dtype.itemsize)))

    @property
    def nbytes(self):
        return self.n * self.dtype.itemsize

    def upload(self, arr):
        a = np.ascontiguousarray(arr, dtype=self.dtype)
        assert a.size == self.n, (a.size, self.n)
        check(lib().b200MemcpyHtoD(self.h.h, self.ptr, a.ctypes.data_as(vp), C.c_size_t(a.nbytes)))
        return self

    def download(self):
        out = np.empty(self.n, self.dtype)
        check(lib().b200MemcpyDtoH(self.h.h, out.ctypes.data_as(vp), self.ptr, C.c_size_t(out.nbytes)))
        return out

    def offset(self, elems):
        return vp(self.ptr.value + elems * self.dtype.itemsize)

    def free(self):
        if self.ptr and self.ptr.value:
            lib().b200Free(self.h.h, self.ptr)
            self.ptr = vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Handle:
    def __init__(self, device=-1):
        self.h = vp()
        check(lib().b200Create(C.byref(self.h), int(device)))

    def close(self):
        if self.h:
            lib().b200Destroy(self.h)
            self.h = vp()

    def sync(self):
        check(lib().b200Synchronize(self.h))

    def array(self, arr, dtype=None):
        a = np.ascontiguousarray(arr, dtype=dtype)
        return DeviceArray(self, a.size, a.dtype).upload(a)

    def empty(self, n, dtype=np.float64):
        return DeviceArray(self, n, dtype)

    def zeros(self, n, dtype=np.float64):
        d = DeviceArray(self, n, dtype)
        check(lib().b200Memset(self.h, d.ptr, 0, C.c_size_t(d.nbytes)))
        return d

    # ---- CSR
    def csr_plan(self, m, n, nnz, d_rowptr, d_colidx):
        p = vp()
        check(lib().b200CsrPlanCreate(self.h, int(m), int(n), i64(nnz), d_rowptr.ptr, d_colidx.ptr, C.byref(p)))
        return p

    def csr_plan_layout(self, plan):
        v = [C.c_int() for _ in range(6)]
        check(lib().b200CsrPlanGetLayout(plan, *[C.byref(x) for x in v]))
        return dict(zip(("lanes", "rows_per_tile", "stages", "grid", "smem", "max_row_nnz"), [x.value for x in v]))

    def csr_plan_set_layout(self, plan, lanes=0, rows=0, stages=0, ctas=0):
        check(lib().b200CsrPlanSetLayout(plan, lanes, rows, stages, ctas))

    def csr_plan_set_hints(self, plan, hints):
        check(lib().b200CsrPlanSetCacheHints(plan, int(hints)))

    def spmv(self, plan, d_val, d_x, d_y):
        check(lib().b200CsrSpMV(self.h, plan, d_val.ptr, d_x.ptr, d_y.ptr))

    def spmv_add(self, plan, d_val, d_x, d_y, d_z):
        check(lib().b200CsrSpMVAdd(self.h, plan, d_val.ptr, d_x.ptr, d_y.ptr, d_z.ptr))

    def spmv_jacobi(self, plan, d_val, d_x, d_dinv, d_w, d_y=None):
        check(lib().b200CsrSpMVJacobi(self.h, plan, d_val.ptr, d_x.ptr, d_dinv.ptr, d_w.ptr, d_y.ptr if d_y is not None else None))

    # ---- BLAS-1 helpers used by tests
    def mdot(self, n, d_x, d_ys):
        nv = len(d_ys)
        ptrs = (vp * nv)(*[(y.ptr if isinstance(y, DeviceArray) else y) for y in d_ys])
        out = (dbl * nv)()
        check(lib().b200VecMDot(self.h, i64(n), nv, d_x.ptr, ptrs, out))
        return np.array(out[:], dtype=np.float64)

    def maxpy(self, n, alpha, d_ys, d_x, want_norm=False):
        nv = len(d_ys)
        ptrs = (vp * nv)(*[(y.ptr if isinstance(y, DeviceArray) else y) for y in d_ys])
        al = (dbl * nv)(*[float(a) for a in alpha])
        nrm = dbl(0)
        check(lib().b200VecMAXPY(self.h, i64(n), nv, al, ptrs, d_x.ptr, C.byref(nrm) if want_norm else None))
        return nrm.value if want_norm else None

    def dot(self, n, d_x, d_y):
        r = dbl()
        check(lib().b200VecDot(self.h, i64(n), d_x.ptr, d_y.ptr, C.byref(r)))
        return r.value

    def norm2(self, n, d_x):
        r = dbl()
        check(lib().b200VecNorm2(self.h, i64(n), d_x.ptr, C.byref(r)))
        return r.value


class Timer:
    """CUDA-event stopwatch on the handle's stream."""

    def __init__(self, handle):
        self.h = handle
        self.a, self.b = vp(), vp()
        check(lib().b200EventCreate(C.byref(self.a)))
        check(lib().b200EventCreate(C.byref(self.b)))

    def start(self):
        check(lib().b200EventRecord(self.h.h, self.a))

    def stop(self):
        check(lib().b200EventRecord(self.h.h, self.b))

    def ms(self):
        r = dbl()
        check(lib().b200EventElapsedMs(self.a, self.b, C.byref(r)))
        return r.value


def launch_count():
    return int(lib().b200KernelLaunchCount())
