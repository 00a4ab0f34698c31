#!/usr/bin/env python
"""The kernels to beat (SURVEY 8d last row): what the REFERENCE's own GPU back end (aijcusparse / veccuda) would launch for the
same operations, on the same B200, on the same device-resident inputs, next to this repository's kernels.

  MatMult_SeqAIJCUSPARSE  ->  cusparseSpMV, CUSPARSE_SPMV_CSR_ALG1 after cusparseSpMV_preprocess (aijcusparse.cu:2517-2529)
  VecMDot_SeqCUDA         ->  cublasDgemv('T') over the slab of basis vectors (vecseqcupm_impl.hpp gemv path)  [stand-in for the
                              reference's hand-written MDot_kernel, which needs a CUDA build of PETSc that does not exist here]
  VecMAXPY_SeqCUDA        ->  cublasDgemv('N')
  VecDot / VecNorm / AXPY ->  cublasDdot / cublasDnrm2 / cublasDaxpy

The library calls are made through ctypes on the CUDA toolkit's libcusparse / libcublas; both sides are timed with CUDA events
on the same stream after warm-up.  Writes one JSON document (and a markdown table with --md).

    python tools/kernel_to_beat.py [--out profiles/round2_kernel_to_beat.json] [--md profiles/round2_kernel_to_beat.md] [--small]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from petsc_b200 import _capi  # noqa: E402

vp = C.c_void_p


def load(name):
    for cand in (name, os.path.join("/usr/local/cuda/lib64", name)):
        try:
            return C.CDLL(cand, mode=C.RTLD_GLOBAL)
        except OSError:
            pass
    raise OSError("cannot load " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--md", default="")
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    L = _capi.lib()
    H = _capi.Handle()
    stream = vp()
    _capi.check(L.b200GetStream(H.h, C.byref(stream)))
    cs, cb = load("libcusparse.so.12"), load("libcublas.so.12")
    sp = vp(); bl = vp()
    assert cs.cusparseCreate(C.byref(sp)) == 0 and cs.cusparseSetStream(sp, stream) == 0
    assert cb.cublasCreate_v2(C.byref(bl)) == 0 and cb.cublasSetStream_v2(bl, stream) == 0
    peak = 6570.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def timed(fn, reps=20, warm=3):
        t = _capi.Timer(H)
        for _ in range(warm):
            fn()
        t.start()
        for _ in range(reps):
            fn()
        t.stop()
        return t.ms() / reps

    one, zero = C.c_double(1.0), C.c_double(0.0)
    rows = []

    def spmv_case(name, n, nnz, gen):
        d_i, d_j, d_a = _capi.DeviceArray(H, n + 1, np.int32), _capi.DeviceArray(H, nnz, np.int32), _capi.DeviceArray(H, nnz, np.float64)
        gen(d_i, d_j, d_a)
        x, y, y2 = _capi.DeviceArray(H, n, np.float64), _capi.DeviceArray(H, n, np.float64), _capi.DeviceArray(H, n, np.float64)
        _capi.check(L.b200VecSet(H.h, C.c_int64(n), C.c_double(1.0), x.ptr))
        # ours: the plan the PETSc plugin creates (automatic layout, automatic column blocks)
        plan = vp()
        _capi.check(L.b200CsrPlanCreate(H.h, n, n, C.c_int64(nnz), d_i.ptr, d_j.ptr, C.byref(plan)))
        nb = C.c_int(0)
        _capi.check(L.b200CsrPlanAutoColumnBlocks(H.h, plan, C.byref(nb)))
        _capi.check(L.b200CsrPlanPackValues(H.h, plan, d_a.ptr))
        ours = timed(lambda: _capi.check(L.b200CsrSpMV(H.h, plan, d_a.ptr, x.ptr, y.ptr)))
        # cuSPARSE generic API, as MatMult_SeqAIJCUSPARSE drives it
        mat, vx, vy = vp(), vp(), vp()
        assert cs.cusparseCreateCsr(C.byref(mat), C.c_int64(n), C.c_int64(n), C.c_int64(nnz), d_i.ptr, d_j.ptr, d_a.ptr, 2, 2, 0, 1) == 0   # INDEX_32I, INDEX_32I, BASE_ZERO, CUDA_R_64F
        assert cs.cusparseCreateDnVec(C.byref(vx), C.c_int64(n), x.ptr, 1) == 0 and cs.cusparseCreateDnVec(C.byref(vy), C.c_int64(n), y2.ptr, 1) == 0
        res = {}
        for alg_name, alg in (("CSR_ALG1", 2), ("CSR_ALG2", 3)):
            bs = C.c_size_t(0)
            assert cs.cusparseSpMV_bufferSize(sp, 0, C.byref(one), mat, vx, C.byref(zero), vy, 1, alg, C.byref(bs)) == 0
            buf = _capi.DeviceArray(H, max(bs.value, 8), np.uint8)
            if hasattr(cs, "cusparseSpMV_preprocess"):
                cs.cusparseSpMV_preprocess(sp, 0, C.byref(one), mat, vx, C.byref(zero), vy, 1, alg, buf.ptr)
            res[alg_name] = timed(lambda: cs.cusparseSpMV(sp, 0, C.byref(one), mat, vx, C.byref(zero), vy, 1, alg, buf.ptr))
            buf.free()
        ya, yb = y.download(), y2.download()
        err = float(np.abs(ya - yb).max() / max(np.abs(yb).max(), 1e-300))
        alg_bytes = nnz * 12 + n * 20
        best = min(res.values())
        rows.append(dict(op="MatMult " + name, n=n, nnz=nnz, ours_ms=round(ours, 4), cusparse_alg1_ms=round(res["CSR_ALG1"], 4), cusparse_alg2_ms=round(res["CSR_ALG2"], 4),
                         speedup_vs_best_cusparse=round(best / ours, 3), ours_gbs=round(alg_bytes / ours / 1e6, 1), ours_frac_of_measured_peak=round(alg_bytes / ours / 1e6 / peak, 3),
                         cusparse_best_gbs=round(alg_bytes / best / 1e6, 1), column_blocks=nb.value, rel_diff=err))
        print(rows[-1], flush=True)
        cs.cusparseDestroySpMat(mat); cs.cusparseDestroyDnVec(vx); cs.cusparseDestroyDnVec(vy)
        L.b200CsrPlanDestroy(plan)
        for o in (d_i, d_j, d_a, x, y, y2):
            o.free()

    def nnz7(n):
        v = C.c_int64()
        _capi.check(L.b200GenLaplace7Nnz(n, n, n, C.c_int64(0), C.c_int64(n ** 3), C.byref(v)))
        return v.value

    def nnz27(n):
        v = C.c_int64()
        _capi.check(L.b200GenLaplace27Nnz(n, C.byref(v)))
        return v.value
    n7, n27, nr = (128, 64, 1_000_000) if a.small else (512, 256, 10_000_000)
    spmv_case("7-pt %d^3" % n7, n7 ** 3, nnz7(n7), lambda i, j, v: _capi.check(L.b200GenLaplace7(H.h, n7, n7, n7, C.c_int64(0), C.c_int64(n7 ** 3), i.ptr, j.ptr, v.ptr)))
    spmv_case("27-pt %d^3" % n27, n27 ** 3, nnz27(n27), lambda i, j, v: _capi.check(L.b200GenLaplace27(H.h, n27, i.ptr, j.ptr, v.ptr)))
    for d in (5, 32, 128, 512):
        n = nr if d < 512 else nr // 4
        spmv_case("random d=%d" % d, n, n * d, lambda i, j, v, n=n, d=d: _capi.check(L.b200GenRandomCsr(H.h, n, n, d, C.c_uint64(20260923 + d), i.ptr, j.ptr, v.ptr)))

    # ---- a power-law matrix: 2.5 M scattered rows of 4 entries + a tail of very long rows (row bins vs one library call)
    if not a.small:
        from oracle import oracle_py as O   # generator of the test inputs only
        npl = 2_500_000
        rng = np.random.default_rng(33)
        ai0, aj0, aa0 = O.random_csr(npl, 4, 5)
        lens = np.full(npl, 4, np.int64)
        longrows = {7: 600_000, 1000: 120_000, 123_456: 40_000, 2_000_001: 9_000, npl - 1: 20_000}
        for r, l in longrows.items():
            lens[r] = l
        ai = np.zeros(npl + 1, np.int64); ai[1:] = np.cumsum(lens)
        aj = np.empty(ai[-1], np.int32); av = np.empty(ai[-1])
        short = np.ones(npl, bool); short[list(longrows)] = False
        inew = (ai[:-1][short][:, None] + np.arange(4)[None, :]).ravel(); iold = (ai0[:-1].astype(np.int64)[short][:, None] + np.arange(4)[None, :]).ravel()
        aj[inew] = aj0[iold]; av[inew] = aa0[iold]
        for r, l in longrows.items():
            aj[ai[r]:ai[r + 1]] = np.sort(rng.choice(npl, l, replace=False)).astype(np.int32); av[ai[r]:ai[r + 1]] = rng.uniform(-1, 1, l)
        ai32 = ai.astype(np.int32)
        spmv_case("power-law 2.5M rows (4/row + tail to 600k)", npl, int(ai[-1]), lambda i, j, v: (i.upload(ai32), j.upload(aj), v.upload(av)))

    # ---- BLAS-1 / orthogonalisation at the size of the headline workload
    n = (128 if a.small else 512) ** 3
    nv = 30
    lda = (n + 31) & ~31
    slab = _capi.DeviceArray(H, lda * nv, np.float64)
    x, res_d = _capi.DeviceArray(H, n, np.float64), _capi.DeviceArray(H, 64, np.float64)
    _capi.check(L.b200VecSet(H.h, C.c_int64(lda * nv), C.c_double(0.5), slab.ptr))
    _capi.check(L.b200VecSet(H.h, C.c_int64(n), C.c_double(1.0), x.ptr))
    ptrs = (vp * nv)(*[vp(slab.ptr.value + 8 * lda * j) for j in range(nv)])
    alpha = (C.c_double * nv)(*[1e-3 * (j + 1) for j in range(nv)])
    hres = (C.c_double * 64)()
    d_alpha = _capi.DeviceArray(H, nv, np.float64).upload(np.array(list(alpha)))
    ours_mdot = timed(lambda: _capi.check(L.b200VecMDot(H.h, C.c_int64(n), nv, x.ptr, ptrs, hres)), reps=10)
    cb.cublasSetPointerMode_v2(bl, 1)  # device pointers for scalars/results: no host sync inside the library call
    d_one = _capi.DeviceArray(H, 1, np.float64).upload(np.array([1.0])); d_zero = _capi.DeviceArray(H, 1, np.float64).upload(np.array([0.0]))
    lib_mdot = timed(lambda: cb.cublasDgemv_v2(bl, 1, n, nv, d_one.ptr, slab.ptr, lda, x.ptr, 1, d_zero.ptr, res_d.ptr, 1), reps=10)   # op T
    rows.append(dict(op="VecMDot nv=30 (n=%d)" % n, ours_ms=round(ours_mdot, 4), cublas_dgemv_T_ms=round(lib_mdot, 4), speedup=round(lib_mdot / ours_mdot, 3),
                     ours_gbs=round(8 * n * (nv + 1) / ours_mdot / 1e6, 1), note="ours returns the results to the host (synchronises); the library call leaves them on the device"))
    print(rows[-1], flush=True)
    ours_maxpy = timed(lambda: _capi.check(L.b200VecMAXPY(H.h, C.c_int64(n), nv, alpha, ptrs, x.ptr, None)), reps=10)
    lib_maxpy = timed(lambda: cb.cublasDgemv_v2(bl, 0, n, nv, d_one.ptr, slab.ptr, lda, d_alpha.ptr, 1, d_one.ptr, x.ptr, 1), reps=10)                # op N: x += V alpha
    rows.append(dict(op="VecMAXPY nv=30 (n=%d)" % n, ours_ms=round(ours_maxpy, 4), cublas_dgemv_N_ms=round(lib_maxpy, 4), speedup=round(lib_maxpy / ours_maxpy, 3),
                     ours_gbs=round(8 * n * (nv + 2) / ours_maxpy / 1e6, 1)))
    print(rows[-1], flush=True)
    y = _capi.DeviceArray(H, n, np.float64)
    _capi.check(L.b200VecSet(H.h, C.c_int64(n), C.c_double(2.0), y.ptr))
    hv = C.c_double()
    for nm, ours_fn, lib_fn, byt in (
            ("VecDot", lambda: _capi.check(L.b200VecDot(H.h, C.c_int64(n), x.ptr, y.ptr, C.byref(hv))), lambda: cb.cublasDdot_v2(bl, n, x.ptr, 1, y.ptr, 1, res_d.ptr), 16 * n),
            ("VecNorm2", lambda: _capi.check(L.b200VecNorm2(H.h, C.c_int64(n), x.ptr, C.byref(hv))), lambda: cb.cublasDnrm2_v2(bl, n, x.ptr, 1, res_d.ptr), 8 * n),
            ("VecAXPY", lambda: _capi.check(L.b200VecAXPY(H.h, C.c_int64(n), C.c_double(1e-3), x.ptr, y.ptr)), lambda: cb.cublasDaxpy_v2(bl, n, d_alpha.ptr, x.ptr, 1, y.ptr, 1), 24 * n)):
        o, l = timed(ours_fn, reps=10), timed(lib_fn, reps=10)
        rows.append(dict(op="%s (n=%d)" % (nm, n), ours_ms=round(o, 4), cublas_ms=round(l, 4), speedup=round(l / o, 3), ours_gbs=round(byt / o / 1e6, 1)))
        print(rows[-1], flush=True)
    doc = dict(peak_gbs=peak, rows=rows, note="both sides timed with CUDA events on one stream after warm-up, device-resident inputs; cuSPARSE generic SpMV with preprocess as aijcusparse.cu:2517-2529 drives it")
    if a.out:
        json.dump(doc, open(a.out, "w"), indent=1)
    if a.md:
        with open(a.md, "w") as f:
            f.write("| operation | ours (ms) | library (ms) | speed-up | ours GB/s (algorithmic) |\n|---|---|---|---|---|\n")
            for r in rows:
                lib = r.get("cusparse_alg1_ms", r.get("cublas_ms", r.get("cublas_dgemv_T_ms", r.get("cublas_dgemv_N_ms"))))
                if "cusparse_alg2_ms" in r:
                    lib = "ALG1 %.4g / ALG2 %.4g" % (r["cusparse_alg1_ms"], r["cusparse_alg2_ms"])
                f.write("| %s | %.4g | %s | %.3g | %.0f |\n" % (r["op"], r["ours_ms"], lib, r.get("speedup_vs_best_cusparse", r.get("speedup")), r["ours_gbs"]))


if __name__ == "__main__":
    main()
