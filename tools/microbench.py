#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA-event timed on the library's stream). Development tool; bench.py is the contract.

usage: python tools/microbench.py [--n 512] [--what spmv,blas1] [--out gpurun_out/micro.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petsc_b200 import _capi  # noqa: E402


def timeit(H, fn, reps=20, warm=3):
    t = _capi.Timer(H)
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        t.start(); fn(); t.stop()
        ms.append(t.ms())
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--what", default="spmv,blas1")
    ap.add_argument("--out", default="")
    ap.add_argument("--layouts", default="auto")
    a = ap.parse_args()
    L = _capi.lib()
    H = _capi.Handle(0)
    res = {"n": a.n}
    n = a.n
    N = n ** 3
    peak = 6570.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    if "spmv" in a.what:
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace7Nnz(n, n, n, C.c_int64(0), C.c_int64(N), C.byref(nnz)))
        nnz = nnz.value
        d_ai, d_aj, d_aa = H.empty(N + 1, np.int32), H.empty(nnz, np.int32), H.empty(nnz)
        _capi.check(L.b200GenLaplace7(H.h, n, n, n, C.c_int64(0), C.c_int64(N), d_ai.ptr, d_aj.ptr, d_aa.ptr))
        d_x, d_y, d_dinv = H.empty(N), H.empty(N), H.empty(N)
        _capi.check(L.b200VecSet(H.h, C.c_int64(N), C.c_double(1.0), d_x.ptr))
        _capi.check(L.b200VecSet(H.h, C.c_int64(N), C.c_double(1.0 / 6), d_dinv.ptr))
        plan = H.csr_plan(N, N, nnz, d_ai, d_aj)
        bytes_alg = nnz * 12 + N * 20
        flops = 2 * nnz - N
        res["spmv"] = []
        layouts = [(0, 0, 0, 0, 3)]
        if a.layouts == "sweep":
            layouts += [(1, r, s, c, hh) for hh in (2,) for (r, s, c) in ((256, 1, 6), (256, 1, 7), (256, 1, 8), (128, 1, 8), (512, 1, 4), (512, 1, 3), (128, 2, 8), (256, 2, 4), (64, 1, 8))]
        for (lanes, rows, stages, ctas, hh) in layouts:
            try:
                H.csr_plan_set_layout(plan, lanes, rows, stages, ctas)
                H.csr_plan_set_hints(plan, hh)
                lay = H.csr_plan_layout(plan)
                if lay["smem"] > 226 * 1024 // max(1, 1):
                    continue
                med, best = timeit(H, lambda: H.spmv(plan, d_aa, d_x, d_y))
            except _capi.B200Error as e:
                print("layout", lanes, rows, stages, ctas, hh, "failed:", e)
                continue
            r = dict(layout=lay, req=[lanes, rows, stages, ctas, hh], ms=med, ms_best=best, gbs=bytes_alg / med / 1e6, gflops=flops / med / 1e6,
                     frac_measured=bytes_alg / med / 1e6 / peak)
            res["spmv"].append(r)
            print("spmv", r)
        H.csr_plan_set_layout(plan, 0, 0, 0, 0)
        H.csr_plan_set_hints(plan, 3)
        med, best = timeit(H, lambda: H.spmv_jacobi(plan, d_aa, d_x, d_dinv, d_y))
        res["spmv_jacobi"] = dict(ms=med, gbs=(bytes_alg + 8 * N) / med / 1e6)
        print("spmv_jacobi", res["spmv_jacobi"])
        # sanity: A*1 has zero interior rows
        y = d_y.download()
        res["spmv_check_rowsum_interior"] = float(y[(n * n + n + 1)] * 6)
        del d_ai, d_aj, d_aa
    if "blas1" in a.what:
        nvmax = 30
        vecs = [H.empty(N) for _ in range(nvmax + 2)]
        for i, v in enumerate(vecs):
            _capi.check(L.b200VecSet(H.h, C.c_int64(N), C.c_double(1.0 / (i + 1)), v.ptr))
        x, y = vecs[nvmax], vecs[nvmax + 1]
        res["blas1"] = {}

        def rec(name, fn, nbytes):
            med, best = timeit(H, fn)
            res["blas1"][name] = dict(ms=med, gbs=nbytes / med / 1e6, frac_measured=nbytes / med / 1e6 / peak)
            print(name, res["blas1"][name])

        rec("copy", lambda: _capi.check(L.b200VecCopy(H.h, C.c_int64(N), x.ptr, y.ptr)), 16 * N)
        rec("axpy", lambda: _capi.check(L.b200VecAXPY(H.h, C.c_int64(N), C.c_double(1e-9), x.ptr, y.ptr)), 24 * N)
        rec("scale", lambda: _capi.check(L.b200VecScale(H.h, C.c_int64(N), C.c_double(1.0), y.ptr)), 16 * N)
        rec("pmult", lambda: _capi.check(L.b200VecPointwiseMult(H.h, C.c_int64(N), x.ptr, y.ptr, vecs[0].ptr)), 24 * N)
        rec("norm2", lambda: H.norm2(N, x), 8 * N)
        rec("dot", lambda: H.dot(N, x, y), 16 * N)
        for nv in (range(1, 31) if a.layouts == 'sweep' else (1, 2, 4, 8, 15, 16, 17, 24, 30)):
            rec("mdot%d" % nv, lambda nv=nv: H.mdot(N, x, vecs[:nv]), 8 * N * (nv + 1))
            al = [1e-12] * nv
            rec("maxpy%d" % nv, lambda nv=nv, al=al: H.maxpy(N, al, vecs[:nv], y), 8 * N * (nv + 2))
            rec("maxpy_norm%d" % nv, lambda nv=nv, al=al: H.maxpy(N, al, vecs[:nv], y, want_norm=True), 8 * N * (nv + 2))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
    H.close()


if __name__ == "__main__":
    main()
