#!/usr/bin/env python
"""Measurements for the other BASELINE.json configs (bench.py is configs[1]); writes one JSON object.

  config 3: 27-point 256^3, KSPCG + PCILU(0)  (ILU(0) numeric on device, sync-free sweeps)
  config 5: random CSR sweep n = 10 M (2.5 M for d = 512: 32-bit PetscInt), d in {5, 32, 128, 512}, MatMult only
  config 1: ex2 100x100 GMRES(30)+Jacobi (latency-bound small case)
  4       : the per-rank block of config 4 (7-point ILU(0) PCApply);  nb (with 5): column-blocked passes;
  tr / coo: transposed product and COO assembly on the 27-point operator (--nasm);  p: KSPPIPECG vs KSPCG

usage: python tools/bench_configs.py [--what 3,5,1,4,nb,tr,coo,p] [--n27 256] [--out gpurun_out/configs.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petsc_b200 import _capi
from harness import petsc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="3,5,1")
    ap.add_argument("--n27", type=int, default=256)
    ap.add_argument("--nrand", type=int, default=10_000_000)
    ap.add_argument("--n7", type=int, default=384)
    ap.add_argument("--nasm", type=int, default=160)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    L = _capi.lib()
    petsc.initialize(device=0)
    H = petsc.handle()

    class Hh:
        h = H
    peak = 6570.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    res = {"peak_gbs": peak}
    what = a.what.split(",")

    def timed(fn, reps, warm=2):
        t = _capi.Timer(Hh)
        for _ in range(warm):
            fn()
        t.start()
        for _ in range(reps):
            fn()
        t.stop()
        return t.ms() / reps

    if "5" in what:
        res["config5_random_csr"] = []
        for d in (5, 32, 128, 512):
            n = a.nrand if d < 512 else a.nrand // 4
            nnz = n * d
            d_i, d_j, d_a = _capi.DeviceArray(Hh, n + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
            _capi.check(L.b200GenRandomCsr(H, n, n, d, C.c_uint64(20260923 + d), d_i.ptr, d_j.ptr, d_a.ptr))
            plan = C.c_void_p()
            _capi.check(L.b200CsrPlanCreate(H, n, n, C.c_int64(nnz), d_i.ptr, d_j.ptr, C.byref(plan)))
            x, y = _capi.DeviceArray(Hh, n, np.float64), _capi.DeviceArray(Hh, n, np.float64)
            _capi.check(L.b200VecSet(H, C.c_int64(n), C.c_double(1.0), x.ptr))
            lay = [C.c_int() for _ in range(6)]
            best = None
            for lanes in (0, 1, 2, 4, 8, 16, 32):
                if lanes and (lanes > max(2 * d, 1)):
                    continue
                _capi.check(L.b200CsrPlanSetLayout(plan, lanes, 0, 0, 0))
                _capi.check(L.b200CsrPlanGetLayout(plan, *[C.byref(v) for v in lay]))
                ms = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 10)
                r = dict(lanes_req=lanes, lanes=lay[0].value, rows_per_tile=lay[1].value, ms=ms)
                if lanes == 0:
                    auto = r
                if best is None or ms < best["ms"]:
                    best = r
            # L2 hints at the automatic layout: 3 = evict hints only, 7 = + persisting access-policy window on x
            _capi.check(L.b200CsrPlanSetLayout(plan, 0, 0, 0, 0))
            hint_ms = {}
            for hints in (3, 7):
                _capi.check(L.b200CsrPlanSetCacheHints(plan, hints))
                hint_ms[str(hints)] = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 10)
            # row-sum association: default = reference order for every lane count; 1 = FMA + shuffle tree
            _capi.check(L.b200CsrPlanSetCacheHints(plan, 3))
            _capi.check(L.b200CsrPlanSetSummation(plan, 1))
            hint_ms["tree_sum"] = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 10)
            _capi.check(L.b200CsrPlanSetSummation(plan, 0))
            hint_ms["ordered_sum"] = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 10)
            # column-blocked passes (x block stays in the L2): nb = 2, 4, 8; values packed once (timed separately)
            blocked = {}
            if d > 5 and "nb" in what:
                y_ref = y.download()
                for nb in (2, 4, 8):
                    _capi.check(L.b200CsrPlanSetColumnBlocks(H, plan, nb))
                    pack_ms = timed(lambda: _capi.check(L.b200CsrPlanPackValues(H, plan, d_a.ptr)), 3, warm=1)
                    ms = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 10)
                    err = float(np.abs(y.download() - y_ref).max() / max(np.abs(y_ref).max(), 1e-300))
                    blocked[str(nb)] = dict(ms=ms, pack_ms=pack_ms, rel_diff_vs_unblocked=err)
                _capi.check(L.b200CsrPlanSetColumnBlocks(H, plan, 0))
            hint_ms["column_blocks"] = blocked
            alg = nnz * 12 + n * 20
            row = dict(n=n, d=d, nnz=nnz, auto=auto, best=best, hints_ms=hint_ms, algorithmic_bytes=alg, gbs_auto=alg / auto["ms"] / 1e6, gbs_best=alg / best["ms"] / 1e6,
                       frac_auto=alg / auto["ms"] / 1e6 / peak, gflops_auto=(2 * nnz - n) / auto["ms"] / 1e6,
                       note="x gather is random over an n-vector (%d MB): beyond the 126 MB L2, so DRAM traffic exceeds the algorithmic bytes" % (n * 8 // 1000000))
            res["config5_random_csr"].append(row)
            print("config5", row, flush=True)
            L.b200CsrPlanDestroy(plan)
            for o in (d_i, d_j, d_a, x, y):
                o.free()

    if "3" in what:
        n = a.n27
        N = n ** 3
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
        nnz = nnz.value
        d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
        _capi.check(L.b200GenLaplace27(H, n, d_i.ptr, d_j.ptr, d_a.ptr))
        petsc.options_clear()
        petsc.options_insert("-ksp_type cg -pc_type ilu -ksp_rtol 1e-8")
        A = petsc.Mat.create(m=N, n=N, M=N, N=N, comm=petsc.COMM_SELF)
        A.set_csr_device(d_i.ptr, d_j.ptr, d_a.ptr)
        for o in (d_i, d_j, d_a):
            o.free()
        x, b = A.create_vecs()
        u = x.duplicate(); u.set(1.0); A.mult(u, b)
        spmv_ms = timed(lambda: A.mult(u, x), 20)
        ksp = petsc.KSP.create(petsc.COMM_SELF)
        ksp.set_operators(A); ksp.set_residual_history(); ksp.set_from_options()
        t0 = time.time()
        pc = ksp.get_pc()
        ksp.solve(b, x)            # includes PCSetUp: host symbolic + level schedule, device numeric
        _capi.check(L.b200Synchronize(H))
        first = time.time() - t0
        its = ksp.its()
        t = _capi.Timer(Hh)
        t.start(); ksp.solve(b, x); t.stop()
        solve_ms = t.ms()
        y = x.duplicate()
        pcapply_ms = timed(lambda: pc.apply(b, y), 10)
        err = float(np.abs(x.array() - 1.0).max())
        alg_spmv = nnz * 12 + N * 20
        alg_sptrsv = nnz * 12 + N * (4 + 4 + 4 + 4 + 8 * 3)
        res["config3_cg_ilu0_27pt"] = dict(n=n, rows=N, nnz=nnz, iterations=its, reason=ksp.reason(), max_error=err, first_solve_incl_setup_s=first, solve_ms=solve_ms,
                                           ms_per_iteration=solve_ms / max(its, 1), iterations_per_sec=its / (solve_ms * 1e-3), spmv_ms=spmv_ms, spmv_gbs=alg_spmv / spmv_ms / 1e6,
                                           pcapply_ilu_ms=pcapply_ms, sptrsv_gbs=alg_sptrsv / pcapply_ms / 1e6, sptrsv_frac_of_peak=alg_sptrsv / pcapply_ms / 1e6 / peak,
                                           levels=3 * n + 4 * (n - 1) - 2)
        print("config3", res["config3_cg_ilu0_27pt"], flush=True)
        ksp.destroy(); A.destroy()

    if "4" in what:
        # the per-rank block of config 4 (GMRES(30) + PCBJACOBI/ILU(0), 7-point): ILU(0) of the 7-point operator, PCApply timing
        n = a.n7
        N = n ** 3
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace7Nnz(n, n, n, C.c_int64(0), C.c_int64(N), C.byref(nnz)))
        nnz = nnz.value
        d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
        _capi.check(L.b200GenLaplace7(H, n, n, n, C.c_int64(0), C.c_int64(N), d_i.ptr, d_j.ptr, d_a.ptr))
        petsc.options_clear()
        petsc.options_insert("-ksp_type gmres -pc_type ilu -ksp_rtol 1e-8 -ksp_max_it 60")
        A = petsc.Mat.create(m=N, n=N, M=N, N=N, comm=petsc.COMM_SELF)
        A.set_csr_device(d_i.ptr, d_j.ptr, d_a.ptr)
        for o in (d_i, d_j, d_a):
            o.free()
        x, b = A.create_vecs()
        u = x.duplicate(); u.set(1.0); A.mult(u, b)
        ksp = petsc.KSP.create(petsc.COMM_SELF)
        ksp.set_operators(A); ksp.set_from_options()
        t0 = time.time()
        ksp.solve(b, x)
        _capi.check(L.b200Synchronize(H))
        first = time.time() - t0
        its = ksp.its()
        t = _capi.Timer(Hh)
        t.start(); ksp.solve(b, x); t.stop()
        pc = ksp.get_pc()
        y = x.duplicate()
        pcapply_ms = timed(lambda: pc.apply(b, y), 10)
        alg_sptrsv = nnz * 12 + N * (4 + 4 + 4 + 4 + 8 * 3)
        res["config4_block_gmres_ilu0_7pt"] = dict(n=n, rows=N, nnz=nnz, iterations=its, reason=ksp.reason(), first_solve_incl_setup_s=first, solve_ms=t.ms(),
                                                   ms_per_iteration=t.ms() / max(its, 1), pcapply_ilu_ms=pcapply_ms, levels=3 * (n - 1) + 1,
                                                   sptrsv_gbs=alg_sptrsv / pcapply_ms / 1e6, sptrsv_frac_of_peak=alg_sptrsv / pcapply_ms / 1e6 / peak)
        print("config4", res["config4_block_gmres_ilu0_7pt"], flush=True)
        ksp.destroy(); A.destroy()

    if "p" in what:
        # fused-reduction CG (KSPPIPECG: 1 synchronisation per iteration) against KSPCG (3) on the 27-point operator, Jacobi
        res["pipecg_vs_cg_27pt"] = []
        for n in (48, 128, 256):
            N = n ** 3
            nnz = C.c_int64()
            _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
            nnz = nnz.value
            d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
            _capi.check(L.b200GenLaplace27(H, n, d_i.ptr, d_j.ptr, d_a.ptr))
            A = petsc.Mat.create(m=N, n=N, M=N, N=N, comm=petsc.COMM_SELF)
            A.set_csr_device(d_i.ptr, d_j.ptr, d_a.ptr)
            for o in (d_i, d_j, d_a):
                o.free()
            x, b = A.create_vecs()
            u = x.duplicate(); u.set(1.0); A.mult(u, b)
            row = dict(n=n, rows=N, nnz=nnz)
            for kt in ("cg", "pipecg", "pipecg_unfused"):
                petsc.options_clear()
                petsc.options_insert("-ksp_type %s -pc_type jacobi -ksp_rtol 1e-8%s" % (kt.split("_")[0], " -ksp_pipecg_b200_fuse_update 0" if kt.endswith("unfused") else ""))
                ksp = petsc.KSP.create(petsc.COMM_SELF)
                ksp.set_operators(A); ksp.set_from_options()
                ksp.solve(b, x)
                t = _capi.Timer(Hh)
                t.start(); ksp.solve(b, x); t.stop()
                row[kt] = dict(iterations=ksp.its(), reason=ksp.reason(), solve_ms=t.ms(), us_per_iteration=1e3 * t.ms() / max(ksp.its(), 1))
                ksp.destroy()
            res["pipecg_vs_cg_27pt"].append(row)
            print("pipecg", row, flush=True)
            for o in (x, b, u):
                o.destroy()
            A.destroy()
        petsc.options_clear()

    if "tr" in what or "coo" in what:
        # widening rows: transposed product and COO assembly on the 27-point operator (n = --nasm)
        n = a.nasm
        N = n ** 3
        nnz = C.c_int64()
        _capi.check(L.b200GenLaplace27Nnz(n, C.byref(nnz)))
        nnz = nnz.value
        d_i, d_j, d_a = _capi.DeviceArray(Hh, N + 1, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
        _capi.check(L.b200GenLaplace27(H, n, d_i.ptr, d_j.ptr, d_a.ptr))
        x, y = _capi.DeviceArray(Hh, N, np.float64), _capi.DeviceArray(Hh, N, np.float64)
        _capi.check(L.b200VecSet(H, C.c_int64(N), C.c_double(1.0), x.ptr))
        if "tr" in what:
            T = C.c_void_p()
            t0 = time.time()
            _capi.check(L.b200CsrTransposeCreate(H, N, N, C.c_int64(nnz), d_i.ptr, d_j.ptr, C.byref(T)))
            _capi.check(L.b200Synchronize(H))
            create_s = time.time() - t0
            gather_ms = timed(lambda: _capi.check(L.b200CsrTransposeSetValues(H, T, d_a.ptr)), 10)
            out = {}
            for lanes in (0, 1):
                tp = C.c_void_p()
                _capi.check(L.b200CsrTransposeGetPlan(T, C.byref(tp)))
                _capi.check(L.b200CsrPlanSetLayout(tp, lanes, 0, 0, 0))
                out["spmv_ms_lanes%d" % lanes] = timed(lambda: _capi.check(L.b200CsrTransposeSpMV(H, T, x.ptr, None, y.ptr)), 20)
            plan = C.c_void_p()
            _capi.check(L.b200CsrPlanCreate(H, N, N, C.c_int64(nnz), d_i.ptr, d_j.ptr, C.byref(plan)))
            fwd_ms = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 20)
            _capi.check(L.b200CsrPlanSetSummation(plan, 1))
            out["forward_spmv_ms_tree_sum"] = timed(lambda: _capi.check(L.b200CsrSpMV(H, plan, d_a.ptr, x.ptr, y.ptr)), 20)
            L.b200CsrPlanDestroy(plan)
            alg = nnz * 12 + N * 20
            res["transpose_27pt"] = dict(n=n, rows=N, nnz=nnz, create_s=create_s, gather_ms=gather_ms, gather_gbs=nnz * 20 / gather_ms / 1e6, forward_spmv_ms=fwd_ms,
                                         forward_gbs=alg / fwd_ms / 1e6, transposed_gbs_auto=alg / out["spmv_ms_lanes0"] / 1e6, transposed_gbs_parity=alg / out["spmv_ms_lanes1"] / 1e6,
                                         frac_of_peak_auto=alg / out["spmv_ms_lanes0"] / 1e6 / peak, **out)
            print("transpose", res["transpose_27pt"], flush=True)
            L.b200CsrTransposeDestroy(T)
        if "coo" in what:
            # the (i,j,v) triples of the same operator in a shuffled order, resident on the device
            hi = np.empty(N + 1, np.int32)
            _capi.check(L.b200MemcpyDtoH(H, hi.ctypes.data_as(C.c_void_p), d_i.ptr, C.c_size_t(4 * (N + 1))))
            rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(hi))
            cols, vals = d_j.download(), d_a.download()
            perm = np.random.default_rng(3).permutation(nnz)
            c_i, c_j, c_v = _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.int32), _capi.DeviceArray(Hh, nnz, np.float64)
            c_i.upload(rows[perm]); c_j.upload(cols[perm]); c_v.upload(vals[perm])
            del rows, cols, perm
            times = []
            for _ in range(3):
                plan = C.c_void_p()
                t = _capi.Timer(Hh); t.start()
                _capi.check(L.b200CooPlanCreate(H, N, N, C.c_int64(nnz), c_i.ptr, c_j.ptr, C.byref(plan)))
                t.stop(); times.append(t.ms())
                if _ < 2:
                    L.b200CooPlanDestroy(plan)
            out_a = _capi.DeviceArray(Hh, nnz + 1, np.float64)
            set_ms = timed(lambda: _capi.check(L.b200CooSetValues(H, plan, c_v.ptr, 1, out_a.ptr)), 10)
            same = bool(np.array_equal(out_a.download()[:nnz], vals))
            res["coo_27pt"] = dict(n=n, rows=N, coo_n=nnz, prealloc_ms=min(times), prealloc_Mentries_per_s=nnz / min(times) / 1e3, setvalues_ms=set_ms,
                                   setvalues_gbs=nnz * (4 + 4 + 8 + 8) / set_ms / 1e6, values_equal_generator=same)
            print("coo", res["coo_27pt"], flush=True)
            L.b200CooPlanDestroy(plan)

    if "1" in what:
        # ex2 -m 100 -n 100 -ksp_type gmres -pc_type jacobi (BASELINE configs[0]); matrix assembled on the host like ex2 does
        m = 100
        ai = [0]; aj = []; aa = []
        # build CSR directly
        for Ii in range(m * m):
            i, j = divmod(Ii, m)
            if i > 0: aj.append(Ii - m); aa.append(-1.0)
            if j > 0: aj.append(Ii - 1); aa.append(-1.0)
            aj.append(Ii); aa.append(4.0)
            if j < m - 1: aj.append(Ii + 1); aa.append(-1.0)
            if i < m - 1: aj.append(Ii + m); aa.append(-1.0)
            ai.append(len(aj))
        A = petsc.Mat.from_csr(np.array(ai), np.array(aj), np.array(aa))
        x, b = A.create_vecs()
        u = x.duplicate(); u.set(1.0); A.mult(u, b)
        petsc.options_clear()
        petsc.options_insert("-ksp_type gmres -pc_type jacobi -ksp_rtol %r" % (1e-2 / 10201))
        ksp = petsc.KSP.create(petsc.COMM_SELF)
        ksp.set_operators(A); ksp.set_from_options()
        ksp.solve(b, x)
        t = _capi.Timer(Hh)
        t.start(); ksp.solve(b, x); t.stop()
        res["config1_ex2_100x100"] = dict(iterations=ksp.its(), reason=ksp.reason(), rnorm=ksp.rnorm(), solve_ms=t.ms(), us_per_iteration=1e3 * t.ms() / ksp.its(),
                                           error_norm=float(np.linalg.norm(x.array() - 1.0)), reference="719 its, residual 4.918891918633e-06, error 0.00920721 (SURVEY 6)")
        print("config1", res["config1_ex2_100x100"], flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
