#!/usr/bin/env python
"""Times b200IndexedOp (the kernel behind VecScatter / PetscSF on device vectors, petsc_b200/csrc/sf.cu) with CUDA events against its
HBM roofline.  Algorithmic bytes per entry of bs doubles: 8*bs read + 8*bs written (+ 8*bs read of the destination for a non-REPLACE
op) + 4 per non-contiguous index array (+ 4 per group for the grouped kernel's offsets).

    python tools/sf_bench.py [--n 67108864] [--out profiles/round2_sf_bench.json]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 26)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    L = _capi.lib()
    H = _capi.Handle()
    peak = 6570.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n = a.n
    rng = np.random.default_rng(3)
    perm = rng.permutation(n).astype(np.int32)
    # banded permutation: indices shuffled inside windows of 4096 (a halo / reordering-like gather with locality)
    band = (np.arange(n, dtype=np.int64).reshape(-1, 4096) if n % 4096 == 0 else None)
    if band is not None:
        band = rng.permuted(band, axis=1).reshape(-1).astype(np.int32)
    many = rng.integers(0, n // 4, n).astype(np.int32)          # reduction: ~4 leaves per root
    tm = _capi.Timer(H)
    rows = []

    def run(label, sidx, didx, bs, op, nsrc, ndst, reps=10):
        p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)   # noqa: E731
        plan = C.c_void_p()
        _capi.check(L.b200IndexedPlanCreate(H.h, C.c_int64(n), p(sidx), 0, p(didx), 0, C.byref(plan)))
        info = [C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int(), C.c_int64(), C.c_int64()]
        _capi.check(L.b200IndexedPlanGetInfo(plan, *[C.byref(v) for v in info]))
        d_s = H.zeros(nsrc * bs); d_d = H.zeros(ndst * bs)
        for _ in range(3):
            _capi.check(L.b200IndexedOp(H.h, plan, 0, bs, op, d_s.ptr, d_d.ptr))
        tm.start()
        for _ in range(reps):
            _capi.check(L.b200IndexedOp(H.h, plan, 0, bs, op, d_s.ptr, d_d.ptr))
        tm.stop()
        ms = tm.ms() / reps
        grouped, ng = info[2].value, info[1].value
        units = ng if grouped else n
        byt = n * bs * 8 + units * bs * 8 * (1 if op == 0 and not grouped else 2)
        byt += 0 if info[3].value else 4 * n
        byt += (8 * ng) if grouped else (0 if info[4].value else 4 * n)
        rows.append(dict(case=label, n=n, bs=bs, op=["replace", "sum"][op], grouped=bool(grouped), ms=round(ms, 4), algorithmic_gb=round(byt / 1e9, 3),
                         gbs=round(byt / ms / 1e6, 1), frac_of_measured_peak=round(byt / ms / 1e6 / peak, 3)))
        print(rows[-1], flush=True)
        _capi.check(L.b200IndexedPlanDestroy(H.h, plan)); d_s.free(); d_d.free()

    run("copy (both sides contiguous)", None, None, 1, 0, n, n)
    run("add  (both sides contiguous)", None, None, 1, 1, n, n)
    if band is not None:
        run("gather, indices shuffled in 4096-windows", band, None, 1, 0, n, n)
        run("scatter-add, same indices", None, band, 1, 1, n, n)
    run("gather, random permutation", perm, None, 1, 0, n, n)
    run("reduce-add, ~4 leaves per root (grouped, ordered)", None, many, 1, 1, n, n // 4)
    doc = dict(kernel="b200IndexedOp (sf_scatter_kernel / sf_scatter_grouped_kernel)", peak_gbs=peak, rows=rows)
    if a.out:
        json.dump(doc, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
