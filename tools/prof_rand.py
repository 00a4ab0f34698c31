#!/usr/bin/env python
"""Launch the random-CSR SpMV a few times (ncu target): python tools/prof_rand.py <n> <d> [lanes] [hints]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petsc_b200 import _capi  # noqa: E402

n, d = int(sys.argv[1]), int(sys.argv[2])
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hints = int(sys.argv[4]) if len(sys.argv) > 4 else 2
H = _capi.Handle(0)
L = _capi.lib()
d_i, d_j, d_a = H.empty(n + 1, np.int32), H.empty(n * d, np.int32), H.empty(n * d)
_capi.check(L.b200GenRandomCsr(H.h, n, n, d, C.c_uint64(20260923 + d), d_i.ptr, d_j.ptr, d_a.ptr))
plan = H.csr_plan(n, n, n * d, d_i, d_j)
H.csr_plan_set_layout(plan, lanes, 0, 0, 0)
H.csr_plan_set_hints(plan, hints)
x, y = H.empty(n), H.empty(n)
_capi.check(L.b200VecSet(H.h, C.c_int64(n), C.c_double(1.0), x.ptr))
t = _capi.Timer(H)
for _ in range(3):
    H.spmv(plan, d_a, x, y)
t.start()
for _ in range(5):
    H.spmv(plan, d_a, x, y)
t.stop()
print("layout", H.csr_plan_layout(plan), "ms", t.ms() / 5, "GB/s", (n * d * 12 + n * 20) / (t.ms() / 5) / 1e6)
