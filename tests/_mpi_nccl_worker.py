"""N-rank GPU check of mpiaijb200/mpib200 over NCCL (one process per GPU, launched by torchrun):
MatMult (halo on the second stream + diag SpMV + compressed off-diag multadd), VecMDot/VecNorm all-reduces,
KSPGMRES + PCBJACOBI(ILU(0)) against the reference golden ex2_2.out and the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from petsc_b200 import _capi
from harness import petsc  # noqa: E402


def sig6(v):
    return float("%.6g" % v)


def gather(a):
    box = [None] * td.get_world_size()
    td.all_gather_object(box, np.asarray(a))
    return np.concatenate(box)


class Checks:
    """Records named checks instead of raising: every rank walks the same sequence of collectives whatever a comparison
    says, so a failing rank cannot leave the others waiting inside an all-reduce.  bench.py --gpus N runs the same checks
    before its timed region and reports the counts (`parity_check`)."""

    def __init__(self):
        self.passed, self.failed = 0, []

    def __call__(self, cond, what):
        if bool(cond):
            self.passed += 1
        else:
            self.failed.append(str(what))

    def merged(self):
        """(passed, failures) summed over ranks."""
        box = [None] * td.get_world_size()
        td.all_gather_object(box, (self.passed, self.failed))
        return sum(b[0] for b in box), [("rank %d: %s" % (r, f)) for r, b in enumerate(box) for f in b[1]]


def main():
    td.init_process_group("gloo")
    rank, size = td.get_rank(), td.get_world_size()
    petsc.initialize(device=int(os.environ.get("LOCAL_RANK", rank)))
    box = [petsc.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    petsc.comm_init(rank, size, box[0])
    ck = Checks()
    run_checks(ck, rank, size)
    passed, failures = ck.merged()
    if rank == 0:
        print("parity checks: %d passed, %d failed" % (passed, len(failures)))
        for f in failures:
            print("FAILED", f)
    td.barrier()
    td.destroy_process_group()
    sys.exit(1 if failures else 0)


def run_checks(ck, rank, size):
    """The communicator (petsc.comm_init) and the gloo process group must exist.  Leaves the options database empty."""
    L = _capi.lib()
    H = petsc.handle()

    class Hh:
        h = H
    rng = np.random.default_rng(5)
    for case, (ai, aj, aa) in (("lap7", O.lap7(12, 10, 4 * size + 1)), ("lap5", O.lap5(31, 17)), ("rand", O.random_csr(1003, 9, 4))):
        n = len(ai) - 1
        rs = O.split_ownership(n, size)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        lai = (ai[r0:r1 + 1] - ai[r0]).astype(np.int32)
        laj = np.ascontiguousarray(aj[ai[r0]:ai[r1]]); laa = np.ascontiguousarray(aa[ai[r0]:ai[r1]])
        xg = rng.uniform(-1, 1, n)
        ref = O.matmult(ai, aj, aa, xg)
        for path in ("host", "device"):
            if path == "host":
                A = petsc.Mat.from_local_csr(lai, laj, laa, n)
            else:
                d_i, d_j, d_a = _capi.DeviceArray(Hh, len(lai), np.int32).upload(lai), _capi.DeviceArray(Hh, len(laj), np.int32).upload(laj), _capi.DeviceArray(Hh, len(laa), np.float64).upload(laa)
                A = petsc.Mat.create(m=r1 - r0, n=r1 - r0, comm=petsc.COMM_WORLD)
                A.set_csr_device(d_i.ptr, d_j.ptr, d_a.ptr)
            ck(A.get_type() == "mpiaijb200" and A.ownership_range() == (r0, r1), "A.get_type() == 'mpiaijb200' and A.ownership_range() == (r0, r1)")
            Ad, Ao, garray = A.mpiaij_blocks()
            oA, oB, og = O.mpiaij_split(lai, laj.astype(np.int64), laa, r0, r1)
            ck(np.array_equal(garray, og), ("np.array_equal(garray, og)", (case, path)))  # IS/index work: bit-exact
            for got, want in zip(Ad.csr_host() + Ao.csr_host(), oA + oB):
                ck(np.array_equal(got, want), ("np.array_equal(got, want)", (case, path)))
            x, y = A.create_vecs()
            ck(x.get_type() == "mpib200" and x.ownership_range() == (r0, r1), "x.get_type() == 'mpib200' and x.ownership_range() == (r0, r1)")
            x.set_array(xg[r0:r1])
            A.mult(x, y)
            yl = y.array()
            # rank-local reference: diagonal block first, then the off-diagonal block starting from y[r] (mpiaij.c:1047-1061)
            lv = xg[og] if len(og) else np.zeros(1)
            want = O.matmultadd(oB[0], oB[1], oB[2], np.ascontiguousarray(lv), O.matmult(oA[0], oA[1], oA[2], np.ascontiguousarray(xg[r0:r1])))
            ck(np.array_equal(yl, want), ("np.array_equal(yl, want)", (case, path)))  # bit-identical to MatMult_MPIAIJ's order
            ck(np.allclose(gather(yl), ref, rtol=1e-13, atol=1e-13), "np.allclose(gather(yl), ref, rtol=1e-13, atol=1e-13)")
            z = y.duplicate()
            A.mult_add(x, y, z)
            ck(np.allclose(gather(z.array()), 2 * ref, rtol=1e-13, atol=1e-13), "np.allclose(gather(z.array()), 2 * ref, rtol=1e-13, atol=1e-13)")
            # reductions
            nv = 5
            V = rng.uniform(-1, 1, (nv, n))
            vs, arr = petsc.duplicate_vecs(x, nv)
            for v, row in zip(vs, V):
                v.set_array(row[r0:r1])
            d = x.mdot(vs)
            ck(np.all(np.abs(d - V @ xg) <= 1e-12 * np.linalg.norm(xg) * np.linalg.norm(V, axis=1)), "np.all(np.abs(d - V @ xg) <= 1e-12 * np.linalg.norm(xg) * np.linalg.norm(V, axis=1))")
            ck(np.isclose(x.norm(), np.linalg.norm(xg), rtol=1e-12), "np.isclose(x.norm(), np.linalg.norm(xg), rtol=1e-12)")
            ck(np.isclose(x.norm(0), np.abs(xg).sum(), rtol=1e-12) and x.norm(3) == np.abs(xg).max(), "np.isclose(x.norm(0), np.abs(xg).sum(), rtol=1e-12) and x.norm(3) == np.abs(xg).max()")
            ck(abs(x.dot(vs[0]) - V[0] @ xg) <= 1e-12 * n, "abs(x.dot(vs[0]) - V[0] @ xg) <= 1e-12 * n")
            x.maxpy(-d, vs)
            xr = xg - V.T @ d
            ck(np.isclose(x.norm(), np.linalg.norm(xr), rtol=1e-11), "np.isclose(x.norm(), np.linalg.norm(xr), rtol=1e-11)")  # fused MAXPY+norm, all-reduced at VecNorm
            dg = y.duplicate(); A.get_diagonal(dg)
            ck(np.array_equal(gather(dg.array()), O.getdiagonal(ai, aj, aa)[0]), "np.array_equal(gather(dg.array()), O.getdiagonal(ai, aj, aa)[0])")
            petsc.destroy_vecs(nv, arr)
            for o in (x, y, z, dg, A):
                o.destroy()
        if rank == 0:
            print("OK matmult", case)

    # ---- ex2_2.out (nsize 2): GMRES + default PC (block Jacobi, ILU(0) on each rank's diagonal block), refine_always
    ai, aj, aa = O.lap5(5, 5)
    n = 25
    rs = O.split_ownership(n, size)
    r0, r1 = int(rs[rank]), int(rs[rank + 1])
    lai = (ai[r0:r1 + 1] - ai[r0]).astype(np.int32)
    A = petsc.Mat.from_local_csr(lai, aj[ai[r0]:ai[r1]], aa[ai[r0]:ai[r1]], n)
    x, b = A.create_vecs()
    u = x.duplicate(); u.set(1.0); A.mult(u, b)
    petsc.options_clear()
    petsc.options_insert("-ksp_gmres_cgs_refinement_type refine_always -ksp_rtol %r" % (1e-2 / 36))
    ksp = petsc.KSP.create(petsc.COMM_WORLD)
    ksp.set_operators(A); ksp.set_residual_history(); ksp.set_from_options()
    ck(ksp.get_pc().get_type() == "bjacobi", "ksp.get_pc().get_type() == 'bjacobi'")
    ksp.solve(b, x)
    hist = ksp.history()
    xs = gather(x.array())
    ox, o = O.ksp_solve("gmres", ai, aj, aa, O.matmult(ai, aj, aa, np.ones(n)), pc="bjacobi", nblocks=size, refine="always", rtol=1e-2 / 36)
    ck(ksp.its() == o["its"] and np.allclose(hist, o["hist"], rtol=1e-10), "ksp.its() == o['its'] and np.allclose(hist, o['hist'], rtol=1e-10)")
    if size == 2:
        ck([sig6(v) for v in hist] == [2.73499, 0.795482, 0.261984, 0.0752998, 0.0230031, 0.00521255, 0.00145783, 0.000277319], "[sig6(v) for v in hist] == [2.73499, 0.795482, 0.261984, 0.0752998, 0.0230031, 0.00521255,")
        ck(ksp.its() == 7 and sig6(np.linalg.norm(xs - 1.0)) == 0.000292349, "ksp.its() == 7 and sig6(np.linalg.norm(xs - 1.0)) == 0.000292349")
    # a larger parallel solve: 7-point, GMRES(30)+Jacobi vs the sequential oracle
    ai, aj, aa = O.lap7(14, 12, 6 * size)
    n = len(ai) - 1
    rs = O.split_ownership(n, size); r0, r1 = int(rs[rank]), int(rs[rank + 1])
    A2 = petsc.Mat.from_local_csr((ai[r0:r1 + 1] - ai[r0]).astype(np.int32), aj[ai[r0]:ai[r1]], aa[ai[r0]:ai[r1]], n)
    x2, b2 = A2.create_vecs()
    u2 = x2.duplicate(); u2.set(1.0); A2.mult(u2, b2)
    for opts, pc in (("-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-9", "jacobi"), ("-ksp_type cg -pc_type bjacobi -ksp_rtol 1e-9", "bjacobi")):
        petsc.options_clear(); petsc.options_insert(opts)
        k2 = petsc.KSP.create(petsc.COMM_WORLD)
        k2.set_operators(A2); k2.set_residual_history(); k2.set_from_options()
        k2.solve(b2, x2)
        ox, o = O.ksp_solve(opts.split()[1], ai, aj, aa, O.matmult(ai, aj, aa, np.ones(n)), pc=pc, nblocks=size, rtol=1e-9)
        h = k2.history()
        ck(abs(k2.its() - o["its"]) <= 1 and np.allclose(h[:25], o["hist"][:25], rtol=1e-8), ("abs(k2.its() - o['its']) <= 1 and np.allclose(h[:25], o['hist'][:25], rtol=1e-8)", (opts, k2.its(), o["its"])))
        k = min(len(h), len(o["hist"]), 25)
        dev = float(np.max(np.abs(h[:k] - o["hist"][:k]))) / o["hist"][0]
        ck(dev <= 1e-12, ("residual history within 1e-12*r0 of the sequential oracle over the first %d iterations" % k, opts, dev))  # north_star tolerance
        ck(np.allclose(gather(x2.array()), 1.0, atol=1e-6), "np.allclose(gather(x2.array()), 1.0, atol=1e-6)")
        k2.destroy()
    if rank == 0:
        print("OK ksp")
    petsc.options_clear()


if __name__ == "__main__":
    main()
