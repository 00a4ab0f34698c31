"""Multi-rank parity of the REAL-PETSc path: mpiaijb200 / mpib200 (petsc_plugin) over NCCL ranks against the oracle.

Each rank writes its row block of a few small operators (oracle inputs), runs `b200_driver -parity <dir>` -- a PETSc program
that builds the mpiaijb200 matrix through the plugin's creators and calls MatMult / MatMultAdd / MatMultTranspose /
MatGetDiagonal / PCApplyBAorAB(PCJACOBI) / VecMDot / VecNorm / VecMAXPY / KSPSolve through PETSc's public API -- and then
compares what came back with the oracle (reference restatement):
  * garray and both blocks' CSR arrays: index-exact / bit-exact against O.mpiaij_split (MatSetUpMultiply_MPIAIJ, mmaij.c:8-126)
  * MatMult: bit-exact against the rank-local reference order (diagonal block, then off-diagonal block continuing y[r],
    mpiaij.c:1047-1061); fused Jacobi bit-identical to the unfused composition
  * reductions: 1e-12; KSP residual histories: 1e-12 * r0 over the first cycle; ex2_2.out digits at 2 ranks
Used by bench.py --gpus N (before the timed region: `parity_check`) and by tests (1 rank on the 1-GPU test box).
Only the checker side imports the oracle; the driver under test does not.
"""
import os
import shutil

import numpy as np


def sig6(v):
    return float("%.6g" % v)


EX2_2_HIST = [2.73499, 0.795482, 0.261984, 0.0752998, 0.0230031, 0.00521255, 0.00145783, 0.000277319]   # src/ksp/ksp/tutorials/output/ex2_2.out


def cases(O, size):
    rng = np.random.default_rng(5)
    out = []
    for name, (ai, aj, aa) in (("lap7", O.lap7(12, 10, 4 * size + 1)), ("lap5", O.lap5(31, 17)), ("rand", O.random_csr(1003, 9, 4))):
        n = len(ai) - 1
        for dev in (0, 1):
            out.append(dict(name="%s_%s" % (name, "dev" if dev else "host"), ai=ai, aj=aj, aa=aa, x=rng.uniform(-1, 1, n), V=rng.uniform(-1, 1, (5, n)), dev=dev, solve=None))
    ai, aj, aa = O.lap5(5, 5)
    out.append(dict(name="ex2_2", ai=ai, aj=aj, aa=aa, x=rng.uniform(-1, 1, 25), V=rng.uniform(-1, 1, (2, 25)), dev=0,
                    solve=dict(opts="-ksp_type gmres -ksp_gmres_cgs_refinement_type refine_always -ksp_rtol %r -pc_type bjacobi -sub_pc_type ilu -sub_pc_factor_mat_solver_type b200" % (1e-2 / 36),
                               ksp="gmres", kw=dict(pc="bjacobi", nblocks=size, refine="always", rtol=1e-2 / 36), golden=True)))
    ai, aj, aa = O.lap7(14, 12, 6 * size)
    n = len(ai) - 1
    out.append(dict(name="lap7_gmres_jacobi", ai=ai, aj=aj, aa=aa, x=rng.uniform(-1, 1, n), V=rng.uniform(-1, 1, (2, n)), dev=1,
                    solve=dict(opts="-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-9", ksp="gmres", kw=dict(pc="jacobi", rtol=1e-9))))
    out.append(dict(name="lap7_cg_bjacobi", ai=ai, aj=aj, aa=aa, x=rng.uniform(-1, 1, n), V=rng.uniform(-1, 1, (2, n)), dev=1,
                    solve=dict(opts="-ksp_type cg -pc_type bjacobi -sub_pc_type ilu -sub_pc_factor_mat_solver_type b200 -ksp_rtol 1e-9", ksp="cg", kw=dict(pc="bjacobi", nblocks=size, rtol=1e-9))))
    return out


def write(dirpath, O, rank, size):
    cs = cases(O, size)
    for c in cs:
        ai, aj, aa = c["ai"], c["aj"], c["aa"]
        n = len(ai) - 1
        rs = O.split_ownership(n, size)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        d = os.path.join(dirpath, "case_" + c["name"], "rank%d" % rank)
        os.makedirs(d, exist_ok=True)
        lai = (ai[r0:r1 + 1] - ai[r0]).astype(np.int32)
        laj = np.ascontiguousarray(aj[ai[r0]:ai[r1]]).astype(np.int32)
        laa = np.ascontiguousarray(aa[ai[r0]:ai[r1]])
        lai.tofile(os.path.join(d, "ai.i32")); laj.tofile(os.path.join(d, "aj.i32")); laa.tofile(os.path.join(d, "aa.f64"))
        np.ascontiguousarray(c["x"][r0:r1]).tofile(os.path.join(d, "x.f64"))
        np.ascontiguousarray(c["V"][:, r0:r1]).tofile(os.path.join(d, "V.f64"))
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write("%d %d %d %d %d %d %d\n" % (r1 - r0, n, r0, len(laj), c["V"].shape[0], c["dev"], 1 if c["solve"] else 0))
            if c["solve"]:
                f.write(c["solve"]["opts"] + "\n")
        c.update(r0=r0, r1=r1, lai=lai, laj=laj, laa=laa, dir=d)
    if rank == 0:
        open(os.path.join(dirpath, "cases.txt"), "w").write("\n".join(c["name"] for c in cs) + "\n")
    return cs


def check(cs, O, rank, size, ck, gather):
    """ck(cond, what) records; gather(array) concatenates over ranks (identity for one rank)."""
    for c in cs:
        d, r0, r1, name = c["dir"], c["r0"], c["r1"], c["name"]
        m = r1 - r0

        def rd(fn, dt=np.float64):
            return np.fromfile(os.path.join(d, fn), dtype=dt)
        oA, oB, og = O.mpiaij_split(c["lai"], c["laj"].astype(np.int64), c["laa"], r0, r1)
        ck(np.array_equal(rd("out_garray.i32", np.int32), og), (name, "garray index-exact (mmaij.c:25-61)"))
        for pre, blk in (("out_A", oA), ("out_B", oB)):
            ck(np.array_equal(rd(pre + "_i.i32", np.int32), blk[0]) and np.array_equal(rd(pre + "_j.i32", np.int32), blk[1]) and np.array_equal(rd(pre + "_a.f64"), blk[2]),
               (name, pre + " block CSR bit-exact"))
        xg, V = c["x"], c["V"]
        ref = O.matmult(c["ai"], c["aj"], c["aa"], xg)
        lv = xg[og] if len(og) else np.zeros(1)
        want = O.matmultadd(oB[0], oB[1], oB[2], np.ascontiguousarray(lv), O.matmult(oA[0], oA[1], oA[2], np.ascontiguousarray(xg[r0:r1])))
        y = rd("out_mult.f64")
        ck(np.array_equal(y, want), (name, "MatMult bit-identical to MatMult_MPIAIJ's order (mpiaij.c:1047-1061)"))
        ck(np.allclose(gather(y), ref, rtol=1e-13, atol=1e-13), (name, "MatMult vs sequential MatMult_SeqAIJ"))
        ck(np.allclose(gather(rd("out_multadd.f64")), 2 * ref, rtol=1e-13, atol=1e-13), (name, "MatMultAdd"))
        reft = O.matmulttranspose(c["ai"], c["aj"], c["aa"], xg)
        if True:
            ck(np.allclose(gather(rd("out_multtr.f64")), reft, rtol=1e-13, atol=1e-13), (name, "MatMultTranspose (reverse scatter, mpiaij.c:1086-1097)"))
        ck(np.array_equal(gather(rd("out_diag.f64")), O.getdiagonal(c["ai"], c["aj"], c["aa"])[0]), (name, "MatGetDiagonal"))
        jf, ju = rd("out_jacobi_fused.f64"), rd("out_jacobi_unfused.f64")
        ck(np.array_equal(jf, ju), (name, "fused MatMult+PCJACOBI bit-identical to the unfused composition"))
        dg = O.getdiagonal(c["ai"], c["aj"], c["aa"])[0][r0:r1]
        ck(np.array_equal(ju, want * (1.0 / dg)), (name, "PCApply_Jacobi o MatMult vs oracle"))
        dots = rd("out_mdot.f64")
        ck(np.all(np.abs(dots - V @ xg) <= 1e-12 * np.linalg.norm(xg) * np.linalg.norm(V, axis=1)), (name, "VecMDot all-reduced, 1e-12"))
        sc = rd("out_scalars.f64")
        ck(np.isclose(sc[0], np.linalg.norm(xg), rtol=1e-12) and np.isclose(sc[1], np.abs(xg).sum(), rtol=1e-12) and sc[2] == np.abs(xg).max(), (name, "VecNorm 2/1/inf"))
        ck(abs(sc[3] - V[0] @ xg) <= 1e-12 * len(xg), (name, "VecDot"))
        xr = xg - V.T @ dots
        ck(np.isclose(sc[4], np.linalg.norm(xr), rtol=1e-11), (name, "fused VecMAXPY+VecNorm, all-reduced at VecNorm"))
        Vl = np.ascontiguousarray(V[:, r0:r1])
        ck(np.array_equal(rd("out_maxpy.f64"), O.vecmaxpy(xg[r0:r1].copy(), -dots, [Vl[j] for j in range(len(Vl))])), (name, "VecMAXPY bit-exact (dvec2.c:658-693 association)"))
        ck(np.isclose(sc[5], ref.sum(), rtol=1e-12, atol=1e-12), (name, "VecSum all-reduced"))
        if c["solve"]:
            s = c["solve"]
            hist, info, sol = rd("out_hist.f64"), rd("out_ksp.f64"), gather(rd("out_sol.f64"))
            b = O.matmult(c["ai"], c["aj"], c["aa"], np.ones(len(xg)))
            ox, o = O.ksp_solve(s["ksp"], c["ai"], c["aj"], c["aa"], b, **s["kw"])
            k = min(len(hist), len(o["hist"]), 25)
            dev = float(np.max(np.abs(hist[:k] - o["hist"][:k]))) / o["hist"][0]
            ck(abs(int(info[0]) - o["its"]) <= 1 and dev <= 1e-12, (name, "KSP history within 1e-12*r0 of the oracle over %d iterations" % k, int(info[0]), o["its"], dev))
            ck(np.allclose(sol, 1.0, atol=1e-3 if s.get("golden") else 1e-6), (name, "solution"))
            if s.get("golden") and size == 2:
                ck([sig6(v) for v in hist] == EX2_2_HIST and int(info[0]) == 7 and sig6(np.linalg.norm(sol - 1.0)) == 0.000292349, (name, "ex2_2.out reproduced to all printed digits"))


def cleanup(dirpath):
    shutil.rmtree(dirpath, ignore_errors=True)
