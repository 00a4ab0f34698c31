#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own CPU code on seeded inputs.

Needs the reference library that oracle/build_ref.sh builds from /root/reference into baseline/_ref/petsc (--with-mpi=0
--with-cuda=0 --with-debugging=0 COPTFLAGS=-O2, OpenBLAS 0.3.15 for BLAS) and oracle/_ref/ref_driver (built by the same
script from ref_driver.c, the program that calls the reference's public API).  It only runs in the build container; the
fixtures it writes are committed and are what the tests read.

    python oracle/gen_golden.py            # (re)write tests/golden/*.npz
    python oracle/gen_golden.py --check    # re-derive every fixture into a scratch directory and compare it, array by array
                                           # and bit for bit, with the committed file; exit status 1 on any difference

Every case records inputs (or the generator parameters) together with the reference's outputs:
  ref_mult / ref_multadd / ref_diag  -- MatMult_SeqAIJ, MatMultAdd_SeqAIJ, MatGetDiagonal_SeqAIJ (bit-exact pins)
  ref_mdot (with -vec_mdot_use_gemv 0 = in-tree loop dvec2.c:83) / ref_mdot_gemv (default BLAS dgemv path)
  ref_maxpy                           -- VecMAXPY_Seq (bit-exact pin)
  ref_ilusolve / ref_jacobi           -- PCApply_ILU (ILU(0) factor + MatSolve_SeqAIJ_NaturalOrdering), PCApply_Jacobi
  ref_dotnorm                         -- VecDot, VecNorm (BLAS; tolerance pins)
  KSP cases: residual history, iteration count, reason, solution
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

REFLIB = os.path.join(ROOT, "baseline", "_ref", "petsc", "lib")
BLASDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
OUT = os.path.join(ROOT, "tests", "golden")


def build_driver():
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])   # no-op when the library is already there
    exe = os.path.join(HERE, "_ref", "ref_driver")
    if not os.path.exists(exe):
        sys.exit("oracle/_ref/ref_driver missing: oracle/build_ref.sh needs /root/reference (build container only)")
    return exe


def run_case(exe, ai, aj, aa, x, y, V, alpha, opts):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "%s:%s:%s" % (REFLIB, BLASDIR, env.get("LD_LIBRARY_PATH", ""))
    with tempfile.TemporaryDirectory() as d:
        m, nv = len(ai) - 1, V.shape[0]
        for name, arr in (("ai.i32", ai), ("aj.i32", aj), ("aa.f64", aa), ("x.f64", x), ("y.f64", y), ("V.f64", V),
                          ("alpha.f64", alpha)):
            np.ascontiguousarray(arr).tofile(os.path.join(d, name))
        open(os.path.join(d, "meta.txt"), "w").write("%d %d %d\n" % (m, len(aj), nv))
        subprocess.check_call([exe, d] + opts, env=env)
        out = {}
        for f in os.listdir(d):
            if f.startswith("ref_") and f.endswith(".f64"):
                out[f[:-4]] = np.fromfile(os.path.join(d, f), dtype=np.float64)
        if os.path.exists(os.path.join(d, "ref_ksp.txt")):
            its, reason, rnorm, nh = open(os.path.join(d, "ref_ksp.txt")).read().split()
            out["ref_its"] = np.int64(its); out["ref_reason"] = np.int64(reason); out["ref_rnorm"] = np.float64(rnorm)
        return out


def check():
    """Regenerate into a scratch directory and diff against the committed fixtures (bit for bit)."""
    global OUT
    committed = OUT
    bad = 0
    with tempfile.TemporaryDirectory() as scratch:
        OUT = scratch
        main()
        names = sorted(f for f in os.listdir(committed) if f.endswith(".npz"))
        fresh = sorted(f for f in os.listdir(scratch) if f.endswith(".npz"))
        if names != fresh:
            print("fixture sets differ:", sorted(set(names) ^ set(fresh))); bad += 1
        for f in sorted(set(names) & set(fresh)):
            a, b = np.load(os.path.join(committed, f)), np.load(os.path.join(scratch, f))
            if sorted(a.files) != sorted(b.files):
                print(f, "array names differ", sorted(set(a.files) ^ set(b.files))); bad += 1; continue
            for k in a.files:
                x, y = a[k], b[k]
                same = x.shape == y.shape and x.dtype == y.dtype and (x.tobytes() == y.tobytes())
                if not same:
                    print(f, k, "DIFFERS"); bad += 1
    print("gen_golden --check: %d fixtures compared, %d differences" % (len(names), bad))
    return 1 if bad else 0


def main():
    exe = build_driver()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260923)

    def vecs(m, nv):
        return (rng.uniform(-1, 1, m), rng.uniform(-1, 1, m), rng.uniform(-1, 1, (nv, m)), rng.uniform(-1, 1, nv))

    # ---- operator-level cases (small, inputs stored) ----
    ops = {
        "ops_lap5_21x17": (O.lap5(21, 17), 7, {"gen": "lap5", "args": [21, 17]}),
        "ops_lap7_9x8x7": (O.lap7(9, 8, 7), 30, {"gen": "lap7", "args": [9, 8, 7]}),
        "ops_lap27_7": (O.lap27(7), 5, {"gen": "lap27", "args": [7]}),
        "ops_rand_402x9": (O.random_csr(402, 9, 20260923 + 9), 6, {"gen": "stored"}),
    }
    for name, ((ai, aj, aa), nv, meta) in ops.items():
        if meta["gen"] == "stored":
            # make the random matrix diagonally dominant so ILU(0) is well defined
            d, pos = O.getdiagonal(ai, aj, aa)
            aa = aa.copy(); aa[pos] = 12.0
        x, y, V, alpha = vecs(len(ai) - 1, nv)
        sym = ["-icc"] if meta["gen"] != "stored" else []   # the generated Laplacians are symmetric: record PCApply_ICC too
        out = run_case(exe, ai, aj, aa, x, y, V, alpha, ["-vec_mdot_use_gemv", "0"] + sym)
        out2 = run_case(exe, ai, aj, aa, x, y, V, alpha, [])
        out["ref_mdot_gemv"] = out2["ref_mdot"]
        rec = dict(x=x, y=y, V=V, alpha=alpha, gen=meta["gen"], args=np.array(meta.get("args", []), dtype=np.int64), **out)
        if meta["gen"] == "stored":
            rec.update(ai=ai, aj=aj, aa=aa)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "ok")

    # ---- COO assembly (MatSetPreallocationCOO / MatSetValuesCOO): repeats, negative (ignored) indices, sorted and unsorted input
    crng = np.random.default_rng(20260924)
    coo = {
        "coo_rand_60x45": (60, 45, 4000, "random"),      # ~1.5 repeats per entry on average, up to ~7: pins the summation order
        "coo_rowsorted_33": (33, 33, 900, "rowsorted"),   # rows arrive sorted: the row sort is skipped (aij.c:4558)
        "coo_fem_q1_7x6": (56, 56, 0, "fem"),             # Q1 element loop on a 7x6 node grid: 16 entries per element, up to 4 repeats
        "coo_empty_rows": (20, 20, 30, "random"),
    }
    for name, (M, N, n, kind) in coo.items():
        if kind == "fem":
            nx, ny = 7, 8
            ii, jj = [], []
            for ey in range(ny - 1):
                for ex in range(nx - 1):
                    nodes = [ey * nx + ex, ey * nx + ex + 1, (ey + 1) * nx + ex, (ey + 1) * nx + ex + 1]
                    for a_ in nodes:
                        for b_ in nodes:
                            ii.append(a_); jj.append(b_)
            ci, cj = np.array(ii, np.int32), np.array(jj, np.int32)
            n = len(ci)
        else:
            ci = crng.integers(-1, M, n).astype(np.int32)
            cj = crng.integers(-1, N, n).astype(np.int32)
            if kind == "rowsorted":
                ci = np.sort(ci[ci >= 0]).astype(np.int32); cj = cj[:len(ci)]; n = len(ci)
        v1, v2 = crng.uniform(-1, 1, n), crng.uniform(-1, 1, n)
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = "%s:%s:%s" % (REFLIB, BLASDIR, env.get("LD_LIBRARY_PATH", ""))
        with tempfile.TemporaryDirectory() as d:
            for fn, arr in (("coo_i.i32", ci), ("coo_j.i32", cj), ("v1.f64", v1), ("v2.f64", v2)):
                np.ascontiguousarray(arr).tofile(os.path.join(d, fn))
            open(os.path.join(d, "meta_coo.txt"), "w").write("%d %d %d\n" % (M, N, n))
            subprocess.check_call([exe, "-coo", d], env=env)
            rec = dict(M=M, N=N, coo_i=ci, coo_j=cj, v1=v1, v2=v2,
                       ref_ai=np.fromfile(os.path.join(d, "ref_ai.i32"), np.int32), ref_aj=np.fromfile(os.path.join(d, "ref_aj.i32"), np.int32),
                       ref_aa1=np.fromfile(os.path.join(d, "ref_aa1.f64")), ref_aa2=np.fromfile(os.path.join(d, "ref_aa2.f64")))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "n", n, "nnz", len(rec["ref_aj"]))

    # ---- PetscSF local broadcast / reduction: every MPI_Op a VecScatter can ask for + PROD, unit = bs scalars; roots
    # with several leaves pin the order of application (PetscSFSetGraph sorts the leaves by location, sf.c:500)
    srng = np.random.default_rng(20260925)
    sfc = {"sf_rand_bs1": (300, 700, 900, 1), "sf_rand_bs3": (120, 260, 400, 3), "sf_contig_leaves": (64, 200, 200, 1), "sf_empty": (5, 0, 4, 1)}
    for name, (nroots, nleaves, leafspan, bs) in sfc.items():
        local = srng.permutation(leafspan)[:nleaves].astype(np.int32) if name != "sf_contig_leaves" else np.arange(nleaves, dtype=np.int32)
        remote = srng.integers(0, nroots, nleaves).astype(np.int32)
        root, leaf = srng.uniform(-1, 1, nroots * bs), srng.uniform(-1, 1, leafspan * bs)
        rooti, leafi = srng.integers(-1000, 1000, nroots).astype(np.int32), srng.integers(-1000, 1000, leafspan).astype(np.int32)
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = "%s:%s:%s" % (REFLIB, BLASDIR, env.get("LD_LIBRARY_PATH", ""))
        with tempfile.TemporaryDirectory() as d:
            for fn, arr in (("local.i32", local), ("remote.i32", remote), ("root.f64", root), ("leaf.f64", leaf), ("root.i32", rooti), ("leaf.i32", leafi)):
                np.ascontiguousarray(arr).tofile(os.path.join(d, fn))
            open(os.path.join(d, "meta_sf.txt"), "w").write("%d %d %d %d\n" % (nroots, nleaves, leafspan, bs))
            subprocess.check_call([exe, "-sf", d], env=env)
            rec = dict(nroots=nroots, leafspan=leafspan, bs=bs, local=local, remote=remote, root=root, leaf=leaf, rooti=rooti, leafi=leafi)
            for f in sorted(os.listdir(d)):
                if f.startswith("ref_"):
                    rec[f.replace(".f64", "").replace(".i32", "_i32")] = np.fromfile(os.path.join(d, f), np.float64 if f.endswith(".f64") else np.int32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "nleaves", nleaves)

    # ---- KSP cases: matrix by generator, b = A*1, reference options recorded ----
    ksp = {
        # ex2_1.out: -m 5 -n 5 -ksp_gmres_cgs_refinement_type refine_always, default PC (ILU), ex2.c rtol = 1e-2/((m+1)(n+1))
        "ksp_ex2_1": ("lap5", [5, 5], ["-ksp_type", "gmres", "-pc_type", "ilu", "-ksp_gmres_cgs_refinement_type", "refine_always",
                                         "-ksp_rtol", repr(1e-2 / 36)]),
        # BASELINE config 1: ex2 100x100 GMRES(30)+Jacobi, rtol = 1e-2/(101*101)
        "ksp_ex2_100_gmres_jacobi": ("lap5", [100, 100], ["-ksp_type", "gmres", "-pc_type", "jacobi", "-ksp_rtol", repr(1e-2 / 10201)]),
        "ksp_lap5_30_gmres_none": ("lap5", [30, 30], ["-ksp_type", "gmres", "-pc_type", "none"]),
        "ksp_lap5_30_gmres_ilu": ("lap5", [30, 30], ["-ksp_type", "gmres", "-pc_type", "ilu", "-ksp_rtol", "1e-10"]),
        "ksp_lap5_30_gmres_ifneeded": ("lap5", [30, 30], ["-ksp_type", "gmres", "-pc_type", "jacobi", "-ksp_gmres_cgs_refinement_type", "refine_ifneeded", "-ksp_rtol", "1e-8"]),
        "ksp_lap5_30_cg_jacobi": ("lap5", [30, 30], ["-ksp_type", "cg", "-pc_type", "jacobi", "-ksp_rtol", "1e-8"]),
        "ksp_lap7_12_gmres_jacobi": ("lap7", [12, 12, 12], ["-ksp_type", "gmres", "-pc_type", "jacobi", "-ksp_rtol", "1e-8"]),
        "ksp_lap7_12_gmres5_jacobi": ("lap7", [12, 11, 10], ["-ksp_type", "gmres", "-ksp_gmres_restart", "5", "-pc_type", "jacobi", "-ksp_rtol", "1e-8"]),
        # BASELINE config 3 shape: 27-pt, CG + ILU(0)
        "ksp_lap27_10_cg_ilu": ("lap27", [10], ["-ksp_type", "cg", "-pc_type", "ilu", "-ksp_rtol", "1e-8"]),
        "ksp_lap27_10_gmres_ilu": ("lap27", [10], ["-ksp_type", "gmres", "-pc_type", "ilu", "-ksp_rtol", "1e-8"]),
        # ICC(0) (ex2's default PC because ex2 marks its matrix symmetric), oracle first
        "ksp_lap27_10_cg_icc": ("lap27", [10], ["-ksp_type", "cg", "-pc_type", "icc", "-ksp_rtol", "1e-8"]),
        "ksp_lap5_30_cg_icc": ("lap5", [30, 30], ["-ksp_type", "cg", "-pc_type", "icc", "-ksp_rtol", "1e-8"]),
        "ksp_lap7_12_gmres_icc": ("lap7", [12, 11, 10], ["-ksp_type", "gmres", "-pc_type", "icc", "-ksp_rtol", "1e-8"]),
        # single-reduction CG, oracle first
        "ksp_lap5_30_pipecg_jacobi": ("lap5", [30, 30], ["-ksp_type", "pipecg", "-pc_type", "jacobi", "-ksp_rtol", "1e-8"]),
        "ksp_lap27_10_pipecg_icc": ("lap27", [10], ["-ksp_type", "pipecg", "-pc_type", "icc", "-ksp_rtol", "1e-8"]),
        # pipelined GMRES (one reduction per iteration, used one iteration later), with restarts and without
        "ksp_lap5_30_pgmres_jacobi": ("lap5", [30, 30], ["-ksp_type", "pgmres", "-pc_type", "jacobi", "-ksp_rtol", "1e-8"]),
        "ksp_lap7_12_pgmres_ilu": ("lap7", [12, 11, 10], ["-ksp_type", "pgmres", "-pc_type", "ilu", "-ksp_rtol", "1e-8"]),
        "ksp_lap5_30_pgmres5_none": ("lap5", [30, 30], ["-ksp_type", "pgmres", "-ksp_gmres_restart", "5", "-pc_type", "none", "-ksp_rtol", "1e-3"]),
    }
    for name, (gen, args, opts) in ksp.items():
        ai, aj, aa = getattr(O, gen)(*args)
        m = len(ai) - 1
        z = np.zeros(m)
        out = run_case(exe, ai, aj, aa, z + 1, z, np.zeros((1, m)), np.zeros(1), ["-solve"] + opts)
        keep = {k: v for k, v in out.items() if k in ("ref_hist", "ref_sol", "ref_its", "ref_reason", "ref_rnorm")}
        np.savez_compressed(os.path.join(OUT, name + ".npz"), gen=gen, args=np.array(args, dtype=np.int64), opts=np.array(opts), **keep)
        print(name, "its", int(out["ref_its"]), "reason", int(out["ref_reason"]), "rnorm %.12e" % float(out["ref_rnorm"]))


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(check())
    main()
