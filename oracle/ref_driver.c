/*
 * ref_driver.c -- fixture generator: runs the REFERENCE's own MATSEQAIJ/VECSEQ/KSP/PC CPU code (public PETSc API only)
 * on inputs written by oracle/gen_golden.py and dumps the outputs as raw little-endian arrays.
 *
 * Test infrastructure only.  It is compiled against a PETSc build of /root/reference when one exists in the build
 * container (PETSC_DIR/PETSC_ARCH; see gen_golden.py) and is never shipped or used at run time: the committed
 * tests/golden/ fixtures are what the tests read.
 *
 * usage: ref_driver <dir> [petsc options...]
 *   reads  <dir>/ai.i32 aj.i32 aa.f64 x.f64 y.f64 V.f64 alpha.f64 meta.txt("m nnz nv")
 *   writes <dir>/ref_mult.f64 ref_multadd.f64 ref_diag.f64 ref_mdot.f64 ref_maxpy.f64 ref_ilusolve.f64
 *          ref_jacobi.f64 ref_dotnorm.f64 ref_hist.f64 ref_sol.f64 ref_ksp.txt
 */
#include <petscksp.h>
#include <petscsf.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* timing mode: ref_driver -bench7 <n> [petsc options]
   builds the 7-point n^3 operator in memory (oracle generator; input data only), then times the REFERENCE's own
   MatMult_SeqAIJ and KSPSolve (options from the command line) and prints one line:
   REFBENCH n=<n> rows=<N> nnz=<nnz> matmult_s=<avg seconds> ksp_its=<its> ksp_s=<seconds> reason=<r> rnorm=<r> */
static PetscErrorCode bench7(int n)
{
  PetscInt     N = n * n * n, its;
  int64_t      nnz = ora_lap7_nnz(n, n, n);
  PetscInt    *ai, *aj;
  PetscScalar *aa;
  Mat          A;
  Vec          u, b, x;
  KSP          ksp;
  double       t0, tm, tk;
  PetscReal    rnorm;
  KSPConvergedReason reason;

  PetscFunctionBeginUser;
  PetscCall(PetscMalloc3(N + 1, &ai, nnz, &aj, nnz, &aa));
  ora_lap7(n, n, n, ai, aj, aa);
  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, N, N, ai, aj, aa, &A));
  PetscCall(MatCreateVecs(A, &u, &b));
  PetscCall(VecDuplicate(u, &x));
  PetscCall(VecSet(u, 1.0));
  PetscCall(MatMult(A, u, b));
  PetscCall(PetscTime(&t0));
  for (int r = 0; r < 5; r++) PetscCall(MatMult(A, u, x));
  PetscCall(PetscTime(&tm));
  tm = (tm - t0) / 5;
  PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
  PetscCall(KSPSetOperators(ksp, A, A));
  PetscCall(KSPSetFromOptions(ksp));
  PetscCall(KSPSetUp(ksp));
  PetscCall(PetscTime(&t0));
  PetscCall(KSPSolve(ksp, b, x));
  PetscCall(PetscTime(&tk));
  tk -= t0;
  PetscCall(KSPGetIterationNumber(ksp, &its));
  PetscCall(KSPGetConvergedReason(ksp, &reason));
  PetscCall(KSPGetResidualNorm(ksp, &rnorm));
  printf("REFBENCH n=%d rows=%d nnz=%lld matmult_s=%.6f ksp_its=%d ksp_s=%.6f reason=%d rnorm=%.12e\n", n, (int)N, (long long)nnz, tm, (int)its, tk, (int)reason, (double)rnorm);
  PetscCall(KSPDestroy(&ksp));
  PetscCall(VecDestroy(&u)); PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&x));
  PetscCall(MatDestroy(&A));
  PetscCall(PetscFree3(ai, aj, aa));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static void *rd(const char *dir, const char *name, size_t bytes);
static void  wr(const char *dir, const char *name, const void *buf, size_t bytes);

/* -coo <dir>: MatSetPreallocationCOO + MatSetValuesCOO(INSERT) + MatSetValuesCOO(ADD) through the public API; the CSR the
   reference builds and both value arrays are written back (the values pin the reference's summation order of repeats) */
static PetscErrorCode coo_case(const char *dir)
{
  int                M, N;
  long               n;
  PetscInt          *ci, *cj;
  const PetscInt    *ia, *ja;
  PetscScalar       *v1, *v2;
  const PetscScalar *a;
  PetscInt           nr;
  PetscBool          done;
  Mat                A;
  char               p[4096];
  FILE              *f;

  PetscFunctionBeginUser;
  snprintf(p, sizeof p, "%s/meta_coo.txt", dir);
  f = fopen(p, "r");
  if (!f || fscanf(f, "%d %d %ld", &M, &N, &n) != 3) exit(2);
  fclose(f);
  ci = (PetscInt *)rd(dir, "coo_i.i32", sizeof(PetscInt) * (size_t)n);
  cj = (PetscInt *)rd(dir, "coo_j.i32", sizeof(PetscInt) * (size_t)n);
  v1 = (PetscScalar *)rd(dir, "v1.f64", sizeof(PetscScalar) * (size_t)n);
  v2 = (PetscScalar *)rd(dir, "v2.f64", sizeof(PetscScalar) * (size_t)n);
  PetscCall(MatCreate(PETSC_COMM_SELF, &A));
  PetscCall(MatSetSizes(A, M, N, M, N));
  PetscCall(MatSetType(A, MATSEQAIJ));
  PetscCall(MatSetPreallocationCOO(A, (PetscCount)n, ci, cj));
  PetscCall(MatSetValuesCOO(A, v1, INSERT_VALUES));
  PetscCall(MatGetRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ia, &ja, &done));
  wr(dir, "ref_ai.i32", ia, sizeof(PetscInt) * (size_t)(M + 1));
  wr(dir, "ref_aj.i32", ja, sizeof(PetscInt) * (size_t)ia[M]);
  PetscCall(MatSeqAIJGetArrayRead(A, &a));
  wr(dir, "ref_aa1.f64", a, sizeof(PetscScalar) * (size_t)ia[M]);
  PetscCall(MatSeqAIJRestoreArrayRead(A, &a));
  PetscCall(MatSetValuesCOO(A, v2, ADD_VALUES));
  PetscCall(MatSeqAIJGetArrayRead(A, &a));
  wr(dir, "ref_aa2.f64", a, sizeof(PetscScalar) * (size_t)ia[M]);
  PetscCall(MatSeqAIJRestoreArrayRead(A, &a));
  PetscCall(MatRestoreRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ia, &ja, &done));
  PetscCall(MatDestroy(&A));
  free(ci); free(cj); free(v1); free(v2);
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* -sf <dir>: a process-local PetscSF (nroots, leaves at local[], leaf k attached to root remote[k]) through the public API:
   PetscSFBcast / PetscSFReduce with each MPI_Op on PetscScalar units of bs entries, and SUM / MAX on PetscInt; every run starts
   from the input root/leaf data.  Outputs ref_bcast_<op>.f64 (leaf data) / ref_reduce_<op>.f64 (root data), *_i32 for PetscInt. */
static PetscErrorCode sf_case(const char *dir)
{
  int           nroots, nleaves, leafspan, bs;
  PetscInt     *local, *rem, *ri, *li, *ri0, *li0;
  PetscSFNode  *remote;
  PetscScalar  *root0, *leaf0, *root, *leaf;
  PetscSF       sf;
  MPI_Datatype  unit;
  char          p[4096];
  FILE         *f;
  const struct {
    MPI_Op      op;
    const char *name;
  } ops[] = {{MPI_REPLACE, "replace"}, {MPI_SUM, "sum"}, {MPI_PROD, "prod"}, {MPI_MAX, "max"}, {MPI_MIN, "min"}};

  PetscFunctionBeginUser;
  snprintf(p, sizeof p, "%s/meta_sf.txt", dir);
  f = fopen(p, "r");
  if (!f || fscanf(f, "%d %d %d %d", &nroots, &nleaves, &leafspan, &bs) != 4) exit(2);
  fclose(f);
  local = (PetscInt *)rd(dir, "local.i32", sizeof(PetscInt) * (size_t)nleaves);
  rem   = (PetscInt *)rd(dir, "remote.i32", sizeof(PetscInt) * (size_t)nleaves);
  root0 = (PetscScalar *)rd(dir, "root.f64", sizeof(PetscScalar) * (size_t)nroots * bs);
  leaf0 = (PetscScalar *)rd(dir, "leaf.f64", sizeof(PetscScalar) * (size_t)leafspan * bs);
  ri0   = (PetscInt *)rd(dir, "root.i32", sizeof(PetscInt) * (size_t)nroots);
  li0   = (PetscInt *)rd(dir, "leaf.i32", sizeof(PetscInt) * (size_t)leafspan);
  PetscCall(PetscMalloc1(nleaves, &remote));
  for (int k = 0; k < nleaves; k++) {
    remote[k].rank  = 0;
    remote[k].index = rem[k];
  }
  PetscCall(PetscMalloc4((size_t)nroots * bs, &root, (size_t)leafspan * bs, &leaf, nroots, &ri, leafspan, &li));
  PetscCall(PetscSFCreate(PETSC_COMM_SELF, &sf));
  PetscCall(PetscSFSetGraph(sf, nroots, nleaves, local, PETSC_COPY_VALUES, remote, PETSC_COPY_VALUES));
  PetscCall(PetscSFSetUp(sf));
  if (bs > 1) { /* the unit VecScatterCreate makes for block index sets (vscat.c:1087-1090) */
    PetscCallMPI(MPI_Type_contiguous(bs, MPIU_SCALAR, &unit));
    PetscCallMPI(MPI_Type_commit(&unit));
  } else unit = MPIU_SCALAR;
  for (int o = 0; o < 5; o++) {
    char name[64];
    memcpy(root, root0, sizeof(PetscScalar) * (size_t)nroots * bs);
    memcpy(leaf, leaf0, sizeof(PetscScalar) * (size_t)leafspan * bs);
    PetscCall(PetscSFBcastBegin(sf, unit, root, leaf, ops[o].op));
    PetscCall(PetscSFBcastEnd(sf, unit, root, leaf, ops[o].op));
    snprintf(name, sizeof name, "ref_bcast_%s.f64", ops[o].name);
    wr(dir, name, leaf, sizeof(PetscScalar) * (size_t)leafspan * bs);
    memcpy(leaf, leaf0, sizeof(PetscScalar) * (size_t)leafspan * bs);
    PetscCall(PetscSFReduceBegin(sf, unit, leaf, root, ops[o].op));
    PetscCall(PetscSFReduceEnd(sf, unit, leaf, root, ops[o].op));
    snprintf(name, sizeof name, "ref_reduce_%s.f64", ops[o].name);
    wr(dir, name, root, sizeof(PetscScalar) * (size_t)nroots * bs);
    if (ops[o].op == MPI_SUM || ops[o].op == MPI_MAX) {
      memcpy(ri, ri0, sizeof(PetscInt) * (size_t)nroots);
      memcpy(li, li0, sizeof(PetscInt) * (size_t)leafspan);
      PetscCall(PetscSFBcastBegin(sf, MPIU_INT, ri, li, ops[o].op));
      PetscCall(PetscSFBcastEnd(sf, MPIU_INT, ri, li, ops[o].op));
      snprintf(name, sizeof name, "ref_bcast_%s.i32", ops[o].name);
      wr(dir, name, li, sizeof(PetscInt) * (size_t)leafspan);
      memcpy(li, li0, sizeof(PetscInt) * (size_t)leafspan);
      PetscCall(PetscSFReduceBegin(sf, MPIU_INT, li, ri, ops[o].op));
      PetscCall(PetscSFReduceEnd(sf, MPIU_INT, li, ri, ops[o].op));
      snprintf(name, sizeof name, "ref_reduce_%s.i32", ops[o].name);
      wr(dir, name, ri, sizeof(PetscInt) * (size_t)nroots);
    }
  }
  if (bs > 1) PetscCallMPI(MPI_Type_free(&unit));
  PetscCall(PetscSFDestroy(&sf));
  PetscCall(PetscFree4(root, leaf, ri, li));
  PetscCall(PetscFree(remote));
  free(local); free(rem); free(root0); free(leaf0); free(ri0); free(li0);
  PetscFunctionReturn(PETSC_SUCCESS);
}

static void *rd(const char *dir, const char *name, size_t bytes)
{
  char  p[4096];
  void *buf = malloc(bytes ? bytes : 1);
  FILE *f;
  snprintf(p, sizeof p, "%s/%s", dir, name);
  f = fopen(p, "rb");
  if (!f || fread(buf, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %s\n", p); exit(2); }
  fclose(f);
  return buf;
}
static void wr(const char *dir, const char *name, const void *buf, size_t bytes)
{
  char  p[4096];
  FILE *f;
  snprintf(p, sizeof p, "%s/%s", dir, name);
  f = fopen(p, "wb");
  fwrite(buf, 1, bytes, f);
  fclose(f);
}

int main(int argc, char **argv)
{
  const char *dir;
  int         m, nv;
  long        nnz;
  PetscInt   *ai, *aj;
  PetscScalar *aa, *x, *y, *V, *alpha;
  Mat          A;
  Vec          vx, vy, vz, vd, *vv, vb, vu, vsol;
  PC           pc;
  KSP          ksp;
  PetscBool    dosolve = PETSC_FALSE;

  if (argc < 2) return 1;
  dir = argv[1];
  PetscCall(PetscInitialize(&argc, &argv, NULL, NULL));
  if (!strcmp(dir, "-bench7")) {
    PetscCall(bench7(atoi(argv[2])));
    PetscCall(PetscFinalize());
    return 0;
  }
  if (!strcmp(dir, "-coo")) {
    PetscCall(coo_case(argv[2]));
    PetscCall(PetscFinalize());
    return 0;
  }
  if (!strcmp(dir, "-sf")) {
    PetscCall(sf_case(argv[2]));
    PetscCall(PetscFinalize());
    return 0;
  }
  {
    char  p[4096];
    FILE *f;
    snprintf(p, sizeof p, "%s/meta.txt", dir);
    f = fopen(p, "r");
    if (!f || fscanf(f, "%d %ld %d", &m, &nnz, &nv) != 3) return 2;
    fclose(f);
  }
  ai    = (PetscInt *)rd(dir, "ai.i32", sizeof(PetscInt) * (size_t)(m + 1));
  aj    = (PetscInt *)rd(dir, "aj.i32", sizeof(PetscInt) * (size_t)nnz);
  aa    = (PetscScalar *)rd(dir, "aa.f64", sizeof(PetscScalar) * (size_t)nnz);
  x     = (PetscScalar *)rd(dir, "x.f64", sizeof(PetscScalar) * (size_t)m);
  y     = (PetscScalar *)rd(dir, "y.f64", sizeof(PetscScalar) * (size_t)m);
  V     = (PetscScalar *)rd(dir, "V.f64", sizeof(PetscScalar) * (size_t)m * nv);
  alpha = (PetscScalar *)rd(dir, "alpha.f64", sizeof(PetscScalar) * (size_t)nv);

  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, m, ai, aj, aa, &A));
  PetscCall(VecCreateSeqWithArray(PETSC_COMM_SELF, 1, m, x, &vx));
  PetscCall(VecCreateSeqWithArray(PETSC_COMM_SELF, 1, m, y, &vy));
  PetscCall(VecDuplicate(vx, &vz));
  PetscCall(VecDuplicate(vx, &vd));
  {
    const PetscScalar *z;
    /* MatMult / MatMultAdd / MatGetDiagonal */
    PetscCall(MatMult(A, vx, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_mult.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(MatMultAdd(A, vx, vy, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_multadd.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(MatGetDiagonal(A, vd));
    PetscCall(VecGetArrayRead(vd, &z)); wr(dir, "ref_diag.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vd, &z));
    /* MatMultTranspose / MatMultTransposeAdd */
    PetscCall(MatMultTranspose(A, vx, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_multtr.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(MatMultTransposeAdd(A, vx, vy, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_multtradd.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
  }
  /* VecMDot / VecMAXPY / VecDot / VecNorm */
  PetscCall(PetscMalloc1(nv, &vv));
  for (int j = 0; j < nv; j++) PetscCall(VecCreateSeqWithArray(PETSC_COMM_SELF, 1, m, V + (size_t)j * m, &vv[j]));
  {
    PetscScalar *dots, dn[3];
    PetscReal    nrm;
    const PetscScalar *z;
    PetscCall(PetscMalloc1(nv, &dots));
    PetscCall(VecMDot(vx, nv, vv, dots));
    wr(dir, "ref_mdot.f64", dots, sizeof(PetscScalar) * nv);
    PetscCall(VecCopy(vy, vz));
    PetscCall(VecMAXPY(vz, nv, alpha, vv));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_maxpy.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(VecDot(vx, vy, &dn[0]));
    PetscCall(VecNorm(vx, NORM_2, &nrm));
    dn[1] = nrm;
    PetscCall(VecNorm(vy, NORM_2, &nrm));
    dn[2] = nrm;
    wr(dir, "ref_dotnorm.f64", dn, sizeof dn);
    PetscCall(PetscFree(dots));
  }
  /* PCILU(0) apply = MatSolve ; PCJACOBI apply */
  {
    const PetscScalar *z;
    PetscCall(PCCreate(PETSC_COMM_SELF, &pc));
    PetscCall(PCSetType(pc, PCILU));
    PetscCall(PCSetOperators(pc, A, A));
    PetscCall(PCSetUp(pc));
    PetscCall(PCApply(pc, vx, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_ilusolve.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(PCDestroy(&pc));
    {
      PetscBool doicc = PETSC_FALSE; /* -icc: the matrix is symmetric, also record PCApply of PCICC (ICC(0), natural ordering) */
      PetscCall(PetscOptionsGetBool(NULL, NULL, "-icc", &doicc, NULL));
      if (doicc) {
        PetscCall(PCCreate(PETSC_COMM_SELF, &pc));
        PetscCall(PCSetType(pc, PCICC));
        PetscCall(PCSetOperators(pc, A, A));
        PetscCall(PCSetUp(pc));
        PetscCall(PCApply(pc, vx, vz));
        PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_iccsolve.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
        PetscCall(PCDestroy(&pc));
      }
    }
    PetscCall(PCCreate(PETSC_COMM_SELF, &pc));
    PetscCall(PCSetType(pc, PCJACOBI));
    PetscCall(PCSetOperators(pc, A, A));
    PetscCall(PCSetUp(pc));
    PetscCall(PCApply(pc, vx, vz));
    PetscCall(VecGetArrayRead(vz, &z)); wr(dir, "ref_jacobi.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vz, &z));
    PetscCall(PCDestroy(&pc));
  }
  /* KSPSolve with b = A*1 (ex2.c / bench_kspsolve.c convention), options from the command line */
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-solve", &dosolve, NULL));
  if (dosolve) {
    PetscReal         *hist;
    PetscInt           nh = 0, its, cap = 200000;
    const PetscReal   *h;
    const PetscScalar *z;
    KSPConvergedReason reason;
    PetscReal          rnorm;
    char               p[4096];
    FILE              *f;
    PetscCall(VecDuplicate(vx, &vb));
    PetscCall(VecDuplicate(vx, &vu));
    PetscCall(VecDuplicate(vx, &vsol));
    PetscCall(VecSet(vu, 1.0));
    PetscCall(MatMult(A, vu, vb));
    PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
    PetscCall(KSPSetOperators(ksp, A, A));
    PetscCall(PetscMalloc1(cap, &hist));
    PetscCall(KSPSetResidualHistory(ksp, hist, cap, PETSC_TRUE));
    PetscCall(KSPSetFromOptions(ksp));
    PetscCall(KSPSolve(ksp, vb, vsol));
    PetscCall(KSPGetResidualHistory(ksp, &h, &nh));
    PetscCall(KSPGetIterationNumber(ksp, &its));
    PetscCall(KSPGetConvergedReason(ksp, &reason));
    PetscCall(KSPGetResidualNorm(ksp, &rnorm));
    wr(dir, "ref_hist.f64", h, sizeof(PetscReal) * (size_t)nh);
    PetscCall(VecGetArrayRead(vsol, &z)); wr(dir, "ref_sol.f64", z, sizeof(PetscScalar) * m); PetscCall(VecRestoreArrayRead(vsol, &z));
    snprintf(p, sizeof p, "%s/ref_ksp.txt", dir);
    f = fopen(p, "w");
    fprintf(f, "%d %d %.17g %d\n", (int)its, (int)reason, (double)rnorm, (int)nh);
    fclose(f);
    PetscCall(KSPDestroy(&ksp));
    PetscCall(VecDestroy(&vb)); PetscCall(VecDestroy(&vu)); PetscCall(VecDestroy(&vsol));
    PetscCall(PetscFree(hist));
  }
  for (int j = 0; j < nv; j++) PetscCall(VecDestroy(&vv[j]));
  PetscCall(PetscFree(vv));
  PetscCall(VecDestroy(&vx)); PetscCall(VecDestroy(&vy)); PetscCall(VecDestroy(&vz)); PetscCall(VecDestroy(&vd));
  PetscCall(MatDestroy(&A));
  PetscCall(PetscFinalize());
  return 0;
}
