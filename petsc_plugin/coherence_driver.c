/*
 * coherence_driver.c -- a PETSc program (public API only) with one check per host/device COHERENCE rule of the plugin that the
 * reference's own programs exposed (profiles/round2_notes.md): each block does the same thing on the plugin's types (whatever
 * -vec_type / -mat_type say) and on the reference's host types and compares bit for bit.
 *   1. norm cache: PETSc caches norms by object state; a device write must bump it (VecSet(y,0); MatMult -> y; VecNorm(y);
 *      VecSet(v,0) after a fused PCApplyBAorAB wrote v)
 *   2. a Vec operation on a vector whose host array is handed out (DMDAVecGetArray(F); VecZeroEntries(F); fill; restore)
 *   3. VecGetSubVector / VecRestoreSubVector written on the device (MatMult into a sub-vector), then the parent read on the device
 *   4. MatHeaderMerge: in-place MatLUFactor of an aijb200 matrix, MatSolve, MatDestroy
 *   5. PCJacobiGetDiagonal after KSPSolve with the PCJACOBI sub-class
 *   6. values changed through MatZeroEntries + MatSetValues(ADD) with an unchanged pattern, second KSPSolve (BiCGStab + Jacobi)
 * Prints "ok <name>" per check, "all ok" at the end, non-zero exit on the first mismatch.  Test infrastructure
 * (tests/test_plugin_logic_mock_cpu.py on the mock device, tests/test_petsc_plugin_gpu.py on a GPU); built by oracle/build_ref_demo.sh.
 */
#include <petscksp.h>

#define CHECK(cond, name) \
  do { \
    if (!(cond)) { \
      PetscCall(PetscPrintf(PETSC_COMM_SELF, "FAILED %s\n", name)); \
      PetscCall(PetscFinalize()); \
      return 1; \
    } \
    PetscCall(PetscPrintf(PETSC_COMM_SELF, "ok %s\n", name)); \
  } while (0)

static PetscErrorCode Lap1D(Mat A, PetscInt n, PetscScalar diag, PetscScalar lower)
{
  PetscFunctionBegin;
  for (PetscInt i = 0; i < n; i++) {
    PetscCall(MatSetValue(A, i, i, diag + 0.01 * i, ADD_VALUES));
    if (i) PetscCall(MatSetValue(A, i, i - 1, lower, ADD_VALUES));
    if (i < n - 1) PetscCall(MatSetValue(A, i, i + 1, -1.0, ADD_VALUES));
  }
  PetscCall(MatAssemblyBegin(A, MAT_FINAL_ASSEMBLY));
  PetscCall(MatAssemblyEnd(A, MAT_FINAL_ASSEMBLY));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode Same(Vec v, Vec r, PetscBool *same)
{
  const PetscScalar *a, *b;
  PetscInt           n;
  PetscFunctionBegin;
  PetscCall(VecGetLocalSize(r, &n));
  PetscCall(VecGetArrayRead(v, &a));
  PetscCall(VecGetArrayRead(r, &b));
  *same = (PetscBool)(memcmp(a, b, sizeof(PetscScalar) * (size_t)n) == 0);
  PetscCall(VecRestoreArrayRead(v, &a));
  PetscCall(VecRestoreArrayRead(r, &b));
  PetscFunctionReturn(PETSC_SUCCESS);
}

int main(int argc, char **argv)
{
  const PetscInt n = 200;
  Mat            A, R;
  Vec            x, y, xr, yr;
  PetscReal      na, nr;
  PetscBool      same;

  PetscCall(PetscInitialize(&argc, &argv, NULL, NULL));
  PetscCall(MatCreate(PETSC_COMM_SELF, &A));
  PetscCall(MatSetSizes(A, n, n, n, n));
  PetscCall(MatSetFromOptions(A));
  PetscCall(MatSetUp(A));
  PetscCall(MatCreateSeqAIJ(PETSC_COMM_SELF, n, n, 3, NULL, &R));
  PetscCall(Lap1D(A, n, 4.0, -1.0));
  PetscCall(Lap1D(R, n, 4.0, -1.0));
  PetscCall(MatCreateVecs(A, &x, &y));
  PetscCall(MatCreateVecs(R, &xr, &yr));
  {
    MatType mt;
    VecType vt;
    PetscCall(MatGetType(A, &mt));
    PetscCall(VecGetType(x, &vt));
    PetscCall(PetscPrintf(PETSC_COMM_SELF, "mat type %s vec type %s\n", mt, vt));
  }

  /* 1. norm cache */
  for (PetscInt i = 0; i < n; i++) PetscCall(VecSetValue(xr, i, 1.0 / (1.0 + i), INSERT_VALUES));
  PetscCall(VecAssemblyBegin(xr));
  PetscCall(VecAssemblyEnd(xr));
  PetscCall(VecCopy(xr, x));
  PetscCall(VecSet(y, 0.0));
  PetscCall(VecSet(yr, 0.0));
  PetscCall(VecNorm(y, NORM_2, &na)); /* caches 0 */
  PetscCall(VecNorm(yr, NORM_2, &nr));
  PetscCall(MatMult(A, x, y));
  PetscCall(MatMult(R, xr, yr));
  PetscCall(VecNorm(y, NORM_2, &na));
  PetscCall(VecNorm(yr, NORM_2, &nr));
  CHECK(na > 0.0 && PetscAbsReal(na - nr) <= 1e-14 * nr, "norm_after_matmult_is_not_the_cached_one");
  PetscCall(VecSet(y, 0.0)); /* must not be skipped because of a stale cached 0 */
  PetscCall(VecNorm(y, NORM_INFINITY, &na));
  CHECK(na == 0.0, "vecset_zero_after_device_write");

  /* 2. a Vec operation while the host array is handed out */
  {
    PetscScalar *a, *b;
    PetscCall(VecSet(y, 7.0));
    PetscCall(VecSet(yr, 7.0));
    PetscCall(VecGetArray(y, &a));
    PetscCall(VecGetArray(yr, &b));
    PetscCall(VecZeroEntries(y));
    PetscCall(VecZeroEntries(yr));
    for (PetscInt i = 0; i < n; i += 2) a[i] = b[i] = 1.0 + i; /* every other entry: the rest must be the zeros of VecZeroEntries */
    PetscCall(VecRestoreArray(y, &a));
    PetscCall(VecRestoreArray(yr, &b));
    PetscCall(VecNorm(y, NORM_1, &na));
    PetscCall(VecNorm(yr, NORM_1, &nr));
    PetscCall(Same(y, yr, &same));
    CHECK(same && na == nr, "vecset_while_host_array_is_handed_out");
  }

  /* 3. sub-vector written on the device */
  {
    IS  is;
    Vec sub, subr, z, zr;
    PetscCall(ISCreateStride(PETSC_COMM_SELF, n, 0, 1, &is));
    PetscCall(VecDuplicate(y, &z));
    PetscCall(VecDuplicate(yr, &zr));
    for (int k = 0; k < 3; k++) {
      PetscCall(VecScale(x, 1.5));
      PetscCall(VecScale(xr, 1.5));
      PetscCall(VecGetSubVector(y, is, &sub));
      PetscCall(VecGetSubVector(yr, is, &subr));
      PetscCall(MatMult(A, x, sub));
      PetscCall(MatMult(R, xr, subr));
      PetscCall(VecRestoreSubVector(y, is, &sub));
      PetscCall(VecRestoreSubVector(yr, is, &subr));
      PetscCall(VecWAXPY(z, 2.0, y, x)); /* reads y on the device */
      PetscCall(VecWAXPY(zr, 2.0, yr, xr));
      PetscCall(Same(z, zr, &same));
      CHECK(same, "subvector_written_on_device_then_parent_read");
    }
    PetscCall(VecDestroy(&z));
    PetscCall(VecDestroy(&zr));
    PetscCall(ISDestroy(&is));
  }

  /* 5. PCJacobiGetDiagonal after a solve with the PCJACOBI sub-class; 6. new values, same pattern, second solve */
  {
    KSP       ksp, kspr;
    PC        pc, pcr;
    Vec       d, dr, b, br;
    PetscInt  its, itsr;
    PetscReal rn, rnr;
    PetscCall(VecDuplicate(x, &d));
    PetscCall(VecDuplicate(xr, &dr));
    PetscCall(VecDuplicate(x, &b));
    PetscCall(VecDuplicate(xr, &br));
    PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
    PetscCall(KSPCreate(PETSC_COMM_SELF, &kspr));
    PetscCall(KSPSetType(ksp, KSPBCGS));
    PetscCall(KSPSetType(kspr, KSPBCGS));
    PetscCall(KSPGetPC(ksp, &pc));
    PetscCall(KSPGetPC(kspr, &pcr));
    PetscCall(PCSetType(pc, PCJACOBI));
    PetscCall(PCSetType(pcr, PCJACOBI));
    PetscCall(KSPSetTolerances(ksp, 1e-10, PETSC_CURRENT, PETSC_CURRENT, 50));
    PetscCall(KSPSetTolerances(kspr, 1e-10, PETSC_CURRENT, PETSC_CURRENT, 50));
    for (int t = 0; t < 2; t++) {
      if (t) { /* ksp/tutorials/ex9.c: MatZeroEntries + ADD_VALUES with the same pattern, then solve again */
        PetscCall(MatZeroEntries(A));
        PetscCall(MatZeroEntries(R));
        PetscCall(Lap1D(A, n, 4.5, -1.5));
        PetscCall(Lap1D(R, n, 4.5, -1.5));
      }
      PetscCall(MatMult(A, x, b));
      PetscCall(MatMult(R, xr, br));
      PetscCall(KSPSetOperators(ksp, A, A));
      PetscCall(KSPSetOperators(kspr, R, R));
      PetscCall(KSPSetUp(ksp));
      PetscCall(KSPSetUp(kspr));
      PetscCall(KSPSolve(ksp, b, y));
      PetscCall(KSPSolve(kspr, br, yr));
      PetscCall(KSPGetIterationNumber(ksp, &its));
      PetscCall(KSPGetIterationNumber(kspr, &itsr));
      PetscCall(KSPGetResidualNorm(ksp, &rn));
      PetscCall(KSPGetResidualNorm(kspr, &rnr));
      PetscCall(PetscPrintf(PETSC_COMM_SELF, "solve %d: %" PetscInt_FMT " / %" PetscInt_FMT " iterations, residual %g / %g\n", t, its, itsr, (double)rn, (double)rnr));
      CHECK(its == itsr && PetscAbsReal(rn - rnr) <= 1e-6 * rnr + 1e-16, t ? "second_solve_after_value_update_bcgs_jacobi" : "first_solve_bcgs_jacobi");
      PetscCall(PCJacobiGetDiagonal(pc, d, NULL));
      PetscCall(PCJacobiGetDiagonal(pcr, dr, NULL));
      PetscCall(Same(d, dr, &same));
      CHECK(same, "pcjacobigetdiagonal_after_solve");
    }
    PetscCall(KSPDestroy(&ksp));
    PetscCall(KSPDestroy(&kspr));
    PetscCall(VecDestroy(&d));
    PetscCall(VecDestroy(&dr));
    PetscCall(VecDestroy(&b));
    PetscCall(VecDestroy(&br));
  }

  /* 4. in-place LU: MatHeaderMerge leaves the plugin's ops on a plain factor's data */
  {
    IS            perm, iperm;
    MatFactorInfo info;
    PetscCall(MatGetOrdering(A, MATORDERINGNATURAL, &perm, &iperm));
    PetscCall(MatFactorInfoInitialize(&info));
    PetscCall(MatMult(A, x, y)); /* rhs before the matrix turns into its factor */
    PetscCall(MatMult(R, xr, yr));
    PetscCall(MatLUFactor(A, perm, iperm, &info));
    PetscCall(MatLUFactor(R, perm, iperm, &info));
    {
      Vec s, sr;
      PetscCall(VecDuplicate(x, &s));
      PetscCall(VecDuplicate(xr, &sr));
      PetscCall(MatSolve(A, y, s));
      PetscCall(MatSolve(R, yr, sr));
      PetscCall(Same(s, sr, &same));
      CHECK(same, "inplace_lu_solve");
      PetscCall(VecDestroy(&s));
      PetscCall(VecDestroy(&sr));
    }
    PetscCall(ISDestroy(&perm));
    PetscCall(ISDestroy(&iperm));
  }
  PetscCall(MatDestroy(&A)); /* used to dereference a NULL context */
  PetscCall(MatDestroy(&R));
  PetscCall(PetscPrintf(PETSC_COMM_SELF, "ok destroy_after_headermerge\n"));
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&y));
  PetscCall(VecDestroy(&xr));
  PetscCall(VecDestroy(&yr));
  PetscCall(PetscPrintf(PETSC_COMM_SELF, "all ok\n"));
  PetscCall(PetscFinalize());
  return 0;
}
