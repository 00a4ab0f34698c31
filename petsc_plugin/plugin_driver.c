/*
 * plugin_driver.c -- a PETSc program (public API only + the C ABI for device buffers) that exercises the paths of the plugin the
 * reference's tutorials do not reach: COO assembly from DEVICE-resident arrays, MatMultTranspose[Add], MatBindToCPU,
 * MatGetCurrentMemType.  Every result is compared with the same operation on the reference's CPU types (MATSEQAIJ / VECSEQ)
 * in the same process; prints one "ok <name>" line per check and exits non-zero on the first mismatch.
 * Test infrastructure (run by tests/test_petsc_plugin_gpu.py with -dll_append <plugin>); built by oracle/build_ref_demo.sh.
 */
#include <petscmat.h>
#include "petscb200.h"

#define CHECK(cond, name) \
  do { \
    if (!(cond)) { \
      PetscCall(PetscPrintf(PETSC_COMM_SELF, "FAILED %s\n", name)); \
      PetscCall(PetscFinalize()); \
      return 1; \
    } \
    PetscCall(PetscPrintf(PETSC_COMM_SELF, "ok %s\n", name)); \
  } while (0)

int main(int argc, char **argv)
{
  const PetscInt M = 57, N = 57, n = 3000;
  PetscInt      *ci, *cj;
  PetscScalar   *v1, *v2;
  Mat            A, R;
  Vec            x, y, z, xr, yr, zr;
  PetscBool      eq, isb200;
  b200Handle     h;
  void          *d_i, *d_j, *d_v1, *d_v2;
  PetscMemType   mt;
  PetscReal      nrm;
  unsigned       s = 12345u;

  PetscCall(PetscInitialize(&argc, &argv, NULL, NULL));
  PetscCall(PetscMalloc4(n, &ci, n, &cj, n, &v1, n, &v2));
  for (PetscInt k = 0; k < n; k++) { /* repeats (about 1 in 3 entries), a few ignored negative indices */
    s     = s * 1664525u + 1013904223u;
    ci[k] = (PetscInt)((s >> 8) % (M + 1)) - 1;
    s     = s * 1664525u + 1013904223u;
    cj[k] = (PetscInt)((s >> 8) % (N + 1)) - 1;
    s     = s * 1664525u + 1013904223u;
    v1[k] = ((double)(s >> 8) / 8388608.0) - 1.0;
    s     = s * 1664525u + 1013904223u;
    v2[k] = ((double)(s >> 8) / 8388608.0) - 1.0;
  }
  /* reference: host COO on MATSEQAIJ (the index arrays are modified in place by the reference: give it copies) */
  {
    PetscInt *ri, *rj;
    PetscCall(PetscMalloc2(n, &ri, n, &rj));
    PetscCall(PetscArraycpy(ri, ci, n));
    PetscCall(PetscArraycpy(rj, cj, n));
    PetscCall(MatCreate(PETSC_COMM_SELF, &R));
    PetscCall(MatSetSizes(R, M, N, M, N));
    PetscCall(MatSetType(R, MATSEQAIJ));
    PetscCall(MatSetPreallocationCOO(R, n, ri, rj));
    PetscCall(MatSetValuesCOO(R, v1, INSERT_VALUES));
    PetscCall(PetscFree2(ri, rj));
  }
  /* plugin: the same arrays resident on the device */
  PetscCall(MatCreate(PETSC_COMM_SELF, &A));
  PetscCall(MatSetSizes(A, M, N, M, N));
  PetscCall(MatSetFromOptions(A)); /* -mat_type aijb200 */
  PetscCall(PetscObjectTypeCompare((PetscObject)A, "seqaijb200", &isb200));
  CHECK(isb200, "mat_type_is_seqaijb200");
  PetscCheck(!b200Create(&h, -1), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200Create");
  PetscCheck(!b200Malloc(h, &d_i, sizeof(PetscInt) * n) && !b200Malloc(h, &d_j, sizeof(PetscInt) * n) && !b200Malloc(h, &d_v1, sizeof(PetscScalar) * n) && !b200Malloc(h, &d_v2, sizeof(PetscScalar) * n), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200Malloc");
  PetscCheck(!b200MemcpyHtoD(h, d_i, ci, sizeof(PetscInt) * n) && !b200MemcpyHtoD(h, d_j, cj, sizeof(PetscInt) * n) && !b200MemcpyHtoD(h, d_v1, v1, sizeof(PetscScalar) * n) && !b200MemcpyHtoD(h, d_v2, v2, sizeof(PetscScalar) * n), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200MemcpyHtoD");
  PetscCall(MatSetPreallocationCOO(A, n, (PetscInt *)d_i, (PetscInt *)d_j));
  PetscCall(MatSetValuesCOO(A, (const PetscScalar *)d_v1, INSERT_VALUES));
  PetscCall(MatEqual(A, R, &eq)); /* pattern and values, bit for bit (the plugin adopts the reference's summation order) */
  CHECK(eq, "coo_device_insert_equals_reference");
  PetscCall(MatSetValuesCOO(A, (const PetscScalar *)d_v2, ADD_VALUES));
  PetscCall(MatSetValuesCOO(R, v2, ADD_VALUES));
  PetscCall(MatEqual(A, R, &eq));
  CHECK(eq, "coo_device_add_equals_reference");
  PetscCall(MatSetValuesCOO(A, v1, INSERT_VALUES)); /* host values on the device type */
  PetscCall(MatSetValuesCOO(R, v1, INSERT_VALUES));
  PetscCall(MatEqual(A, R, &eq));
  CHECK(eq, "coo_host_values_equals_reference");

  /* products: device types vs reference types, bit for bit (MatMult, MatMultTranspose, MatMultTransposeAdd) */
  PetscCall(MatCreateVecs(A, &x, &y));
  PetscCall(MatCreateVecs(R, &xr, &yr));
  PetscCall(PetscObjectTypeCompare((PetscObject)x, "seqb200", &isb200));
  CHECK(isb200, "matcreatevecs_gives_seqb200");
  PetscCall(VecDuplicate(x, &z));
  PetscCall(VecDuplicate(xr, &zr));
  {
    PetscScalar *a;
    PetscCall(VecGetArrayWrite(xr, &a));
    for (PetscInt k = 0; k < N; k++) a[k] = v2[k];
    PetscCall(VecRestoreArrayWrite(xr, &a));
    PetscCall(VecCopy(xr, x));
    PetscCall(VecGetArrayWrite(zr, &a));
    for (PetscInt k = 0; k < N; k++) a[k] = v1[k + 100];
    PetscCall(VecRestoreArrayWrite(zr, &a));
    PetscCall(VecCopy(zr, z));
  }
  PetscCall(MatMult(A, x, y));
  PetscCall(MatMult(R, xr, yr));
  PetscCall(VecAXPY(yr, -1.0, y));
  PetscCall(VecNorm(yr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmult_bit_exact");
  PetscCall(MatMultTranspose(A, x, y));
  PetscCall(MatMultTranspose(R, xr, yr));
  PetscCall(VecAXPY(yr, -1.0, y));
  PetscCall(VecNorm(yr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmulttranspose_bit_exact");
  PetscCall(MatMultTransposeAdd(A, x, z, y));
  PetscCall(MatMultTransposeAdd(R, xr, zr, yr));
  PetscCall(VecAXPY(yr, -1.0, y));
  PetscCall(VecNorm(yr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmulttransposeadd_bit_exact");
  PetscCall(MatMultTransposeAdd(A, x, z, z)); /* in place */
  PetscCall(MatMultTransposeAdd(R, xr, zr, zr));
  PetscCall(VecAXPY(zr, -1.0, z));
  PetscCall(VecNorm(zr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmulttransposeadd_inplace_bit_exact");
  /* values change through the host API -> mirror and transposed copy follow */
  PetscCall(MatScale(A, 0.5));
  PetscCall(MatScale(R, 0.5));
  PetscCall(MatMultTranspose(A, x, y));
  PetscCall(MatMultTranspose(R, xr, yr));
  PetscCall(VecAXPY(yr, -1.0, y));
  PetscCall(VecNorm(yr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmulttranspose_after_matscale");
  PetscCall(MatGetCurrentMemType(A, &mt));
  CHECK(PetscMemTypeDevice(mt), "current_memtype_is_device");
  PetscCall(MatBindToCPU(A, PETSC_TRUE));
  PetscCall(MatMult(A, x, y));
  PetscCall(MatMult(R, xr, yr));
  PetscCall(VecAXPY(yr, -1.0, y));
  PetscCall(VecNorm(yr, NORM_INFINITY, &nrm));
  CHECK(nrm == 0.0, "matmult_bound_to_cpu");
  PetscCall(MatBindToCPU(A, PETSC_FALSE));

  b200Free(h, d_i); b200Free(h, d_j); b200Free(h, d_v1); b200Free(h, d_v2);
  b200Destroy(h);
  PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&y)); PetscCall(VecDestroy(&z));
  PetscCall(VecDestroy(&xr)); PetscCall(VecDestroy(&yr)); PetscCall(VecDestroy(&zr));
  PetscCall(MatDestroy(&A)); PetscCall(MatDestroy(&R));
  PetscCall(PetscFree4(ci, cj, v1, v2));
  PetscCall(PetscPrintf(PETSC_COMM_SELF, "all ok\n"));
  PetscCall(PetscFinalize());
  return 0;
}
