/*
 * b200_driver.c -- a PETSc PROGRAM (real libpetsc: KSPSolve is the reference's own gmres.c / cg.c / bjacobi.c / ilu.c) that runs
 * the BASELINE.json workloads on the b200 types and reports device-timed numbers.  bench.py, tools/ and the GPU tests call it;
 * it is also a stand-alone executable (ncu captures).  Everything that touches PETSc goes through its public API; the types
 * come from libpetscb200plugin.so (linked, registered with PetscDLLibraryRegister_petscb200plugin) and timing uses CUDA
 * events on the library's stream through the C ABI (include/petscb200.h).
 *
 *   b200_driver -bench gmres7 [-n 512 | -nx -ny -nzl] [-steps K -warmup W] [-e2e 1] [-kernels 1] [PETSc options]
 *       BASELINE configs[1] (and [3] with -pc_type bjacobi -sub_pc_type ilu): 7-point Laplacian, nx*ny*nzl rows per rank,
 *       row-partitioned over the NCCL ranks (mpiaijb200/mpib200) when PETSCB200_NRANKS > 1; one step = one GMRES(30) cycle.
 *   b200_driver -bench cg27 [-n 256]          configs[2]: 27-point operator of bench_kspsolve.c, KSPCG + PCILU(0)
 *   b200_driver -bench rand [-rand_n N -rand_d d]   configs[4]: random CSR, MatMult only
 *   b200_driver -bench ex2 [-m 100]           configs[0]: ex2's 5-point operator assembled with MatSetValues like ex2.c
 *   b200_driver -parity <dir>                 multi-rank parity: reads this rank's row block (written by the Python side from
 *                                             oracle inputs), runs the mpiaijb200 ops, writes results for comparison
 *
 * Output: one line "B200JSON {...}" per measurement on stdout and, with -json_out <file>, the same objects appended there.
 */
#include <petscksp.h>
#include <petsctime.h>
#include <stdarg.h>
#include "petscb200.h"

PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_petscb200plugin(void);
PETSC_EXTERN PetscErrorCode PetscB200GetHandle(void **, int *, int *);
PETSC_EXTERN PetscErrorCode MatCreateMPIAIJB200WithArrays(PetscInt, PetscInt, PetscInt, const PetscInt[], const PetscInt[], const PetscScalar[], Mat *);
PETSC_EXTERN PetscErrorCode MatCreateMPIAIJB200WithSplitArrays(PetscInt, PetscInt, PetscInt, PetscInt[], PetscInt[], PetscScalar[], PetscInt[], PetscInt[], PetscScalar[], Mat *);
PETSC_EXTERN PetscErrorCode MatCreateSeqAIJB200WithDeviceArrays(PetscInt, PetscInt, PetscInt *, PetscInt *, PetscScalar *, Mat *);
PETSC_EXTERN PetscErrorCode MatMPIAIJB200GetSeqAIJ(Mat, Mat *, Mat *, const PetscInt **);

#define CB(...) \
  do { \
    int e_ = (__VA_ARGS__); \
    PetscCheck(!e_, PETSC_COMM_SELF, (PetscErrorCode)e_, "%s", b200GetLastErrorString()); \
  } while (0)

static b200Handle H;
static int        RANK, SIZE;
static FILE      *JOUT;

static void emit(const char *fmt, ...)
{
  char    buf[16384];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (RANK == 0) {
    printf("B200JSON %s\n", buf);
    fflush(stdout);
    if (JOUT) {
      fprintf(JOUT, "%s\n", buf);
      fflush(JOUT);
    }
  }
}

/* device timing: events on the library stream, barrier + synchronise on both sides, max over ranks */
typedef struct {
  b200Event a, b;
} Timer;
static PetscErrorCode TimerStart(Timer *t)
{
  PetscFunctionBeginUser;
  if (!t->a) {
    CB(b200EventCreate(&t->a));
    CB(b200EventCreate(&t->b));
  }
  CB(b200CommBarrier(H)); /* all-reduce + stream synchronise (b200Synchronize alone on one rank) */
  CB(b200EventRecord(H, t->a));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode TimerStop(Timer *t, double *ms_max)
{
  double ms, *d = NULL;
  PetscFunctionBeginUser;
  CB(b200EventRecord(H, t->b));
  CB(b200EventElapsedMs(t->a, t->b, &ms));
  CB(b200Synchronize(H));
  if (SIZE > 1) {
    CB(b200Malloc(H, (void **)&d, sizeof(double)));
    CB(b200MemcpyHtoD(H, d, &ms, sizeof(double)));
    CB(b200CommAllreduceMax(H, d, 1));
    CB(b200MemcpyDtoH(H, &ms, d, sizeof(double)));
    CB(b200Free(H, d));
  }
  *ms_max = ms;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode SumOverRanks(double *v)
{
  double *d = NULL;
  PetscFunctionBeginUser;
  if (SIZE > 1) {
    CB(b200Malloc(H, (void **)&d, sizeof(double)));
    CB(b200MemcpyHtoD(H, d, v, sizeof(double)));
    CB(b200CommAllreduceSum(H, d, 1));
    CB(b200MemcpyDtoH(H, v, d, sizeof(double)));
    CB(b200Free(H, d));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* this rank's rows of the 7-point operator, generated on the device, as a b200 matrix */
static PetscErrorCode Laplace7Mat(PetscInt nx, PetscInt ny, PetscInt nzl, Mat *A, PetscInt *nloc_out, int64_t *nnz_out)
{
  const int64_t nloc = (int64_t)nx * ny * nzl, r0 = nloc * RANK, r1 = r0 + nloc;
  const PetscInt nzg = nzl * SIZE;
  int64_t  nnz;
  int     *d_i, *d_j;
  double  *d_a;
  PetscFunctionBeginUser;
  PetscCheck(nloc * SIZE < 2147483647LL, PETSC_COMM_SELF, PETSC_ERR_SUP, "global size exceeds the 32-bit PetscInt of this PETSc");
  CB(b200GenLaplace7Nnz((int)nx, (int)ny, (int)nzg, r0, r1, &nnz));
  CB(b200Malloc(H, (void **)&d_i, sizeof(int) * ((size_t)nloc + 1)));
  CB(b200Malloc(H, (void **)&d_j, sizeof(int) * (size_t)nnz));
  CB(b200Malloc(H, (void **)&d_a, sizeof(double) * (size_t)nnz));
  CB(b200GenLaplace7(H, (int)nx, (int)ny, (int)nzg, r0, r1, d_i, d_j, d_a));
  if (SIZE == 1) PetscCall(MatCreateSeqAIJB200WithDeviceArrays((PetscInt)nloc, (PetscInt)nloc, d_i, d_j, d_a, A)); /* adopts the arrays */
  else {
    PetscCall(MatCreateMPIAIJB200WithArrays((PetscInt)nloc, (PetscInt)(nloc * SIZE), (PetscInt)r0, d_i, d_j, d_a, A)); /* copies */
    CB(b200Free(H, d_i));
    CB(b200Free(H, d_j));
    CB(b200Free(H, d_a));
  }
  *nloc_out = (PetscInt)nloc;
  *nnz_out  = nnz;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* time one PETSc call sequence with events: reps after warm-ups */
#define TIME_LOOP(ms_out, warm, reps, stmt) \
  do { \
    Timer  t_ = {0}; \
    double m_; \
    for (int w_ = 0; w_ < (warm); w_++) { stmt; } \
    PetscCall(TimerStart(&t_)); \
    for (int r_ = 0; r_ < (reps); r_++) { stmt; } \
    PetscCall(TimerStop(&t_, &m_)); \
    (ms_out) = m_ / (reps); \
    CB(b200EventDestroy(t_.a)); \
    CB(b200EventDestroy(t_.b)); \
  } while (0)

static PetscErrorCode BenchGmres7(void)
{
  PetscInt  n = 512, nx, ny, nzl, steps = 5, warmup = 3, restart = 30, nloc = 0, its;
  PetscBool e2e = PETSC_FALSE, kernels = PETSC_TRUE, flg;
  Mat       A;
  Vec       x, b, u;
  KSP       ksp;
  int64_t   nnz = 0;
  double    ms, rnorm_d, bsum;
  Timer     T = {0};
  long long l0, l1, h2d0, d2h0, h2d1, d2h1;
  PetscReal rnorm;
  char      pctype[64] = "jacobi", ksptype[64] = "gmres";
  PC        pc;
  KSPConvergedReason reason;

  PetscFunctionBeginUser;
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-n", &n, NULL));
  nx = ny = nzl = n;
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-nx", &nx, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-ny", &ny, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-nzl", &nzl, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-steps", &steps, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-warmup", &warmup, NULL));
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-e2e", &e2e, NULL));
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-kernels", &kernels, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-ksp_gmres_restart", &restart, NULL));
  PetscCall(PetscOptionsGetString(NULL, NULL, "-pc_type", pctype, sizeof pctype, &flg));
  PetscCall(PetscOptionsGetString(NULL, NULL, "-ksp_type", ksptype, sizeof ksptype, &flg));
  /* never converge, never diverge: exactly steps*restart iterations (the reference's own tolerances otherwise) */
  PetscCall(PetscOptionsInsertString(NULL, "-ksp_rtol 1e-300 -ksp_atol 1e-300 -ksp_divtol 1e300 -mat_no_inode"));
  if (!flg) PetscCall(PetscOptionsInsertString(NULL, "-ksp_type gmres"));

  PetscCall(Laplace7Mat(nx, ny, nzl, &A, &nloc, &nnz));
  PetscCall(MatCreateVecs(A, &x, &b));
  PetscCall(VecDuplicate(x, &u));
  PetscCall(VecSet(u, 1.0));
  PetscCall(MatMult(A, u, b)); /* b = A*1 (ex2.c:140, bench_kspsolve.c:373) */
  PetscCall(VecSum(b, &bsum)); /* exact integer checksum of the operator: sum(A*1) = 7N - nnz summed over ranks */
  PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
  PetscCall(KSPSetOperators(ksp, A, A));
  PetscCall(KSPGetPC(ksp, &pc));
  PetscCall(PCSetType(pc, PCJACOBI));
  PetscCall(KSPSetFromOptions(ksp));
  PetscCall(KSPSetTolerances(ksp, PETSC_CURRENT, PETSC_CURRENT, PETSC_CURRENT, restart * PetscMax(warmup, 1)));
  PetscCall(KSPSolve(ksp, b, x)); /* warm-up cycles: also allocates the Krylov basis (gmres.c:399-418) and sets up the PC */
  PetscCall(KSPSetTolerances(ksp, PETSC_CURRENT, PETSC_CURRENT, PETSC_CURRENT, restart * steps));
  l0 = b200KernelLaunchCount();
  CB(b200TransferCounters(&h2d0, &d2h0));
  PetscCall(TimerStart(&T));
  PetscCall(KSPSolve(ksp, b, x));
  PetscCall(TimerStop(&T, &ms));
  l1 = b200KernelLaunchCount();
  CB(b200TransferCounters(&h2d1, &d2h1));
  PetscCall(KSPGetIterationNumber(ksp, &its));
  PetscCall(KSPGetConvergedReason(ksp, &reason));
  PetscCall(KSPGetResidualNorm(ksp, &rnorm));
  rnorm_d = (double)rnorm;
  PetscCheck(its == restart * steps, PETSC_COMM_SELF, PETSC_ERR_PLIB, "expected %d iterations, KSPSolve did %d (reason %d)", (int)(restart * steps), (int)its, (int)reason);
  {
    double nnz_g = (double)nnz, bs = bsum;
    PetscCall(SumOverRanks(&nnz_g));
    emit("{\"kind\":\"solve\",\"bench\":\"gmres7\",\"ksp_type\":\"%s\",\"pc_type\":\"%s\",\"nx\":%d,\"ny\":%d,\"nz_local\":%d,\"n_ranks\":%d,\"rows_per_rank\":%d,\"nnz_per_rank\":%lld,"
         "\"steps\":%d,\"warmup\":%d,\"restart\":%d,\"iterations\":%d,\"ms_total\":%.4f,\"ms_per_step\":%.4f,\"iterations_per_sec\":%.4f,\"rnorm\":%.15e,\"reason\":%d,"
         "\"gpu_launches\":%lld,\"h2d_bytes_in_timed_region\":%lld,\"d2h_bytes_in_timed_region\":%lld,\"sum_A_ones\":%.1f,\"expected_sum_A_ones\":%.1f}",
         ksptype, pctype, (int)nx, (int)ny, (int)nzl, SIZE, (int)nloc, (long long)nnz, (int)steps, (int)warmup, (int)restart, (int)its, ms, ms / steps, its / (ms * 1e-3), rnorm_d, (int)reason,
         l1 - l0, h2d1 - h2d0, d2h1 - d2h0, bs, 7.0 * (double)nloc * SIZE - nnz_g);
  }

  if (kernels) {
    /* the three kernels that make up the cycle, each alone, through the same PETSc calls KSPGMRESCycle makes */
    Mat          Ad = A;
    Vec          xs, ys, *V;
    PetscScalar *al, dots[32];
    double       spmv_ms, maxpy_ms, mdot_ms, pcba_ms = 0;
    const int    nv  = 30;
    int64_t      nzd = nnz;
    if (SIZE > 1) {
      MatInfo info;
      PetscCall(MatMPIAIJB200GetSeqAIJ(A, &Ad, NULL, NULL));
      PetscCall(MatGetInfo(Ad, MAT_LOCAL, &info));
      nzd = (int64_t)info.nz_used;
    }
    PetscCall(MatCreateVecs(Ad, &xs, &ys));
    PetscCall(VecSet(xs, 1.0));
    TIME_LOOP(spmv_ms, 3, 20, PetscCall(MatMult(Ad, xs, ys)));
    TIME_LOOP(pcba_ms, 3, 20, PetscCall(PCApplyBAorAB(pc, PC_LEFT, x, u, b)));
    PetscCall(VecDuplicateVecs(xs, nv, &V));
    PetscCall(PetscMalloc1(nv, &al));
    for (int j = 0; j < nv; j++) {
      PetscCall(VecSet(V[j], 1.0 / (j + 1)));
      al[j] = 1e-3 * (j + 1);
    }
    TIME_LOOP(maxpy_ms, 2, 10, PetscCall(VecMAXPY(ys, nv, al, V)));
    TIME_LOOP(mdot_ms, 2, 10, PetscCall(VecMDot(ys, nv, V, dots)));
    emit("{\"kind\":\"kernels\",\"bench\":\"gmres7\",\"rows\":%d,\"nnz_diag_block\":%lld,\"spmv_ms\":%.5f,\"spmv_algorithmic_bytes\":%lld,\"spmv_flops\":%lld,"
         "\"pcapplyba_ms\":%.5f,\"maxpy_nv\":%d,\"maxpy_ms\":%.5f,\"maxpy_algorithmic_bytes\":%lld,\"mdot_nv\":%d,\"mdot_ms\":%.5f,\"mdot_algorithmic_bytes\":%lld}",
         (int)nloc, (long long)nzd, spmv_ms, (long long)(nzd * 12 + (int64_t)nloc * 20), (long long)(2 * nzd - nloc), pcba_ms, nv, maxpy_ms, (long long)(8LL * nloc * (nv + 2)), nv, mdot_ms,
         (long long)(8LL * nloc * (nv + 1)));
    PetscCall(VecDestroyVecs(nv, &V));
    PetscCall(PetscFree(al));
    PetscCall(VecDestroy(&xs));
    PetscCall(VecDestroy(&ys));
  }
  PetscCall(KSPDestroy(&ksp));
  PetscCall(VecDestroy(&u));

  if (e2e) {
    /* end to end through PETSc's public API with HOST buffers.  Untimed prelude: the user's data -- this rank's CSR (split
       into the diagonal / off-diagonal blocks the reference's MatCreateMPIAIJWithSplitArrays takes when there are several
       ranks), the right-hand side b = A*1 and a result buffer -- is put into PINNED host memory.  Timed: Mat creation on those
       arrays (zero-copy adoption by PETSc, host->device mirror by the plugin), Vec creation on the user's arrays
       (VecPlaceArray), KSP set-up, then steps x [b declared modified on the host -> H2D, KSPSolve one GMRES cycle,
       x device->host into the user's buffer]. */
    PetscInt    *hi, *hj, *hoi = NULL, *hoj = NULL, m = nloc;
    PetscScalar *ha, *hoa = NULL, *hb, *hx;
    const PetscScalar *rb;
    Mat          Ad = A, Ao = NULL, A2;
    const PetscInt *ci, *cj;
    const PetscScalar *ca;
    PetscInt     nr, nzA, nzB = 0;
    PetscBool    done;
    Vec          b2, x2;
    KSP          k2;
    double       t0, t_mat, t_vec, t_first = 0, t_solve = 0, t_d2h = 0, ms2, csum = 0;
    long long    ha0, da0, ha1, da1, hs0, ds0;
    const PetscInt *garray = NULL;

    if (SIZE > 1) PetscCall(MatMPIAIJB200GetSeqAIJ(A, &Ad, &Ao, &garray));
    PetscCall(MatGetRowIJ(Ad, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
    PetscCall(MatSeqAIJGetArrayRead(Ad, &ca));
    nzA = ci[m];
    CB(b200MallocHost((void **)&hi, sizeof(PetscInt) * ((size_t)m + 1)));
    CB(b200MallocHost((void **)&hj, sizeof(PetscInt) * ((size_t)nzA + 1)));
    CB(b200MallocHost((void **)&ha, sizeof(PetscScalar) * ((size_t)nzA + 1)));
    CB(b200MallocHost((void **)&hb, sizeof(PetscScalar) * (size_t)m));
    CB(b200MallocHost((void **)&hx, sizeof(PetscScalar) * (size_t)m));
    PetscCall(PetscArraycpy(hi, ci, m + 1));
    PetscCall(PetscArraycpy(hj, cj, nzA));
    PetscCall(PetscArraycpy(ha, ca, nzA));
    PetscCall(MatSeqAIJRestoreArrayRead(Ad, &ca));
    PetscCall(MatRestoreRowIJ(Ad, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
    if (Ao) { /* off-diagonal block back to GLOBAL columns, as a user would hold it */
      PetscCall(MatGetRowIJ(Ao, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
      PetscCall(MatSeqAIJGetArrayRead(Ao, &ca));
      nzB = ci[m];
      CB(b200MallocHost((void **)&hoi, sizeof(PetscInt) * ((size_t)m + 1)));
      CB(b200MallocHost((void **)&hoj, sizeof(PetscInt) * ((size_t)nzB + 1)));
      CB(b200MallocHost((void **)&hoa, sizeof(PetscScalar) * ((size_t)nzB + 1)));
      PetscCall(PetscArraycpy(hoi, ci, m + 1));
      for (PetscInt k = 0; k < nzB; k++) hoj[k] = garray[cj[k]];
      PetscCall(PetscArraycpy(hoa, ca, nzB));
      PetscCall(MatSeqAIJRestoreArrayRead(Ao, &ca));
      PetscCall(MatRestoreRowIJ(Ao, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
    }
    PetscCall(VecGetArrayRead(b, &rb));
    PetscCall(PetscArraycpy(hb, rb, m));
    PetscCall(VecRestoreArrayRead(b, &rb));
    PetscCall(VecDestroy(&x));
    PetscCall(VecDestroy(&b));
    PetscCall(MatDestroy(&A)); /* frees the device copy: the timed region starts from host data only */

    /* one-time set-up (host clock; reported, not part of the per-step metric -- the reference arm does not time its
       MatCreateSeqAIJWithArrays / first-solve set-up either): matrix adoption + host->device mirror, vectors on the user's
       buffers, KSP/PC set-up and one warm-up solve (Krylov basis allocation) */
    CB(b200TransferCounters(&hs0, &ds0));
    CB(b200Synchronize(H));
    PetscCall(PetscTime(&t0));
    if (SIZE == 1) {
      PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, m, hi, hj, ha, &A2));
      PetscCall(MatSetType(A2, "aijb200"));
    } else PetscCall(MatCreateMPIAIJB200WithSplitArrays(m, m * SIZE, m * RANK, hi, hj, ha, hoi, hoj, hoa, &A2));
    PetscCall(PetscTime(&t_mat));
    t_mat -= t0;
    PetscCall(MatCreateVecs(A2, &x2, &b2));
    PetscCall(VecPlaceArray(b2, hb)); /* the user's buffers ARE the host storage of the vectors */
    PetscCall(VecPlaceArray(x2, hx));
    PetscCall(KSPCreate(PETSC_COMM_SELF, &k2));
    PetscCall(KSPSetOperators(k2, A2, A2));
    PetscCall(KSPGetPC(k2, &pc));
    PetscCall(PCSetType(pc, PCJACOBI));
    PetscCall(KSPSetFromOptions(k2));
    PetscCall(KSPSetTolerances(k2, PETSC_CURRENT, PETSC_CURRENT, PETSC_CURRENT, restart));
    PetscCall(PetscTime(&t_vec));
    t_vec -= t0 + t_mat;
    PetscCall(KSPSolve(k2, b2, x2)); /* first solve: matrix host->device (11.8 GB at 512^3), SpMV plan, PC set-up, basis allocation */
    CB(b200Synchronize(H));
    PetscCall(PetscTime(&t_first));
    t_first -= t0 + t_mat + t_vec;
    CB(b200TransferCounters(&ha0, &da0));
    /* timed region (CUDA events, barrier + synchronise both sides, max over ranks): every step moves its input from the user's
       pinned host buffer to the device and its result back */
    PetscCall(TimerStart(&T));
    for (PetscInt s = 0; s < steps; s++) {
      PetscScalar       *wb;
      const PetscScalar *rx;
      double             ta, tb2, tc;
      PetscCall(PetscTime(&ta));
      PetscCall(VecGetArray(b2, &wb)); /* this step's right-hand side arrives in the user's host buffer */
      PetscCall(VecRestoreArray(b2, &wb));
      PetscCall(KSPSolve(k2, b2, x2));
      CB(b200Synchronize(H));
      PetscCall(PetscTime(&tb2));
      PetscCall(VecGetArrayRead(x2, &rx)); /* device -> the user's result buffer */
      csum = 0;
      for (int q = 0; q < 1000 && q < m; q++) csum += rx[q];
      PetscCall(VecRestoreArrayRead(x2, &rx));
      PetscCall(PetscTime(&tc));
      t_solve += tb2 - ta;
      t_d2h += tc - tb2;
    }
    PetscCall(TimerStop(&T, &ms2));
    CB(b200TransferCounters(&ha1, &da1));
    {
      double setup_ms = 1e3 * (t_mat + t_vec + t_first);
      emit("{\"kind\":\"e2e\",\"bench\":\"gmres7\",\"n_ranks\":%d,\"steps\":%d,\"iterations\":%d,\"ms_total\":%.3f,\"ms_per_step\":%.3f,\"iterations_per_sec\":%.4f,\"h2d_bytes_per_step\":%lld,\"d2h_bytes_per_step\":%lld,"
           "\"phases_ms\":{\"solves_incl_b_h2d\":%.2f,\"x_d2h\":%.2f},\"setup_ms\":{\"total\":%.2f,\"mat_create_host_adoption\":%.2f,\"vec_ksp_create\":%.2f,\"first_solve_matrix_h2d_plan_pcsetup_basis\":%.2f,\"h2d_bytes\":%lld,\"d2h_bytes\":%lld},"
           "\"iterations_per_sec_incl_one_time_setup\":%.4f,\"x_checksum\":%.15e}",
           SIZE, (int)steps, (int)(restart * steps), ms2, ms2 / steps, restart * steps / (ms2 * 1e-3), (ha1 - ha0) / steps, (da1 - da0) / steps, 1e3 * t_solve, 1e3 * t_d2h, setup_ms, 1e3 * t_mat, 1e3 * t_vec, 1e3 * t_first,
           ha0 - hs0, da0 - ds0, restart * steps / ((ms2 + setup_ms) * 1e-3), csum);
    }
    PetscCall(KSPDestroy(&k2));
    PetscCall(VecResetArray(b2));
    PetscCall(VecResetArray(x2));
    PetscCall(VecDestroy(&b2));
    PetscCall(VecDestroy(&x2));
    PetscCall(MatDestroy(&A2));
    CB(b200FreeHost(hi)); CB(b200FreeHost(hj)); CB(b200FreeHost(ha)); CB(b200FreeHost(hb)); CB(b200FreeHost(hx));
    if (hoi) { CB(b200FreeHost(hoi)); CB(b200FreeHost(hoj)); CB(b200FreeHost(hoa)); }
  } else {
    PetscCall(VecDestroy(&x));
    PetscCall(VecDestroy(&b));
    PetscCall(MatDestroy(&A));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* configs[2]: 27-point 256^3, KSPCG + PCILU(0) (MatSolverType b200: device factorisation + level-scheduled sweeps) */
static PetscErrorCode BenchCg27(void)
{
  PetscInt n = 256, N, its;
  int64_t  nnz;
  int     *d_i, *d_j;
  double  *d_a, t0, t1, first_s, solve_ms, spmv_ms, pc_ms, err;
  Mat      A;
  Vec      x, b, u, y;
  KSP      ksp;
  PC       pc;
  Timer    T = {0};
  KSPConvergedReason reason;
  PetscReal enorm;

  PetscFunctionBeginUser;
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-n", &n, NULL));
  N = n * n * n;
  CB(b200GenLaplace27Nnz((int)n, &nnz));
  CB(b200Malloc(H, (void **)&d_i, sizeof(int) * ((size_t)N + 1)));
  CB(b200Malloc(H, (void **)&d_j, sizeof(int) * (size_t)nnz));
  CB(b200Malloc(H, (void **)&d_a, sizeof(double) * (size_t)nnz));
  CB(b200GenLaplace27(H, (int)n, d_i, d_j, d_a));
  PetscCall(PetscOptionsInsertString(NULL, "-mat_no_inode"));
  PetscCall(MatCreateSeqAIJB200WithDeviceArrays(N, N, d_i, d_j, d_a, &A));
  PetscCall(MatCreateVecs(A, &x, &b));
  PetscCall(VecDuplicate(x, &u));
  PetscCall(VecDuplicate(x, &y));
  PetscCall(VecSet(u, 1.0));
  PetscCall(MatMult(A, u, b));
  TIME_LOOP(spmv_ms, 3, 20, PetscCall(MatMult(A, u, y)));
  PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
  PetscCall(KSPSetOperators(ksp, A, A));
  PetscCall(KSPSetType(ksp, KSPCG));
  PetscCall(KSPGetPC(ksp, &pc));
  PetscCall(PCSetType(pc, PCILU));
  PetscCall(PCFactorSetMatSolverType(pc, "b200"));
  PetscCall(KSPSetTolerances(ksp, 1e-8, PETSC_CURRENT, PETSC_CURRENT, PETSC_CURRENT));
  PetscCall(KSPSetFromOptions(ksp));
  PetscCall(PetscTime(&t0));
  PetscCall(KSPSolve(ksp, b, x)); /* includes PCSetUp: symbolic (host pattern + level schedule) and the device numeric factorisation */
  CB(b200Synchronize(H));
  PetscCall(PetscTime(&t1));
  first_s = t1 - t0;
  PetscCall(TimerStart(&T));
  PetscCall(KSPSolve(ksp, b, x));
  PetscCall(TimerStop(&T, &solve_ms));
  PetscCall(KSPGetIterationNumber(ksp, &its));
  PetscCall(KSPGetConvergedReason(ksp, &reason));
  TIME_LOOP(pc_ms, 2, 10, PetscCall(PCApply(pc, b, y)));
  PetscCall(VecAXPY(x, -1.0, u));
  PetscCall(VecNorm(x, NORM_INFINITY, &enorm));
  err = (double)enorm;
  emit("{\"kind\":\"solve\",\"bench\":\"cg27\",\"n\":%d,\"rows\":%d,\"nnz\":%lld,\"iterations\":%d,\"reason\":%d,\"max_error\":%.3e,\"first_solve_incl_setup_s\":%.3f,\"solve_ms\":%.3f,"
       "\"ms_per_iteration\":%.4f,\"iterations_per_sec\":%.3f,\"spmv_ms\":%.5f,\"spmv_algorithmic_bytes\":%lld,\"pcapply_ilu_ms\":%.5f,\"sptrsv_algorithmic_bytes\":%lld}",
       (int)n, (int)N, (long long)nnz, (int)its, (int)reason, err, first_s, solve_ms, solve_ms / PetscMax(its, 1), its / (solve_ms * 1e-3), spmv_ms, (long long)(nnz * 12 + (int64_t)N * 20), pc_ms,
       (long long)(nnz * 12 + (int64_t)N * (4 + 4 + 4 + 4 + 8 * 3)));
  PetscCall(KSPDestroy(&ksp));
  PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&u)); PetscCall(VecDestroy(&y));
  PetscCall(MatDestroy(&A));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* configs[4]: random CSR, fixed row length d, MatMult only */
static PetscErrorCode BenchRand(void)
{
  PetscInt n = 10000000, d = 32;
  int64_t  nnz;
  int     *d_i, *d_j;
  double  *d_a, ms;
  Mat      A;
  Vec      x, y;
  PetscFunctionBeginUser;
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-rand_n", &n, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-rand_d", &d, NULL));
  nnz = (int64_t)n * d;
  PetscCheck(nnz < 2147483647LL, PETSC_COMM_SELF, PETSC_ERR_SUP, "n*d exceeds the 32-bit PetscInt of this PETSc (SURVEY 7 hard part 2): reduce -rand_n");
  CB(b200Malloc(H, (void **)&d_i, sizeof(int) * ((size_t)n + 1)));
  CB(b200Malloc(H, (void **)&d_j, sizeof(int) * (size_t)nnz));
  CB(b200Malloc(H, (void **)&d_a, sizeof(double) * (size_t)nnz));
  CB(b200GenRandomCsr(H, (int)n, (int)n, (int)d, (uint64_t)(20260923 + d), d_i, d_j, d_a));
  PetscCall(PetscOptionsInsertString(NULL, "-mat_no_inode"));
  PetscCall(MatCreateSeqAIJB200WithDeviceArrays(n, n, d_i, d_j, d_a, &A));
  PetscCall(MatCreateVecs(A, &x, &y));
  PetscCall(VecSet(x, 1.0));
  TIME_LOOP(ms, 3, 10, PetscCall(MatMult(A, x, y)));
  emit("{\"kind\":\"matmult\",\"bench\":\"rand\",\"n\":%d,\"d\":%d,\"nnz\":%lld,\"ms\":%.5f,\"algorithmic_bytes\":%lld,\"flops\":%lld}", (int)n, (int)d, (long long)nnz, ms, (long long)(nnz * 12 + (int64_t)n * 20),
       (long long)(2 * nnz - n));
  PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&y));
  PetscCall(MatDestroy(&A));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* configs[0]: ex2's operator assembled with MatSetValues exactly as ex2.c:70-92 does, KSP from the options */
static PetscErrorCode BenchEx2(void)
{
  PetscInt    m = 100, n = 100, its, Ii, J, i, j;
  PetscScalar v;
  Mat         A;
  Vec         x, b, u;
  KSP         ksp;
  double      ms;
  Timer       T = {0};
  PetscReal   rnorm, enorm;
  KSPConvergedReason reason;
  PetscFunctionBeginUser;
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-m", &m, NULL));
  PetscCall(PetscOptionsGetInt(NULL, NULL, "-n", &n, NULL));
  PetscCall(MatCreate(PETSC_COMM_SELF, &A));
  PetscCall(MatSetSizes(A, m * n, m * n, m * n, m * n));
  PetscCall(MatSetType(A, "aijb200"));
  PetscCall(MatSetFromOptions(A));
  PetscCall(MatSeqAIJSetPreallocation(A, 5, NULL));
  for (Ii = 0; Ii < m * n; Ii++) {
    v = -1.0; i = Ii / n; j = Ii - i * n;
    if (i > 0) { J = Ii - n; PetscCall(MatSetValues(A, 1, &Ii, 1, &J, &v, ADD_VALUES)); }
    if (i < m - 1) { J = Ii + n; PetscCall(MatSetValues(A, 1, &Ii, 1, &J, &v, ADD_VALUES)); }
    if (j > 0) { J = Ii - 1; PetscCall(MatSetValues(A, 1, &Ii, 1, &J, &v, ADD_VALUES)); }
    if (j < n - 1) { J = Ii + 1; PetscCall(MatSetValues(A, 1, &Ii, 1, &J, &v, ADD_VALUES)); }
    v = 4.0; PetscCall(MatSetValues(A, 1, &Ii, 1, &Ii, &v, ADD_VALUES));
  }
  PetscCall(MatAssemblyBegin(A, MAT_FINAL_ASSEMBLY));
  PetscCall(MatAssemblyEnd(A, MAT_FINAL_ASSEMBLY));
  PetscCall(MatCreateVecs(A, &x, &b));
  PetscCall(VecDuplicate(x, &u));
  PetscCall(VecSet(u, 1.0));
  PetscCall(MatMult(A, u, b));
  PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
  PetscCall(KSPSetOperators(ksp, A, A));
  PetscCall(KSPSetTolerances(ksp, 1.e-2 / ((m + 1) * (n + 1)), 1.e-50, PETSC_CURRENT, PETSC_CURRENT)); /* ex2.c:160 */
  PetscCall(KSPSetFromOptions(ksp));
  PetscCall(KSPSolve(ksp, b, x));
  PetscCall(TimerStart(&T));
  PetscCall(KSPSolve(ksp, b, x));
  PetscCall(TimerStop(&T, &ms));
  PetscCall(KSPGetIterationNumber(ksp, &its));
  PetscCall(KSPGetConvergedReason(ksp, &reason));
  PetscCall(KSPGetResidualNorm(ksp, &rnorm));
  PetscCall(VecAXPY(x, -1.0, u));
  PetscCall(VecNorm(x, NORM_2, &enorm));
  {
    KSPType kt;
    PCType  pt;
    PC      pc;
    PetscCall(KSPGetType(ksp, &kt));
    PetscCall(KSPGetPC(ksp, &pc));
    PetscCall(PCGetType(pc, &pt));
    emit("{\"kind\":\"solve\",\"bench\":\"ex2\",\"m\":%d,\"n\":%d,\"ksp_type\":\"%s\",\"pc_type\":\"%s\",\"iterations\":%d,\"reason\":%d,\"rnorm\":%.12e,\"error_norm\":%.6g,\"solve_ms\":%.4f,\"us_per_iteration\":%.3f}", (int)m, (int)n, kt, pt,
         (int)its, (int)reason, (double)rnorm, (double)enorm, ms, 1e3 * ms / PetscMax(its, 1));
  }
  PetscCall(KSPDestroy(&ksp));
  PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&u));
  PetscCall(MatDestroy(&A));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- multi-rank parity: <dir>/case_<name>/rank<r>/{meta.txt, ai.i32, aj.i32 (GLOBAL columns), aa.f64, x.f64, V.f64} in,
   results out; the Python side (bench.py --gpus N, tests) holds the oracle and compares */
static void *rd(const char *dir, const char *name, size_t bytes)
{
  char  p[4096];
  FILE *f;
  void *buf = malloc(bytes + 8);
  snprintf(p, sizeof p, "%s/%s", dir, name);
  f = fopen(p, "rb");
  if (!f || fread(buf, 1, bytes, f) != bytes) {
    fprintf(stderr, "b200_driver: cannot read %s\n", p);
    exit(3);
  }
  fclose(f);
  return buf;
}
static void wr(const char *dir, const char *name, const void *buf, size_t bytes)
{
  char  p[4096];
  FILE *f;
  snprintf(p, sizeof p, "%s/%s", dir, name);
  f = fopen(p, "wb");
  if (!f || fwrite(buf, 1, bytes, f) != bytes) {
    fprintf(stderr, "b200_driver: cannot write %s\n", p);
    exit(3);
  }
  fclose(f);
}
static PetscErrorCode WriteVec(const char *dir, const char *name, Vec v)
{
  const PetscScalar *a;
  PetscInt           n;
  PetscFunctionBeginUser;
  PetscCall(VecGetLocalSize(v, &n));
  PetscCall(VecGetArrayRead(v, &a));
  wr(dir, name, a, sizeof(PetscScalar) * (size_t)n);
  PetscCall(VecRestoreArrayRead(v, &a));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode WriteSeqAIJ(const char *dir, const char *prefix, Mat M)
{
  const PetscInt    *ci, *cj;
  const PetscScalar *ca;
  PetscInt           nr;
  PetscBool          done;
  char               nm[256];
  PetscFunctionBeginUser;
  PetscCall(MatGetRowIJ(M, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
  PetscCall(MatSeqAIJGetArrayRead(M, &ca));
  snprintf(nm, sizeof nm, "%s_i.i32", prefix); wr(dir, nm, ci, sizeof(PetscInt) * ((size_t)nr + 1));
  snprintf(nm, sizeof nm, "%s_j.i32", prefix); wr(dir, nm, cj, sizeof(PetscInt) * (size_t)ci[nr]);
  snprintf(nm, sizeof nm, "%s_a.f64", prefix); wr(dir, nm, ca, sizeof(PetscScalar) * (size_t)ci[nr]);
  PetscCall(MatSeqAIJRestoreArrayRead(M, &ca));
  PetscCall(MatRestoreRowIJ(M, 0, PETSC_FALSE, PETSC_FALSE, &nr, &ci, &cj, &done));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode ParityCase(const char *cdir)
{
  char         dir[4096];
  int          m, N, r0, nv, dev, solve;
  long         nz;
  FILE        *f;
  PetscInt    *ai, *aj, ec;
  PetscScalar *aa, *xl, *Vl, dots[64], sc[8];
  Mat          A, Ad, Ao;
  Vec          x, y, z, dg, *V;
  const PetscInt *garray;
  PetscReal    nrm[3];
  char         opts[2048] = "";

  PetscFunctionBeginUser;
  snprintf(dir, sizeof dir, "%s/rank%d", cdir, RANK);
  {
    char p[4200];
    snprintf(p, sizeof p, "%s/meta.txt", dir);
    f = fopen(p, "r");
    PetscCheck(f && fscanf(f, "%d %d %d %ld %d %d %d", &m, &N, &r0, &nz, &nv, &dev, &solve) == 7, PETSC_COMM_SELF, PETSC_ERR_FILE_READ, "bad %s", p);
    if (solve && fgets(opts, sizeof opts, f) && fgets(opts, sizeof opts, f)) opts[strcspn(opts, "\n")] = 0;
    fclose(f);
  }
  ai = (PetscInt *)rd(dir, "ai.i32", sizeof(PetscInt) * ((size_t)m + 1));
  aj = (PetscInt *)rd(dir, "aj.i32", sizeof(PetscInt) * (size_t)nz);
  aa = (PetscScalar *)rd(dir, "aa.f64", sizeof(PetscScalar) * (size_t)nz);
  xl = (PetscScalar *)rd(dir, "x.f64", sizeof(PetscScalar) * (size_t)m);
  Vl = (PetscScalar *)rd(dir, "V.f64", sizeof(PetscScalar) * (size_t)m * nv);
  if (dev) { /* device-resident input: the column split runs on the device */
    int    *d_i, *d_j;
    double *d_a;
    CB(b200Malloc(H, (void **)&d_i, sizeof(int) * ((size_t)m + 1)));
    CB(b200Malloc(H, (void **)&d_j, sizeof(int) * ((size_t)nz + 1)));
    CB(b200Malloc(H, (void **)&d_a, sizeof(double) * ((size_t)nz + 1)));
    CB(b200MemcpyHtoD(H, d_i, ai, sizeof(int) * ((size_t)m + 1)));
    CB(b200MemcpyHtoD(H, d_j, aj, sizeof(int) * (size_t)nz));
    CB(b200MemcpyHtoD(H, d_a, aa, sizeof(double) * (size_t)nz));
    PetscCall(MatCreateMPIAIJB200WithArrays(m, N, r0, d_i, d_j, d_a, &A));
    CB(b200Free(H, d_i)); CB(b200Free(H, d_j)); CB(b200Free(H, d_a));
  } else PetscCall(MatCreateMPIAIJB200WithArrays(m, N, r0, ai, aj, aa, &A));
  PetscCall(MatMPIAIJB200GetSeqAIJ(A, &Ad, &Ao, &garray));
  PetscCall(MatGetSize(Ao, NULL, &ec));
  wr(dir, "out_garray.i32", garray, sizeof(PetscInt) * (size_t)ec);
  PetscCall(WriteSeqAIJ(dir, "out_A", Ad));
  PetscCall(WriteSeqAIJ(dir, "out_B", Ao));
  PetscCall(MatCreateVecs(A, &x, &y));
  PetscCall(VecDuplicate(y, &z));
  PetscCall(VecDuplicate(y, &dg));
  {
    PetscScalar *xa;
    PetscCall(VecGetArrayWrite(x, &xa));
    PetscCall(PetscArraycpy(xa, xl, m));
    PetscCall(VecRestoreArrayWrite(x, &xa));
  }
  PetscCall(MatMult(A, x, y));
  PetscCall(WriteVec(dir, "out_mult.f64", y));
  PetscCall(MatMultAdd(A, x, y, z));
  PetscCall(WriteVec(dir, "out_multadd.f64", z));
  PetscCall(MatMultTranspose(A, x, z));
  PetscCall(WriteVec(dir, "out_multtr.f64", z));
  PetscCall(MatGetDiagonal(A, dg));
  PetscCall(WriteVec(dir, "out_diag.f64", dg));
  { /* fused Jacobi through PCApplyBAorAB against the unfused composition */
    PC  pc;
    Vec w1, w2, wk;
    PetscCall(VecDuplicate(y, &w1)); PetscCall(VecDuplicate(y, &w2)); PetscCall(VecDuplicate(y, &wk));
    PetscCall(PCCreate(PETSC_COMM_SELF, &pc));
    PetscCall(PCSetOperators(pc, A, A));
    PetscCall(PCSetType(pc, PCJACOBI));
    PetscCall(PCSetUp(pc));
    PetscCall(PCApplyBAorAB(pc, PC_LEFT, x, w1, wk));
    PetscCall(WriteVec(dir, "out_jacobi_fused.f64", w1));
    PetscCall(MatMult(A, x, wk));
    PetscCall(PCApply(pc, wk, w2));
    PetscCall(WriteVec(dir, "out_jacobi_unfused.f64", w2));
    PetscCall(PCDestroy(&pc));
    PetscCall(VecDestroy(&w1)); PetscCall(VecDestroy(&w2)); PetscCall(VecDestroy(&wk));
  }
  PetscCall(VecDuplicateVecs(x, nv, &V));
  for (int j = 0; j < nv; j++) {
    PetscScalar *va;
    PetscCall(VecGetArrayWrite(V[j], &va));
    PetscCall(PetscArraycpy(va, Vl + (size_t)j * m, m));
    PetscCall(VecRestoreArrayWrite(V[j], &va));
  }
  PetscCall(VecMDot(x, nv, V, dots));
  wr(dir, "out_mdot.f64", dots, sizeof(PetscScalar) * (size_t)nv);
  PetscCall(VecNorm(x, NORM_2, &nrm[0]));
  PetscCall(VecNorm(x, NORM_1, &nrm[1]));
  PetscCall(VecNorm(x, NORM_INFINITY, &nrm[2]));
  PetscCall(VecDot(x, V[0], &sc[3]));
  sc[0] = nrm[0]; sc[1] = nrm[1]; sc[2] = nrm[2];
  for (int j = 0; j < nv; j++) dots[j] = -dots[j];
  PetscCall(VecMAXPY(x, nv, dots, V));
  PetscCall(VecNorm(x, NORM_2, &nrm[0])); /* fused MAXPY+norm, all-reduced inside VecNorm */
  sc[4] = nrm[0];
  PetscCall(VecSum(y, &sc[5]));
  wr(dir, "out_scalars.f64", sc, sizeof(PetscScalar) * 6);
  PetscCall(WriteVec(dir, "out_maxpy.f64", x));
  PetscCall(VecDestroyVecs(nv, &V));
  if (solve) { /* b = A*1, KSP from the options line of meta.txt (e.g. ex2_2.out: GMRES + bjacobi/ILU(0)) */
    KSP        ksp;
    Vec        b, u, s;
    PetscInt   its, nh;
    PetscReal *hist;
    const PetscReal *h;
    PetscScalar info[2];
    KSPConvergedReason reason;
    PetscCall(PetscOptionsInsertString(NULL, opts));
    PetscCall(VecDuplicate(y, &b)); PetscCall(VecDuplicate(y, &u)); PetscCall(VecDuplicate(y, &s));
    PetscCall(VecSet(u, 1.0));
    PetscCall(MatMult(A, u, b));
    PetscCall(KSPCreate(PETSC_COMM_SELF, &ksp));
    PetscCall(KSPSetOperators(ksp, A, A));
    PetscCall(PetscMalloc1(10000, &hist));
    PetscCall(KSPSetResidualHistory(ksp, hist, 10000, PETSC_TRUE));
    PetscCall(KSPSetFromOptions(ksp));
    PetscCall(KSPSolve(ksp, b, s));
    PetscCall(KSPGetIterationNumber(ksp, &its));
    PetscCall(KSPGetConvergedReason(ksp, &reason));
    PetscCall(KSPGetResidualHistory(ksp, &h, &nh));
    wr(dir, "out_hist.f64", h, sizeof(PetscReal) * (size_t)nh);
    info[0] = its; info[1] = reason;
    wr(dir, "out_ksp.f64", info, sizeof info);
    PetscCall(WriteVec(dir, "out_sol.f64", s));
    PetscCall(KSPDestroy(&ksp));
    PetscCall(PetscFree(hist));
    PetscCall(VecDestroy(&b)); PetscCall(VecDestroy(&u)); PetscCall(VecDestroy(&s));
    PetscCall(PetscOptionsClear(NULL));
  }
  PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&y)); PetscCall(VecDestroy(&z)); PetscCall(VecDestroy(&dg));
  PetscCall(MatDestroy(&A));
  free(ai); free(aj); free(aa); free(xl); free(Vl);
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode Parity(const char *dir)
{
  char  p[4200], name[256];
  FILE *f;
  PetscFunctionBeginUser;
  snprintf(p, sizeof p, "%s/cases.txt", dir);
  f = fopen(p, "r");
  PetscCheck(f, PETSC_COMM_SELF, PETSC_ERR_FILE_OPEN, "cannot open %s", p);
  while (fscanf(f, "%255s", name) == 1) {
    snprintf(p, sizeof p, "%s/case_%s", dir, name);
    PetscCall(ParityCase(p));
    if (RANK == 0) printf("parity case %s done\n", name);
  }
  fclose(f);
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* entry point shared by the executable and the in-process use from Python (ctypes: b200_driver_main(argc, argv)) */
PETSC_EXTERN int b200_driver_main(int argc, char **argv)
{
  char      bench[64] = "", pdir[4096] = "", jout[4096] = "";
  PetscBool flg, pflg, jflg;
  void     *h;

  PetscFunctionBeginUser;
  PetscCall(PetscInitialize(&argc, &argv, NULL, "b200_driver: BASELINE workloads on the b200 PETSc types"));
  PetscCall(PetscDLLibraryRegister_petscb200plugin());
  PetscCall(PetscB200GetHandle(&h, &RANK, &SIZE));
  H = (b200Handle)h;
  PetscCall(PetscOptionsGetString(NULL, NULL, "-bench", bench, sizeof bench, &flg));
  PetscCall(PetscOptionsGetString(NULL, NULL, "-parity", pdir, sizeof pdir, &pflg));
  PetscCall(PetscOptionsGetString(NULL, NULL, "-json_out", jout, sizeof jout, &jflg));
  JOUT = (jflg && RANK == 0) ? fopen(jout, "a") : NULL;
  PetscCall(PetscOptionsInsertString(NULL, "-mat_type aijb200 -vec_type b200"));
  if (pflg) PetscCall(Parity(pdir));
  if (flg) {
    if (!strcmp(bench, "gmres7")) PetscCall(BenchGmres7());
    else if (!strcmp(bench, "cg27")) PetscCall(BenchCg27());
    else if (!strcmp(bench, "rand")) PetscCall(BenchRand());
    else if (!strcmp(bench, "ex2")) PetscCall(BenchEx2());
    else SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_UNKNOWN_TYPE, "unknown -bench %s", bench);
  }
  if (JOUT) fclose(JOUT);
  JOUT = NULL;
  PetscCall(PetscFinalize());
  return 0;
}

#if !defined(B200_DRIVER_NO_MAIN)
int main(int argc, char **argv)
{
  return b200_driver_main(argc, argv);
}
#endif
