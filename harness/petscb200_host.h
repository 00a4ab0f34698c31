/*
 * petscb200_host.h -- stand-alone C host mirror of the part of PETSc's Mat/Vec/KSP/PC interface that the Krylov hot path
 * uses (libpetscb200host.so, plain C, calls the sm_100a kernels through include/petscb200.h).
 *
 * Why it exists: the reference's toolchain (C) is present but an MPI and a configured PETSc are not part of the GPU box,
 * so the path is exercised through a mirror with the SAME names, argument meaning and error behaviour as the reference's
 * public API -- a driver written against petscksp.h for this path (ex2.c, bench_kspsolve.c style) compiles against this
 * header unchanged, and `-mat_type aijb200 -vec_type b200 -ksp_type gmres -pc_type jacobi` mean what they mean there.
 * The real-PETSc plugin that registers the same types with MatRegister/VecRegister/PCRegister is petsc_plugin/.
 * This library must not be loaded into a process that also loads libpetsc (identical symbol names).
 *
 * Each declaration cites the reference interface it mirrors (path:line under the PETSc source root).
 * Types: seqb200/mpib200 (Vec), seqaijb200/mpiaijb200 (Mat), gmres/cg/preonly (KSP), none/jacobi/bjacobi/ilu (PC).
 * One process per GPU; ranks = torchrun's RANK/WORLD_SIZE; NCCL replaces MPI (PetscB200CommInit).
 */
#ifndef PETSCB200_HOST_H
#define PETSCB200_HOST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int      PetscErrorCode; /* include/petscsystypes.h: 0 = PETSC_SUCCESS, PETSC_ERR_* otherwise */
typedef int      PetscInt;       /* 32-bit indices (default configuration) */
typedef int64_t  PetscInt64;
typedef int      PetscMPIInt;
typedef double   PetscScalar;
typedef double   PetscReal;
typedef int      PetscBool;
typedef int      MPI_Comm;       /* as MPIUNI does (include/petsc/mpiuni/mpi.h) */
typedef int      PetscMemType;
typedef int64_t  PetscCount;
#define PETSC_TRUE  1
#define PETSC_FALSE 0
#define PETSC_SUCCESS 0
#define PETSC_COMM_WORLD 1
#define PETSC_COMM_SELF  2
#define PETSC_DECIDE   (-1)
#define PETSC_DETERMINE PETSC_DECIDE
#define PETSC_DEFAULT  (-2)
#define PETSC_CURRENT  (-3)
#define PETSC_MEMTYPE_HOST 0
#define PETSC_MEMTYPE_CUDA 1

typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_PC  *PC;
typedef struct _p_KSP *KSP;
typedef const char *VecType;
typedef const char *MatType;
typedef const char *PCType;
typedef const char *KSPType;

#define VECSEQB200    "seqb200"
#define VECMPIB200    "mpib200"
#define VECB200       "b200"
#define VECSTANDARD   "standard"     /* accepted as an alias of b200: there is no host vector type in this library */
#define MATSEQAIJB200 "seqaijb200"
#define MATMPIAIJB200 "mpiaijb200"
#define MATAIJB200    "aijb200"
#define MATAIJ        "aij"          /* alias of aijb200 */
#define PCNONE    "none"
#define PCJACOBI  "jacobi"
#define PCBJACOBI "bjacobi"
#define PCILU     "ilu"
#define KSPGMRES   "gmres"
#define KSPCG      "cg"
#define KSPPIPECG  "pipecg" /* cg/pipecg/pipecg.c: the three reductions of an iteration fused into one kernel + one synchronisation */
#define KSPPREONLY "preonly"

typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;   /* include/petscvec.h */
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MAT_SYMMETRIC = 1, MAT_SPD = 2, MAT_OPTION_OTHER = 99 } MatOption;
typedef enum {                                                                              /* include/petscksp.h */
  KSP_CONVERGED_RTOL = 2, KSP_CONVERGED_ATOL = 3, KSP_CONVERGED_ITS = 4, KSP_CONVERGED_HAPPY_BREAKDOWN = 7,
  KSP_DIVERGED_NULL = -2, KSP_DIVERGED_ITS = -3, KSP_DIVERGED_DTOL = -4, KSP_DIVERGED_BREAKDOWN = -5,
  KSP_DIVERGED_INDEFINITE_PC = -8, KSP_DIVERGED_NANORINF = -9, KSP_DIVERGED_INDEFINITE_MAT = -10,
  KSP_DIVERGED_PC_FAILED = -11, KSP_CONVERGED_ITERATING = 0
} KSPConvergedReason;
typedef enum { KSP_GMRES_CGS_REFINE_NEVER, KSP_GMRES_CGS_REFINE_IFNEEDED, KSP_GMRES_CGS_REFINE_ALWAYS } KSPGMRESCGSRefinementType;

/* ---- Sys: init, options database, plugin registries, error text ---- */
PetscErrorCode PetscInitialize(int *argc, char ***args, const char file[], const char help[]);   /* src/sys/objects/pinit.c */
PetscErrorCode PetscInitializeNoArguments(void);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscOptionsSetValue(void *options, const char name[], const char value[]);       /* src/sys/objects/options.c */
PetscErrorCode PetscOptionsInsertString(void *options, const char in_str[]);
PetscErrorCode PetscOptionsClearValue(void *options, const char name[]);
PetscErrorCode PetscOptionsClear(void *options);
PetscErrorCode PetscOptionsGetInt(void *options, const char pre[], const char name[], PetscInt *ivalue, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(void *options, const char pre[], const char name[], PetscReal *dvalue, PetscBool *set);
PetscErrorCode PetscOptionsGetBool(void *options, const char pre[], const char name[], PetscBool *bvalue, PetscBool *set);
PetscErrorCode PetscOptionsGetString(void *options, const char pre[], const char name[], char str[], size_t len, PetscBool *set);
PetscErrorCode MPI_Comm_rank(MPI_Comm comm, PetscMPIInt *rank);
PetscErrorCode MPI_Comm_size(MPI_Comm comm, PetscMPIInt *size);
const char    *PetscB200GetLastErrorMessage(void);   /* what PetscError() would have printed (message + call stack) */
/* NCCL bootstrap replacing MPI_Init: rank 0 calls PetscB200CommGetUniqueId, the 128 bytes travel out of band (bench.py:
   torch.distributed gloo broadcast), every rank calls PetscB200CommInit.  Without it WORLD has size 1. */
PetscErrorCode PetscB200CommGetUniqueId(void *id128);
PetscErrorCode PetscB200CommInit(PetscMPIInt rank, PetscMPIInt size, const void *id128);
PetscErrorCode PetscB200SetDevice(int device);        /* before PetscInitialize; default LOCAL_RANK % ndev (cupmdevice.cxx:293) */
PetscErrorCode PetscB200GetHandle(void **b200Handle); /* the PetscCUBLASGetHandle analogue */
/* registries = the reference's plugin API (matreg.c:293, vecreg.c:252, pcregis.c, itregis.c) */
PetscErrorCode VecRegister(const char sname[], PetscErrorCode (*function)(Vec));
PetscErrorCode MatRegister(const char sname[], PetscErrorCode (*function)(Mat));
PetscErrorCode MatRegisterRootName(const char rname[], const char sname[], const char mname[]);   /* matreg.c:328 */
PetscErrorCode PCRegister(const char sname[], PetscErrorCode (*function)(PC));
PetscErrorCode KSPRegister(const char sname[], PetscErrorCode (*function)(KSP));

/* ---- Vec (src/vec/vec/interface/{vector.c,rvector.c,vecreg.c}) ---- */
PetscErrorCode VecCreate(MPI_Comm comm, Vec *vec);
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N);
PetscErrorCode VecSetType(Vec vec, VecType newType);                                              /* vecreg.c:51 */
PetscErrorCode VecSetFromOptions(Vec vec);                                                        /* -vec_type */
PetscErrorCode VecGetType(Vec vec, VecType *type);
PetscErrorCode VecDuplicate(Vec v, Vec *newv);
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec *V[]);                                     /* contiguous slab, bvec2.c:670 */
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecDestroyVecs(PetscInt m, Vec *vv[]);
PetscErrorCode VecGetSize(Vec x, PetscInt *size);
PetscErrorCode VecGetLocalSize(Vec x, PetscInt *size);
PetscErrorCode VecGetOwnershipRange(Vec x, PetscInt *low, PetscInt *high);
PetscErrorCode VecSet(Vec x, PetscScalar alpha);
PetscErrorCode VecZeroEntries(Vec x);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecSwap(Vec x, Vec y);
PetscErrorCode VecScale(Vec x, PetscScalar alpha);                                                /* rvector.c:1010 */
PetscErrorCode VecShift(Vec v, PetscScalar shift);
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);                                          /* rvector.c:663 */
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);
PetscErrorCode VecAXPBY(Vec y, PetscScalar alpha, PetscScalar beta, Vec x);
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y);
PetscErrorCode VecMAXPY(Vec y, PetscInt nv, const PetscScalar alpha[], Vec x[]);                  /* rvector.c:1365 */
PetscErrorCode VecMAXPBY(Vec y, PetscInt nv, const PetscScalar alpha[], PetscScalar beta, Vec x[]); /* rvector.c:1394 */
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);                                            /* rvector.c:108 */
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecMDot(Vec x, PetscInt nv, const Vec y[], PetscScalar val[]);                     /* rvector.c:1300 */
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val);                                     /* rvector.c:199, with the norm cache */
PetscErrorCode VecNormalize(Vec x, PetscReal *val);                                               /* rvector.c:289 */
PetscErrorCode VecSum(Vec v, PetscScalar *sum);
PetscErrorCode VecMax(Vec x, PetscInt *p, PetscReal *val);
PetscErrorCode VecMin(Vec x, PetscInt *p, PetscReal *val);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecReciprocal(Vec vec);
PetscErrorCode VecSetValues(Vec x, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode iora); /* owned entries only */
PetscErrorCode VecAssemblyBegin(Vec vec);
PetscErrorCode VecAssemblyEnd(Vec vec);
PetscErrorCode VecGetArray(Vec x, PetscScalar **a);                                               /* host mirror, offload mask protocol */
PetscErrorCode VecRestoreArray(Vec x, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec x, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec x, const PetscScalar **a);
PetscErrorCode VecGetArrayWrite(Vec x, PetscScalar **a);
PetscErrorCode VecRestoreArrayWrite(Vec x, PetscScalar **a);
PetscErrorCode VecGetArrayAndMemType(Vec x, PetscScalar **a, PetscMemType *mtype);                /* rvector.c:2365: device pointer */
PetscErrorCode VecRestoreArrayAndMemType(Vec x, PetscScalar **a);
PetscErrorCode VecGetArrayReadAndMemType(Vec x, const PetscScalar **a, PetscMemType *mtype);
PetscErrorCode VecRestoreArrayReadAndMemType(Vec x, const PetscScalar **a);
PetscErrorCode VecGetLocalVector(Vec v, Vec w);                                                   /* aliases the local device array */
PetscErrorCode VecRestoreLocalVector(Vec v, Vec w);
PetscErrorCode VecGetLocalVectorRead(Vec v, Vec w);
PetscErrorCode VecRestoreLocalVectorRead(Vec v, Vec w);
PetscErrorCode VecCreateSeqWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, const PetscScalar array[], Vec *V); /* bvec2.c: the user's HOST array is the host storage */
PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, PetscInt N, const PetscScalar array[], Vec *V); /* pbvec.c */
PetscErrorCode VecPlaceArray(Vec vec, const PetscScalar array[]);                                                         /* rvector.c:2593 */
PetscErrorCode VecResetArray(Vec vec);

/* ---- Mat (src/mat/interface/{matrix.c,matreg.c}, impls/aij) ---- */
PetscErrorCode MatCreate(MPI_Comm comm, Mat *A);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetType(Mat mat, MatType matype);                                               /* matreg.c:107 */
PetscErrorCode MatSetFromOptions(Mat B);                                                          /* -mat_type */
PetscErrorCode MatGetType(Mat mat, MatType *type);
PetscErrorCode MatSetUp(Mat A);
PetscErrorCode MatSeqAIJSetPreallocation(Mat B, PetscInt nz, const PetscInt nnz[]);
PetscErrorCode MatMPIAIJSetPreallocation(Mat B, PetscInt d_nz, const PetscInt d_nnz[], PetscInt o_nz, const PetscInt o_nnz[]);
PetscErrorCode MatSetValues(Mat mat, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode addv); /* owned rows */
PetscErrorCode MatAssemblyBegin(Mat mat, MatAssemblyType type);
PetscErrorCode MatAssemblyEnd(Mat mat, MatAssemblyType type);                                     /* aij.c:1085 / mpiaij.c:823 */
PetscErrorCode MatSetOption(Mat mat, MatOption op, PetscBool flg);
PetscErrorCode MatCreateSeqAIJWithArrays(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt i[], PetscInt j[], PetscScalar a[], Mat *mat); /* aij.c (copies to device) */
PetscErrorCode MatSeqAIJSetPreallocationCSR(Mat B, const PetscInt i[], const PetscInt j[], const PetscScalar v[]);  /* aij.c:4023 */
PetscErrorCode MatMPIAIJSetPreallocationCSR(Mat B, const PetscInt i[], const PetscInt j[], const PetscScalar v[]);  /* local rows, global cols */
/* device-resident CSR hand-over (no host copy): local rows, GLOBAL columns; the matrix takes ownership of nothing and copies */
PetscErrorCode MatB200SetCSRDevice(Mat B, const PetscInt *d_i, const PetscInt *d_j, const PetscScalar *d_a);
PetscErrorCode MatGetSize(Mat mat, PetscInt *m, PetscInt *n);
PetscErrorCode MatGetLocalSize(Mat mat, PetscInt *m, PetscInt *n);
PetscErrorCode MatGetOwnershipRange(Mat mat, PetscInt *m, PetscInt *n);
PetscErrorCode MatCreateVecs(Mat mat, Vec *right, Vec *left);                                     /* matrix.c:10069 (defaultvectype) */
PetscErrorCode MatMult(Mat mat, Vec x, Vec y);                                                    /* matrix.c:2696 */
PetscErrorCode MatMultAdd(Mat mat, Vec v1, Vec v2, Vec v3);
PetscErrorCode MatMultTranspose(Mat mat, Vec x, Vec y);                /* matrix.c MatMultTranspose -> aij.c:1434 (seqaijb200 only) */
PetscErrorCode MatMultTransposeAdd(Mat mat, Vec v1, Vec v2, Vec v3);   /* aij.c:1383 */
/* COO assembly (matrix.c MatSetPreallocationCOO / MatSetValuesCOO -> aij.c:4524,4710): coo_i/coo_j/v may be host or device
   arrays (detected like PetscGetMemType); negative indices are ignored; seqaijb200 only */
PetscErrorCode MatSetPreallocationCOO(Mat A, PetscCount ncoo, PetscInt coo_i[], PetscInt coo_j[]);
PetscErrorCode MatSetValuesCOO(Mat A, const PetscScalar coo_v[], InsertMode imode);
PetscErrorCode MatGetDiagonal(Mat mat, Vec v);
PetscErrorCode MatGetDiagonalBlock(Mat A, Mat *a);                                                /* mpiaij.c:2758 */
PetscErrorCode MatDestroy(Mat *A);
typedef struct { /* include/petscmat.h MatInfo */
  double block_size, nz_allocated, nz_used, nz_unneeded, memory, assemblies, mallocs, fill_ratio_given, fill_ratio_needed, factor_mallocs;
} MatInfo;
typedef enum { MAT_LOCAL = 1, MAT_GLOBAL_MAX = 2, MAT_GLOBAL_SUM = 3 } MatInfoType;
PetscErrorCode MatGetInfo(Mat mat, MatInfoType flag, MatInfo *info);                              /* matrix.c MatGetInfo (nz_used, memory) */
/* MPIAIJ internals exposed for parity tests: diag/off-diag blocks and garray (mpiaij.h:41-76, mmaij.c:8-126) */
PetscErrorCode MatMPIAIJGetSeqAIJ(Mat A, Mat *Ad, Mat *Ao, const PetscInt *colmap[]);
PetscErrorCode MatSeqAIJGetCSRHost(Mat A, PetscInt *m, const PetscInt **i, const PetscInt **j, const PetscScalar **a); /* host copy (downloads) */
PetscErrorCode MatB200SetSpMVLayout(Mat A, PetscInt lanes_per_row, PetscInt rows_per_tile, PetscInt stages, PetscInt ctas_per_sm); /* -mat_b200_spmv_* */
PetscErrorCode MatB200SetSpMVOrdered(Mat A, PetscBool ordered); /* -mat_b200_spmv_ordered: reference-order row sums for every layout (bit-exact MatMult) */

/* ---- PC (src/ksp/pc/interface/precon.c) ---- */
PetscErrorCode PCCreate(MPI_Comm comm, PC *newpc);
PetscErrorCode PCSetType(PC pc, PCType type);
PetscErrorCode PCGetType(PC pc, PCType *type);
PetscErrorCode PCSetFromOptions(PC pc);                                                           /* -pc_type, -sub_pc_type */
PetscErrorCode PCSetOperators(PC pc, Mat Amat, Mat Pmat);
PetscErrorCode PCSetUp(PC pc);
PetscErrorCode PCApply(PC pc, Vec x, Vec y);                                                      /* precon.c:523 */
PetscErrorCode PCApplyBAorAB(PC pc, int side, Vec x, Vec y, Vec work);                            /* precon.c:810 (left only) */
PetscErrorCode PCDestroy(PC *pc);

/* ---- KSP (src/ksp/ksp/interface/{itcreate.c,itfunc.c,iterativ.c}) ---- */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *inksp);
PetscErrorCode KSPSetType(KSP ksp, KSPType type);
PetscErrorCode KSPGetType(KSP ksp, KSPType *type);
PetscErrorCode KSPSetOperators(KSP ksp, Mat Amat, Mat Pmat);
PetscErrorCode KSPGetPC(KSP ksp, PC *pc);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetInitialGuessNonzero(KSP ksp, PetscBool flg);
PetscErrorCode KSPSetFromOptions(KSP ksp);   /* -ksp_type -ksp_rtol -ksp_atol -ksp_divtol -ksp_max_it -ksp_gmres_restart -ksp_gmres_cgs_refinement_type -ksp_monitor -pc_type ... */
PetscErrorCode KSPSetUp(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);                                                   /* itfunc.c:1106 */
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm);
PetscErrorCode KSPGetConvergedReason(KSP ksp, KSPConvergedReason *reason);
PetscErrorCode KSPSetResidualHistory(KSP ksp, PetscReal a[], PetscInt na, PetscBool reset);
PetscErrorCode KSPGetResidualHistory(KSP ksp, const PetscReal *a[], PetscInt *na);
PetscErrorCode KSPMonitorSet(KSP ksp, PetscErrorCode (*monitor)(KSP, PetscInt, PetscReal, void *), void *ctx, PetscErrorCode (*monitordestroy)(void **));
PetscErrorCode KSPGMRESSetRestart(KSP ksp, PetscInt restart);
PetscErrorCode KSPGMRESSetCGSRefinementType(KSP ksp, KSPGMRESCGSRefinementType type);
PetscErrorCode KSPDestroy(KSP *ksp);

/* ---- ICC(0): host-side symbolic phase of the planned device factorisation (csrc/host/iccsym.c) ----------
   Index work only, caller-allocated outputs; the numeric kernels that consume these schedules are the next round's. */
/* MatICCFactorSymbolic_SeqAIJ levels 0, natural ordering (aijfact.c:2078-2094): ui[n+1], uj[<= ai[n]], udiag[n] */
PetscErrorCode PetscB200ICC0Symbolic(PetscInt n, const PetscInt *ai, const PetscInt *aj, PetscInt *ui, PetscInt *uj, PetscInt *udiag);
/* merge order of MatCholeskyFactorNumeric_SeqAIJ (aijfact.c:1750-1800) + dependency levels: mptr[n+1], mrow/mpos[ui[n]-n], level[n] */
PetscErrorCode PetscB200ICC0MergeSchedule(PetscInt n, const PetscInt *ui, const PetscInt *uj, PetscInt *mptr, PetscInt *mrow, PetscInt *mpos, PetscInt *level, PetscInt *nlevels);
/* column view for the scatter-free forward sweep of MatSolve_SeqSBAIJ_1_NaturalOrdering: tptr[n+1], trow/tpos[ui[n]-n] */
PetscErrorCode PetscB200ICC0ColumnView(PetscInt n, const PetscInt *ui, const PetscInt *uj, PetscInt *tptr, PetscInt *trow, PetscInt *tpos);

#ifdef __cplusplus
}
#endif
#endif
