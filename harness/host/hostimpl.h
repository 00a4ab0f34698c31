/* Private structures of the host mirror (the analogue of petsc/private/{vecimpl,matimpl,pcimpl,kspimpl}.h). */
#ifndef B200_HOSTIMPL_H
#define B200_HOSTIMPL_H
#ifndef _POSIX_C_SOURCE
#define _POSIX_C_SOURCE 200809L
#endif
#include "petscb200.h"
#include "petscb200_host.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- error handling: PetscCall/SETERRQ with a traceback buffer (src/sys/error/err.c) ---- */
#define PETSC_ERR_MEM            55
#define PETSC_ERR_SUP            56
#define PETSC_ERR_ORDER          58
#define PETSC_ERR_ARG_SIZ        60
#define PETSC_ERR_ARG_IDN        61
#define PETSC_ERR_ARG_WRONG      62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_MAT_LU_ZRPVT   71
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_ARG_INCOMP     75
#define PETSC_ERR_LIB            76
#define PETSC_ERR_PLIB           77
#define PETSC_ERR_ARG_NULL       85
#define PETSC_ERR_ARG_UNKNOWN_TYPE 86
#define PETSC_ERR_NOT_CONVERGED  91
#define PETSC_ERR_GPU            97

PetscErrorCode PetscB200Error(const char *file, int line, const char *func, PetscErrorCode code, int initial, const char *fmt, ...);
#define SETERRQ(comm, code, ...) return PetscB200Error(__FILE__, __LINE__, __func__, code, 1, __VA_ARGS__)
#define PetscCheck(cond, comm, code, ...) \
  do { \
    if (!(cond)) SETERRQ(comm, code, __VA_ARGS__); \
  } while (0)
#define PetscCall(...) \
  do { \
    PetscErrorCode ierr_ = (__VA_ARGS__); \
    if (ierr_) return PetscB200Error(__FILE__, __LINE__, __func__, ierr_, 0, " "); \
  } while (0)
/* calls into the kernel library: its error text is attached (the PetscCallCUDA analogue, petscdevice_cuda.h:59) */
#define PetscCallB200(...) \
  do { \
    int ierr_ = (__VA_ARGS__); \
    if (ierr_) return PetscB200Error(__FILE__, __LINE__, __func__, ierr_, 1, "%s", b200GetLastErrorString()); \
  } while (0)
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define PetscValidPointer(p, n) PetscCheck((p) != NULL, 0, PETSC_ERR_ARG_NULL, "Null pointer: argument # %d", n)
#define PetscValidHeader(o, n)  PetscCheck((o) != NULL, 0, PETSC_ERR_ARG_NULL, "Null object: argument # %d", n)

/* ---- global state ---- */
typedef struct {
  int        initialized;
  b200Handle h;
  int        device;
  int        rank, size; /* WORLD */
} PetscB200Globals;
extern PetscB200Globals PetscB200;
PetscErrorCode PetscB200EnsureInit(void);
static inline int PetscB200CommSize(MPI_Comm c) { return c == PETSC_COMM_SELF ? 1 : PetscB200.size; }
static inline int PetscB200CommRank(MPI_Comm c) { return c == PETSC_COMM_SELF ? 0 : PetscB200.rank; }
/* sum/max all-reduce of a few host scalars over WORLD through NCCL (the MPIU_Allreduce analogue) */
PetscErrorCode PetscB200AllreduceHost(MPI_Comm comm, double *vals, int n, int op /*0 sum 1 max*/);
PetscErrorCode PetscB200AllgatherInt64(MPI_Comm comm, int64_t mine, int64_t *all);
PetscErrorCode PetscB200AlltoallvInt(MPI_Comm comm, const int *sendcounts, const int *sendbuf, int *recvcounts, int **recvbuf);
/* PetscSplitOwnership (src/sys/utils/psplit.c:88) */
PetscErrorCode PetscSplitOwnership(MPI_Comm comm, PetscInt *n, PetscInt *N);

/* ---- options ---- */
const char *PetscB200OptionsFind(const char *pre, const char *name);

/* ---- function lists (PetscFunctionList, src/sys/dll/reg.c) ---- */
typedef struct _FL {
  char       *name;
  void       *fn;
  struct _FL *next;
} *PetscFunctionList;
PetscErrorCode PetscFunctionListAdd(PetscFunctionList *fl, const char name[], void *fn);
PetscErrorCode PetscFunctionListFind(PetscFunctionList fl, const char name[], void **fn);

/* ---- object header (PETSCHEADER, include/petsc/private/petscimpl.h) ---- */
typedef struct {
  MPI_Comm comm;
  char     type_name[64];
  char     prefix[64];
  int      refct;
  int64_t  state; /* PetscObjectState: bumped on every write access */
} PetscObjectHeader;

/* ---- Vec ---- */
typedef enum { OFFLOAD_UNALLOCATED = 0, OFFLOAD_CPU = 1, OFFLOAD_GPU = 2, OFFLOAD_BOTH = 3 } PetscOffloadMask; /* petscdevicetypes.h:240 */
struct _VecOps { /* include/petsc/private/vecimpl.h:18-110 (subset) */
  PetscErrorCode (*duplicate)(Vec, Vec *);
  PetscErrorCode (*duplicatevecs)(Vec, PetscInt, Vec **);
  PetscErrorCode (*destroy)(Vec);
  PetscErrorCode (*dot)(Vec, Vec, PetscScalar *);
  PetscErrorCode (*mdot)(Vec, PetscInt, const Vec[], PetscScalar *);
  PetscErrorCode (*norm)(Vec, NormType, PetscReal *);
  PetscErrorCode (*scale)(Vec, PetscScalar);
  PetscErrorCode (*copy)(Vec, Vec);
  PetscErrorCode (*set)(Vec, PetscScalar);
  PetscErrorCode (*axpy)(Vec, PetscScalar, Vec);
  PetscErrorCode (*aypx)(Vec, PetscScalar, Vec);
  PetscErrorCode (*axpby)(Vec, PetscScalar, PetscScalar, Vec);
  PetscErrorCode (*waxpy)(Vec, PetscScalar, Vec, Vec);
  PetscErrorCode (*maxpy)(Vec, PetscInt, const PetscScalar *, Vec *);
  PetscErrorCode (*pointwisemult)(Vec, Vec, Vec);
  PetscErrorCode (*pointwisedivide)(Vec, Vec, Vec);
  PetscErrorCode (*reciprocal)(Vec);
  PetscErrorCode (*shift)(Vec, PetscScalar);
  PetscErrorCode (*sum)(Vec, PetscScalar *);
  PetscErrorCode (*max)(Vec, PetscInt *, PetscReal *);
  PetscErrorCode (*min)(Vec, PetscInt *, PetscReal *);
  PetscErrorCode (*dot_local)(Vec, Vec, PetscScalar *);
  PetscErrorCode (*mdot_local)(Vec, PetscInt, const Vec[], PetscScalar *);
  PetscErrorCode (*norm_local)(Vec, NormType, PetscReal *);
};
struct _p_Vec {
  PetscObjectHeader hdr;
  struct _VecOps    ops;
  PetscInt          n, N, rstart, rend; /* PetscLayout */
  int               sizes_set, type_set;
  /* data (Vec_Seq / Vec_MPI + device mirror) */
  double          *d_array;     /* device array (owned unless slab/alias) */
  double          *h_array;     /* pinned host mirror, lazily allocated -- or the user's array (VecCreateSeqWithArray / VecPlaceArray) */
  int              h_user;      /* h_array belongs to the user: never freed here */
  double          *h_saved;     /* our own mirror while a user array is placed (VecResetArray restores it) */
  int              h_saved_user;
  PetscOffloadMask offloadmask;
  int              owns_device; /* 0: part of a VecDuplicateVecs slab or an alias (VecGetLocalVector) */
  void            *slab;        /* slab base pointer shared by the group (freed by the first vector) */
  int             *slab_ref;
  int              array_gotten; /* 0 none, 1 read, 2 write host access outstanding */
  /* norm cache (rvector.c:211,232): valid while state == norm_state */
  int64_t   norm_state[4];
  PetscReal norm_val[4];
  /* fused MAXPY+norm: local sum of squares left on the device by the last VecMAXPY */
  double   *d_sumsq;
  int64_t   sumsq_state;
  Vec       localrep_owner; /* for VecGetLocalVector aliases */
};
PetscErrorCode VecCreate_SeqB200(Vec v);
PetscErrorCode VecCreate_MPIB200(Vec v);
PetscErrorCode VecCreate_B200(Vec v);
/* device access for ops (syncs the mirror, sets the offload mask) */
PetscErrorCode VecB200GetArrayRead(Vec v, const double **d);
PetscErrorCode VecB200GetArrayWrite(Vec v, double **d); /* contents will be overwritten completely */
PetscErrorCode VecB200GetArray(Vec v, double **d);      /* read-modify-write */
static inline void VecStateIncrease(Vec v) { v->hdr.state++; }

/* ---- Mat ---- */
struct _MatOps { /* include/petsc/private/matimpl.h (subset) */
  PetscErrorCode (*mult)(Mat, Vec, Vec);
  PetscErrorCode (*multadd)(Mat, Vec, Vec, Vec);
  PetscErrorCode (*getdiagonal)(Mat, Vec);
  PetscErrorCode (*assemblyend)(Mat, MatAssemblyType);
  PetscErrorCode (*destroy)(Mat);
  PetscErrorCode (*getdiagonalblock)(Mat, Mat *);
  PetscErrorCode (*setvalues)(Mat, PetscInt, const PetscInt[], PetscInt, const PetscInt[], const PetscScalar[], InsertMode);
  PetscErrorCode (*setcsr)(Mat, const PetscInt *, const PetscInt *, const PetscScalar *, int on_device);
  /* fused MatMult + PCApply_Jacobi: w = dinv .* (A x) (the ops->applyBA hook of precon.c:810-865 lands here) */
  PetscErrorCode (*multjacobi)(Mat, Vec x, Vec dinv, Vec w);
  PetscErrorCode (*multtranspose)(Mat, Vec, Vec);
  PetscErrorCode (*multtransposeadd)(Mat, Vec, Vec, Vec);
  PetscErrorCode (*setpreallocationcoo)(Mat, PetscCount, const PetscInt[], const PetscInt[]);
  PetscErrorCode (*setvaluescoo)(Mat, const PetscScalar[], InsertMode);
};
typedef struct { /* Mat_SeqAIJ (aij.h:47-92) device mirror */
  PetscInt    m, n;
  int64_t     nz;
  int        *d_i, *d_j;
  double     *d_a;
  b200CsrPlan plan;
  /* compressed-row view (aij.h compressedrow): rows with at least one entry */
  int         cr_use, cr_nrows;
  int        *d_cr_i, *d_cr_rindex;
  /* host copies kept only on request */
  int    *h_i, *h_j;
  double *h_a;
  PetscInt nonzerorowcnt;
  /* explicit transpose for MatMultTranspose (built on first use; values re-gathered when the matrix state moved on) */
  b200CsrTranspose T;
  int64_t          T_state;
  /* COO assembly plan (MatSetPreallocationCOO) */
  b200CooPlan coo;
} Mat_SeqAIJB200;
typedef struct { /* Mat_MPIAIJ (mpiaij.h:41-76) */
  Mat       A, B;     /* diag and off-diag blocks (seqaijb200) */
  PetscInt *garray;   /* global column of each compacted off-diag column, sorted */
  PetscInt  ec;
  Vec       lvec;     /* halo vector, length ec */
  b200Halo  Mvctx;    /* the scatter */
  int64_t  *ranges;   /* column ownership ranges [size+1] */
} Mat_MPIAIJB200;
typedef struct _COOEntry {
  PetscInt row, col;
  double   v;
  size_t   seq; /* arrival order: makes the (row,col) sort stable so that INSERT_VALUES keeps the LAST value */
} COOEntry;
struct _p_Mat {
  PetscObjectHeader hdr;
  struct _MatOps    ops;
  PetscInt          m, n, M, N, rstart, rend, cstart, cend;
  int               sizes_set, type_set, assembled;
  void             *data;
  char              defaultvectype[32];
  /* MatSetValues staging (the stash of matstash.c, local rows only) */
  COOEntry *coo;
  size_t    ncoo, coocap;
  int       stash_mode;    /* 0 = nothing staged, else the InsertMode of the staged entries (mixing is an error, matrix.c:1480) */
  int       ever_assembled;
  int       spmv_layout[4];
  int       spmv_ordered; /* -mat_b200_spmv_ordered: row sums in the reference's order for every lane count (bit-exact MatMult) */
  int64_t   coo_n; /* length of the arrays given to MatSetPreallocationCOO */
};
PetscErrorCode MatCreate_SeqAIJB200(Mat A);
PetscErrorCode PetscB200MPIAIJSplit(PetscInt m, PetscInt cstart, PetscInt cend, const PetscInt *ai, const PetscInt *aj, const PetscScalar *aa, PetscInt **Ai, PetscInt **Aj, PetscScalar **Aa, PetscInt **Bi, PetscInt **Bj, PetscScalar **Ba, PetscInt **garray, PetscInt *ec);
PetscErrorCode PetscB200HaloPlanRecv(PetscInt ec, const PetscInt *garray, int size, const int64_t *ranges, int *recv_counts, int *recv_offsets);
PetscErrorCode MatCreate_MPIAIJB200(Mat A);

/* ---- PC ---- */
struct _PCOps { /* include/petsc/private/pcimpl.h:11-31 (subset) */
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*applyBA)(PC, int, Vec, Vec, Vec);
  PetscErrorCode (*setfromoptions)(PC);
  PetscErrorCode (*destroy)(PC);
  PetscErrorCode (*reset)(PC);
};
struct _p_PC {
  PetscObjectHeader hdr;
  struct _PCOps     ops;
  Mat               mat, pmat;
  int               setupcalled, type_set;
  int64_t           matstate; /* pmat->hdr.state at the last set-up: PCSetUp refreshes when the operator changed (precon.c:1080-1110) */
  void             *data;
};

/* ---- KSP ---- */
struct _KSPOps {
  PetscErrorCode (*setup)(KSP);
  PetscErrorCode (*solve)(KSP);
  PetscErrorCode (*setfromoptions)(KSP);
  PetscErrorCode (*destroy)(KSP);
  PetscErrorCode (*reset)(KSP);
};
struct _p_KSP {
  PetscObjectHeader hdr;
  struct _KSPOps    ops;
  PC                pc;
  Vec               vec_rhs, vec_sol;
  PetscReal         rtol, abstol, divtol, ttol, rnorm0, rnorm;
  PetscInt          max_it, its;
  int               guess_zero, setupcalled, type_set, monitor_stdout;
  KSPConvergedReason reason;
  PetscReal        *res_hist;
  PetscInt          res_hist_len, res_hist_max;
  int               res_hist_reset, res_hist_alloc;
  PetscErrorCode (*monitor)(KSP, PetscInt, PetscReal, void *);
  void *monitorctx;
  Vec  *work;
  PetscInt nwork;
  void *data;
};
PetscErrorCode KSPLogResidualHistory(KSP ksp, PetscReal norm);
PetscErrorCode KSPMonitor(KSP ksp, PetscInt it, PetscReal rnorm);
PetscErrorCode KSPConvergedDefault(KSP ksp, PetscInt n, PetscReal rnorm, KSPConvergedReason *reason);
PetscErrorCode KSPInitialResidual(KSP ksp, Vec vsoln, Vec vt1, Vec vt2, Vec vres, Vec vb);
PetscErrorCode KSPCreate_GMRES(KSP ksp);
PetscErrorCode KSPCreate_CG(KSP ksp);
PetscErrorCode KSPCreate_PIPECG(KSP ksp);
PetscErrorCode KSPCreate_PREONLY(KSP ksp);
PetscErrorCode PCCreate_None(PC pc);
PetscErrorCode PCCreate_Jacobi(PC pc);
PetscErrorCode PCCreate_BJacobi(PC pc);
PetscErrorCode PCCreate_ILU(PC pc);

#endif
