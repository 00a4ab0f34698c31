/*
 * petscb200_plugin.c -- the PETSc-side binding of libpetscb200.so: new types registered with PETSc's own registries so
 * that an UNMODIFIED PETSc program runs the Krylov hot path on the sm_100a kernels with
 *
 *     ./ex2 -dll_append /path/libpetscb200plugin.so -mat_type aijb200 -vec_type b200 -ksp_type gmres -pc_type jacobi
 *
 * Registered (PetscDLLibraryRegister_petscb200plugin, loaded by src/sys/dll/reg.c:79,150):
 *   VecRegister            "seqb200", "b200"                         (src/vec/vec/interface/vecreg.c:252)
 *   MatRegisterRootName    "aijb200" -> "seqaijb200" / "mpiaijb200"  (src/mat/interface/matreg.c:328)
 *   MatRegister            "seqaijb200"                              (matreg.c:293)
 *   MatSolverTypeRegister  "b200" for seqaijb200, MAT_FACTOR_ILU     (src/mat/interface/matrix.c:4720)
 *   PCRegister             "jacobib200": PCJACOBI with a fused ops->applyBA (src/ksp/pc/interface/pcregis.c, precon.c:810-865)
 *
 * Structure mirrors the reference's own device subclassing (aijcusparse.cu:2807-2868, veccupmimpl.h:994-1047): create the
 * parent (MATSEQAIJ / VECSEQ), keep its host data structures, overwrite the ops of the hot path with functions that run
 * on a device mirror, and keep host and device coherent with an offload mask.  Everything the plugin does not override
 * falls back to the parent's host implementation, which reaches the data through VecGetArray*() -> our hooks.
 *
 * Sequential types only: the PETSc this was built against (MPIUNI) has a single rank; the row-partitioned mpiaijb200 /
 * mpib200 types live in the stand-alone host mirror (petsc_b200/csrc/host) where NCCL replaces MPI.
 * Compile: see petsc_plugin/Makefile (needs PETSC_DIR/PETSC_ARCH of the PETSc the application links).
 */
#include <petsc/private/vecimpl.h>
#include <petsc/private/matimpl.h>
#include <petsc/private/pcimpl.h>
#include <../src/vec/vec/impls/dvecimpl.h>
#include <../src/mat/impls/aij/seq/aij.h>
#include <petscksp.h>
#include "petscb200.h"

#define VECSEQB200    "seqb200"
#define VECB200       "b200"
#define MATSEQAIJB200 "seqaijb200"
#define MATAIJB200    "aijb200"
#define MATSOLVERB200 "b200"
#define PCJACOBIB200  "jacobib200"

static b200Handle PB_h = NULL;

/* Mat::boundtocpu exists only in a PETSc configured with a device back end (include/petsc/private/matimpl.h:493-497) */
#if PetscDefined(HAVE_DEVICE)
  #define PB_BoundToCPU(A) ((A)->boundtocpu)
#else
  #define PB_BoundToCPU(A) PETSC_FALSE
#endif

#define PetscCallB200(...) \
  do { \
    int b200_ierr_ = (__VA_ARGS__); \
    PetscCheck(!b200_ierr_, PETSC_COMM_SELF, (PetscErrorCode)b200_ierr_, "%s", b200GetLastErrorString()); \
  } while (0)

static PetscErrorCode PB_Init(void)
{
  PetscFunctionBegin;
  if (!PB_h) {
    PetscCallB200(b200Create(&PB_h, -1));
#if defined(PETSC_HAVE_CUDA)
    PetscCallB200(b200SetStream(PB_h, (void *)PetscDefaultCudaStream)); /* include/petscdevice_cuda.h:180 */
#endif
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== Vec: seqb200 */
enum { PB_UNALLOCATED = 0, PB_CPU = 1, PB_GPU = 2, PB_BOTH = 3 }; /* PetscOffloadMask, include/petscdevicetypes.h:240 */

typedef struct {
  Vec_Seq seq;    /* MUST be first: the parent's data, untouched (host array, VECHEADER) */
  double *d;      /* device mirror */
  int     mask;
  double *d_sumsq; /* fused MAXPY+norm */
  PetscObjectState sumsq_state;
} Vec_SeqB200;

static struct _VecOps PB_VecSeqOps; /* the parent's ops, captured at first creation */
static PetscBool      PB_VecSeqOpsSet = PETSC_FALSE;

static PetscErrorCode VecGetArray_SeqB200(Vec, PetscScalar **);
#define PB_IsB200(v) ((v)->ops->getarray == VecGetArray_SeqB200)

static PetscErrorCode PB_VecAlloc(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  if (!b->d && v->map->n) PetscCallB200(b200Malloc(PB_h, (void **)&b->d, sizeof(double) * (size_t)v->map->n));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecToDevice(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecAlloc(v));
  if (b->mask == PB_CPU || b->mask == PB_UNALLOCATED) {
    if (v->map->n) PetscCallB200(b200MemcpyHtoDAsync(PB_h, b->d, b->seq.array, sizeof(double) * (size_t)v->map->n));
    b->mask = PB_BOTH;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecToHost(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  if (b->mask == PB_GPU) {
    if (v->map->n) PetscCallB200(b200MemcpyDtoH(PB_h, b->seq.array, b->d, sizeof(double) * (size_t)v->map->n));
    b->mask = PB_BOTH;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecRead(Vec v, const double **p)
{
  PetscFunctionBegin;
  PetscCall(PB_VecToDevice(v));
  *p = ((Vec_SeqB200 *)v->data)->d;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecRW(Vec v, double **p)
{
  PetscFunctionBegin;
  PetscCall(PB_VecToDevice(v));
  ((Vec_SeqB200 *)v->data)->mask = PB_GPU;
  *p = ((Vec_SeqB200 *)v->data)->d;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecWrite(Vec v, double **p)
{
  PetscFunctionBegin;
  PetscCall(PB_VecAlloc(v));
  ((Vec_SeqB200 *)v->data)->mask = PB_GPU;
  *p = ((Vec_SeqB200 *)v->data)->d;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* host access hooks (rvector.c:2070-2160 dispatch to these because ops->getarray is set) */
static PetscErrorCode VecGetArray_SeqB200(Vec v, PetscScalar **a)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecToHost(v));
  b->mask = PB_CPU;
  *a      = b->seq.array;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecGetArrayRead_SeqB200(Vec v, const PetscScalar **a)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecToHost(v));
  if (b->mask == PB_UNALLOCATED) b->mask = PB_CPU;
  *a = b->seq.array;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecGetArrayWrite_SeqB200(Vec v, PetscScalar **a)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  b->mask = PB_CPU;
  *a      = b->seq.array;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecRestoreArray_SeqB200(Vec v, PetscScalar **a)
{
  PetscFunctionBegin;
  (void)v;
  if (a) *a = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecRestoreArrayRead_SeqB200(Vec v, const PetscScalar **a)
{
  PetscFunctionBegin;
  (void)v;
  if (a) *a = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* device pointer hand-out: how VecScatter/PetscSF learn the data is on the device (rvector.c:2365-2381) */
static PetscErrorCode VecGetArrayAndMemType_SeqB200(Vec v, PetscScalar **a, PetscMemType *m)
{
  PetscFunctionBegin;
  PetscCall(PB_VecRW(v, a));
  if (m) *m = PETSC_MEMTYPE_CUDA;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecGetArrayReadAndMemType_SeqB200(Vec v, const PetscScalar **a, PetscMemType *m)
{
  PetscFunctionBegin;
  PetscCall(PB_VecRead(v, a));
  if (m) *m = PETSC_MEMTYPE_CUDA;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecGetArrayWriteAndMemType_SeqB200(Vec v, PetscScalar **a, PetscMemType *m)
{
  PetscFunctionBegin;
  PetscCall(PB_VecWrite(v, a));
  if (m) *m = PETSC_MEMTYPE_CUDA;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecRestoreArrayAndMemType_SeqB200(Vec v, PetscScalar **a)
{
  PetscFunctionBegin;
  (void)v;
  if (a) *a = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecRestoreArrayReadAndMemType_SeqB200(Vec v, const PetscScalar **a)
{
  PetscFunctionBegin;
  (void)v;
  if (a) *a = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- the BLAS-1 ops of the Krylov loop (the ops table of bvec2.c:694-790, device versions) ---- */
#define N_(v) ((int64_t)(v)->map->n)

static PetscErrorCode VecSet_SeqB200(Vec x, PetscScalar a)
{
  double *d;
  PetscFunctionBegin;
  PetscCall(PB_VecWrite(x, &d));
  PetscCallB200(b200VecSet(PB_h, N_(x), a, d));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecCopy_SeqB200(Vec x, Vec y)
{
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (!PB_IsB200(y)) { /* device -> plain host vector */
    PetscScalar       *ya;
    const PetscScalar *xa;
    PetscCall(VecGetArrayRead(x, &xa));
    PetscCall(VecGetArrayWrite(y, &ya));
    PetscCall(PetscArraycpy(ya, xa, x->map->n));
    PetscCall(VecRestoreArrayWrite(y, &ya));
    PetscCall(VecRestoreArrayRead(x, &xa));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(y, &dy));
  PetscCallB200(b200VecCopy(PB_h, N_(x), dx, dy));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecScale_SeqB200(Vec x, PetscScalar a)
{
  double *d;
  PetscFunctionBegin;
  PetscCall(PB_VecRW(x, &d));
  PetscCallB200(b200VecScale(PB_h, N_(x), a, d));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecAXPY_SeqB200(Vec y, PetscScalar a, Vec x)
{
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (!PB_IsB200(x)) PetscFunctionReturn((*PB_VecSeqOps.axpy)(y, a, x));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRW(y, &dy));
  PetscCallB200(b200VecAXPY(PB_h, N_(y), a, dx, dy));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecAYPX_SeqB200(Vec y, PetscScalar b, Vec x)
{
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (!PB_IsB200(x)) PetscFunctionReturn((*PB_VecSeqOps.aypx)(y, b, x));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRW(y, &dy));
  PetscCallB200(b200VecAYPX(PB_h, N_(y), b, dx, dy));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecAXPBY_SeqB200(Vec y, PetscScalar a, PetscScalar b, Vec x)
{
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (!PB_IsB200(x)) PetscFunctionReturn((*PB_VecSeqOps.axpby)(y, a, b, x));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRW(y, &dy));
  PetscCallB200(b200VecAXPBY(PB_h, N_(y), a, b, dx, dy));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecWAXPY_SeqB200(Vec w, PetscScalar a, Vec x, Vec y)
{
  const double *dx, *dy;
  double       *dw;
  PetscFunctionBegin;
  if (!PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*PB_VecSeqOps.waxpy)(w, a, x, y));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(y, &dy));
  PetscCall(PB_VecWrite(w, &dw));
  PetscCallB200(b200VecWAXPY(PB_h, N_(w), a, dx, dy, dw));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecPointwiseMult_SeqB200(Vec w, Vec x, Vec y)
{
  const double *dx, *dy;
  double       *dw;
  PetscFunctionBegin;
  if (!PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*PB_VecSeqOps.pointwisemult)(w, x, y));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(y, &dy));
  if (w == x || w == y) PetscCall(PB_VecRW(w, &dw));
  else PetscCall(PB_VecWrite(w, &dw));
  PetscCallB200(b200VecPointwiseMult(PB_h, N_(w), dx, dy, dw));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecPointwiseDivide_SeqB200(Vec w, Vec x, Vec y)
{
  const double *dx, *dy;
  double       *dw;
  PetscFunctionBegin;
  if (!PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*PB_VecSeqOps.pointwisedivide)(w, x, y));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(y, &dy));
  if (w == x || w == y) PetscCall(PB_VecRW(w, &dw));
  else PetscCall(PB_VecWrite(w, &dw));
  PetscCallB200(b200VecPointwiseDivide(PB_h, N_(w), dx, dy, dw));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecReciprocal_SeqB200(Vec x)
{
  double *d;
  PetscFunctionBegin;
  PetscCall(PB_VecRW(x, &d));
  PetscCallB200(b200VecReciprocal(PB_h, N_(x), d));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecDot_SeqB200(Vec x, Vec y, PetscScalar *z)
{
  const double *dx, *dy;
  PetscFunctionBegin;
  if (!PB_IsB200(y)) PetscFunctionReturn((*PB_VecSeqOps.dot)(x, y, z));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(y, &dy));
  PetscCallB200(b200VecDot(PB_h, N_(x), dx, dy, z));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMDot_SeqB200(Vec x, PetscInt nv, const Vec y[], PetscScalar *z)
{
  const double  *dx;
  const double **yp;
  PetscFunctionBegin;
  for (PetscInt j = 0; j < nv; j++)
    if (!PB_IsB200(y[j])) PetscFunctionReturn((*PB_VecSeqOps.mdot)(x, nv, y, z));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PetscMalloc1(nv, &yp));
  for (PetscInt j = 0; j < nv; j++) PetscCall(PB_VecRead(y[j], &yp[j]));
  PetscCallB200(b200VecMDot(PB_h, N_(x), (int)nv, dx, yp, z)); /* one kernel for all nv: x read once */
  PetscCall(PetscFree(yp));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMAXPY_SeqB200(Vec x, PetscInt nv, const PetscScalar *alpha, Vec *y)
{
  Vec_SeqB200   *b = (Vec_SeqB200 *)x->data;
  double        *dx;
  const double **yp;
  PetscFunctionBegin;
  for (PetscInt j = 0; j < nv; j++)
    if (!PB_IsB200(y[j])) PetscFunctionReturn((*PB_VecSeqOps.maxpy)(x, nv, alpha, y));
  PetscCall(PetscMalloc1(nv, &yp));
  for (PetscInt j = 0; j < nv; j++) PetscCall(PB_VecRead(y[j], &yp[j]));
  PetscCall(PB_VecRW(x, &dx));
  if (!b->d_sumsq) PetscCallB200(b200Malloc(PB_h, (void **)&b->d_sumsq, sizeof(double)));
  /* one pass, reference association (bit-identical to VecMAXPY_Seq) + ||x||^2 of the result for the VecNorm that
     KSPGMRESCycle issues next; the interface bumps the object state right after this returns (rvector.c:1385) */
  PetscCallB200(b200VecMAXPYAsync(PB_h, N_(x), (int)nv, alpha, yp, dx, b->d_sumsq));
  PetscCall(PetscObjectStateGet((PetscObject)x, &b->sumsq_state));
  b->sumsq_state += 1;
  PetscCall(PetscFree(yp));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecNorm_SeqB200(Vec x, NormType type, PetscReal *z)
{
  Vec_SeqB200     *b = (Vec_SeqB200 *)x->data;
  const double    *dx;
  PetscObjectState st;
  PetscFunctionBegin;
  PetscCall(PetscObjectStateGet((PetscObject)x, &st));
  if ((type == NORM_2 || type == NORM_FROBENIUS) && b->d_sumsq && b->sumsq_state == st) {
    double ss;
    PetscCallB200(b200MemcpyDtoH(PB_h, &ss, b->d_sumsq, sizeof(double)));
    *z = PetscSqrtReal(ss);
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_VecRead(x, &dx));
  if (type == NORM_1_AND_2) {
    PetscCallB200(b200VecNorm(PB_h, N_(x), dx, 0, &z[0]));
    PetscCallB200(b200VecNorm(PB_h, N_(x), dx, 1, &z[1]));
  } else PetscCallB200(b200VecNorm(PB_h, N_(x), dx, type == NORM_1 ? 0 : (type == NORM_INFINITY ? 3 : 1), z));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecDestroy_SeqB200(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  if (b) {
    if (b->d) PetscCallB200(b200Free(PB_h, b->d));
    if (b->d_sumsq) PetscCallB200(b200Free(PB_h, b->d_sumsq));
    b->d = b->d_sumsq = NULL;
  }
  PetscCall((*PB_VecSeqOps.destroy)(v)); /* VecDestroy_Seq frees the host array and v->data */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecResetArray_SeqB200(Vec v)
{
  PetscFunctionBegin;
  PetscCall(PB_VecToHost(v));
  PetscCall((*PB_VecSeqOps.resetarray)(v));
  ((Vec_SeqB200 *)v->data)->mask = PB_CPU;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecPlaceArray_SeqB200(Vec v, const PetscScalar *a)
{
  PetscFunctionBegin;
  PetscCall(PB_VecToHost(v));
  PetscCall((*PB_VecSeqOps.placearray)(v, a));
  ((Vec_SeqB200 *)v->data)->mask = PB_CPU;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecReplaceArray_SeqB200(Vec v, const PetscScalar *a)
{
  PetscFunctionBegin;
  PetscCall((*PB_VecSeqOps.replacearray)(v, a));
  ((Vec_SeqB200 *)v->data)->mask = PB_CPU;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode VecCreate_SeqB200(Vec v);
static PetscErrorCode       VecDuplicate_SeqB200(Vec win, Vec *V)
{
  PetscFunctionBegin;
  PetscCall(VecCreate(PetscObjectComm((PetscObject)win), V));
  PetscCall(PetscLayoutReference(win->map, &(*V)->map));
  PetscCall(VecSetType(*V, VECSEQB200));
  (*V)->stash.ignorenegidx = win->stash.ignorenegidx;
  PetscCall(PetscObjectListDuplicate(((PetscObject)win)->olist, &((PetscObject)*V)->olist));
  PetscCall(PetscFunctionListDuplicate(((PetscObject)win)->qlist, &((PetscObject)*V)->qlist));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode VecCreate_SeqB200(Vec v)
{
  PetscMPIInt  size;
  Vec_SeqB200 *b;
  Vec_Seq     *s;

  PetscFunctionBegin;
  PetscCallMPI(MPI_Comm_size(PetscObjectComm((PetscObject)v), &size));
  PetscCheck(size == 1, PetscObjectComm((PetscObject)v), PETSC_ERR_ARG_WRONG, "Cannot create VECSEQB200 on more than one process");
  PetscCall(PB_Init());
  PetscCall(VecSetType(v, VECSEQ)); /* parent: host array + Vec_Seq (VecCreate_Seq is not exported; VecSetType is) */
  if (!PB_VecSeqOpsSet) {
    PB_VecSeqOps    = *v->ops;
    PB_VecSeqOpsSet = PETSC_TRUE;
  }
  /* grow the parent's data structure in place: Vec_SeqB200 starts with a Vec_Seq */
  s = (Vec_Seq *)v->data;
  PetscCall(PetscNew(&b));
  b->seq = *s;
  PetscCall(PetscFree(s));
  v->data = (void *)b;
  b->mask = PB_UNALLOCATED;

  v->ops->duplicate                  = VecDuplicate_SeqB200;
  v->ops->destroy                    = VecDestroy_SeqB200;
  v->ops->getarray                   = VecGetArray_SeqB200;
  v->ops->restorearray               = VecRestoreArray_SeqB200;
  v->ops->getarrayread               = VecGetArrayRead_SeqB200;
  v->ops->restorearrayread           = VecRestoreArrayRead_SeqB200;
  v->ops->getarraywrite              = VecGetArrayWrite_SeqB200;
  v->ops->restorearraywrite          = VecRestoreArray_SeqB200;
  v->ops->getarrayandmemtype         = VecGetArrayAndMemType_SeqB200;
  v->ops->restorearrayandmemtype     = VecRestoreArrayAndMemType_SeqB200;
  v->ops->getarrayreadandmemtype     = VecGetArrayReadAndMemType_SeqB200;
  v->ops->restorearrayreadandmemtype = VecRestoreArrayReadAndMemType_SeqB200;
  v->ops->getarraywriteandmemtype    = VecGetArrayWriteAndMemType_SeqB200;
  v->ops->placearray                 = VecPlaceArray_SeqB200;
  v->ops->resetarray                 = VecResetArray_SeqB200;
  v->ops->replacearray               = VecReplaceArray_SeqB200;
  v->ops->set                        = VecSet_SeqB200;
  v->ops->copy                       = VecCopy_SeqB200;
  v->ops->scale                      = VecScale_SeqB200;
  v->ops->axpy                       = VecAXPY_SeqB200;
  v->ops->aypx                       = VecAYPX_SeqB200;
  v->ops->axpby                      = VecAXPBY_SeqB200;
  v->ops->waxpy                      = VecWAXPY_SeqB200;
  v->ops->maxpy                      = VecMAXPY_SeqB200;
  v->ops->pointwisemult              = VecPointwiseMult_SeqB200;
  v->ops->pointwisedivide            = VecPointwiseDivide_SeqB200;
  v->ops->reciprocal                 = VecReciprocal_SeqB200;
  v->ops->dot                        = VecDot_SeqB200;
  v->ops->tdot                       = VecDot_SeqB200;
  v->ops->mdot                       = VecMDot_SeqB200;
  v->ops->mtdot                      = VecMDot_SeqB200;
  v->ops->norm                       = VecNorm_SeqB200;
  v->ops->dot_local                  = VecDot_SeqB200;
  v->ops->tdot_local                 = VecDot_SeqB200;
  v->ops->mdot_local                 = VecMDot_SeqB200;
  v->ops->mtdot_local                = VecMDot_SeqB200;
  v->ops->norm_local                 = VecNorm_SeqB200;
  PetscCall(PetscObjectChangeTypeName((PetscObject)v, VECSEQB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== Mat: seqaijb200 */
typedef struct {
  int             *d_i, *d_j;
  double          *d_a;
  b200CsrPlan      plan;
  PetscObjectState nonzerostate, valstate;
  PetscBool        valid;
  PetscErrorCode (*destroy_seqaij)(Mat);
  PetscErrorCode (*duplicate_seqaij)(Mat, MatDuplicateOption, Mat *);
  PetscErrorCode (*mult_seqaij)(Mat, Vec, Vec);
  PetscErrorCode (*multadd_seqaij)(Mat, Vec, Vec, Vec);
  PetscErrorCode (*multtranspose_seqaij)(Mat, Vec, Vec);
  PetscErrorCode (*multtransposeadd_seqaij)(Mat, Vec, Vec, Vec);
  /* explicit transposed pattern for MatMultTranspose (built on first use) */
  b200CsrTranspose T;
  PetscObjectState T_nonzerostate, T_valstate;
  /* COO assembly: the parent's composed host functions + a device plan that adopts the parent's jmap/perm */
  PetscErrorCode (*coo_prealloc_seqaij)(Mat, PetscCount, PetscInt[], PetscInt[]);
  PetscErrorCode (*coo_setvalues_seqaij)(Mat, const PetscScalar[], InsertMode);
  b200CooPlan coo;
} Mat_B200;

static PetscErrorCode PB_MatFreeDevice(Mat_B200 *m)
{
  PetscFunctionBegin;
  if (m->plan) b200CsrPlanDestroy(m->plan);
  m->plan = NULL;
  if (m->T) b200CsrTransposeDestroy(m->T);
  m->T = NULL;
  PetscCallB200(b200Free(PB_h, m->d_i));
  PetscCallB200(b200Free(PB_h, m->d_j));
  PetscCallB200(b200Free(PB_h, m->d_a));
  m->d_i = m->d_j = NULL;
  m->d_a   = NULL;
  m->valid = PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* host -> device mirror, keyed on nonzerostate (pattern) and the object state (values): the protocol of
   MatSeqAIJCUSPARSECopyToGPU (aijcusparse.cu:1477-1590) */
/* -mat_b200_spmv_lanes <0|1|2|4|8|16|32>: lanes per row of the SpMV kernels; 1 = the parity mode that reproduces
   MatMult_SeqAIJ's left-to-right row sums bit for bit (default 0 = chosen from the row-length statistics) */
static PetscErrorCode PB_PlanSetFromOptions(Mat A, b200CsrPlan plan)
{
  PetscInt  lanes = 0;
  PetscBool set   = PETSC_FALSE;
  PetscFunctionBegin;
  PetscCall(PetscOptionsGetInt(((PetscObject)A)->options, ((PetscObject)A)->prefix, "-mat_b200_spmv_lanes", &lanes, &set));
  if (set) PetscCallB200(b200CsrPlanSetLayout(plan, (int)lanes, 0, 0, 0));
  {
    /* -mat_b200_spmv_ordered: reference-order row sums for every lane count (bit-identical MatMult at 0.6-1.0x the speed) */
    PetscBool ordered = PETSC_FALSE;
    PetscCall(PetscOptionsGetBool(((PetscObject)A)->options, ((PetscObject)A)->prefix, "-mat_b200_spmv_ordered", &ordered, NULL));
    if (ordered) PetscCallB200(b200CsrPlanSetSummation(plan, 0));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PB_MatSyncEx(Mat A, PetscBool need_assembled);
static PetscErrorCode PB_MatSync(Mat A)
{
  return PB_MatSyncEx(A, PETSC_TRUE);
}
static PetscErrorCode PB_MatSyncEx(Mat A, PetscBool need_assembled)
{
  Mat_B200        *m = (Mat_B200 *)A->spptr;
  Mat_SeqAIJ      *a = (Mat_SeqAIJ *)A->data;
  PetscObjectState st;
  const PetscInt   nr = A->rmap->n;
  const size_t     nz = (size_t)a->nz;

  PetscFunctionBegin;
  PetscCheck(A->assembled || !need_assembled, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "Not for unassembled matrix");
  PetscCall(PetscObjectStateGet((PetscObject)A, &st));
  if (!m->valid || m->nonzerostate != A->nonzerostate) {
    PetscCall(PB_MatFreeDevice(m));
    PetscCallB200(b200Malloc(PB_h, (void **)&m->d_i, sizeof(int) * ((size_t)nr + 1)));
    PetscCallB200(b200Malloc(PB_h, (void **)&m->d_j, sizeof(int) * (nz + 1)));
    PetscCallB200(b200Malloc(PB_h, (void **)&m->d_a, sizeof(double) * (nz + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_i, a->i, sizeof(int) * ((size_t)nr + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_j, a->j, sizeof(int) * nz));
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_a, a->a, sizeof(double) * nz));
    PetscCallB200(b200CsrPlanCreate(PB_h, (int)nr, (int)A->cmap->n, (int64_t)nz, m->d_i, m->d_j, &m->plan));
    PetscCall(PB_PlanSetFromOptions(A, m->plan));
    m->nonzerostate = A->nonzerostate;
    m->valstate     = st;
    m->valid        = PETSC_TRUE;
  } else if (m->valstate != st) {
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_a, a->a, sizeof(double) * nz));
    m->valstate = st;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode MatMult_SeqAIJB200(Mat A, Vec x, Vec y)
{
  Mat_B200     *m = (Mat_B200 *)A->spptr;
  Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*m->mult_seqaij)(A, x, y)); /* host vectors: parent */
  PetscCall(PB_MatSync(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(y, &dy));
  PetscCallB200(b200CsrSpMV(PB_h, m->plan, m->d_a, dx, dy));
  PetscCall(PetscLogFlops(2.0 * a->nz - a->nonzerorowcnt)); /* aij.c:1497 */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultAdd_SeqAIJB200(Mat A, Vec x, Vec y, Vec z)
{
  Mat_B200     *m = (Mat_B200 *)A->spptr;
  Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
  const double *dx, *dy;
  double       *dz;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y) || !PB_IsB200(z)) PetscFunctionReturn((*m->multadd_seqaij)(A, x, y, z));
  PetscCall(PB_MatSync(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(y, &dy));
  if (z == y) PetscCall(PB_VecRW(z, &dz));
  else PetscCall(PB_VecWrite(z, &dz));
  PetscCallB200(b200CsrSpMVAdd(PB_h, m->plan, m->d_a, dx, dy, dz));
  PetscCall(PetscLogFlops(2.0 * a->nz));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetDiagonal_SeqAIJB200(Mat A, Vec v)
{
  Mat_B200 *m = (Mat_B200 *)A->spptr;
  double   *dv;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(v)) { /* host vector: the parent's loop over a->diag */
    const PetscInt  *diag;
    Mat_SeqAIJ      *a = (Mat_SeqAIJ *)A->data;
    PetscScalar     *va;
    PetscCall(MatGetDiagonalMarkers_SeqAIJ(A, &diag, NULL));
    PetscCall(VecGetArrayWrite(v, &va));
    for (PetscInt i = 0; i < A->rmap->n; i++) va[i] = (diag[i] < a->i[i + 1] && a->j[diag[i]] == i) ? a->a[diag[i]] : 0.0;
    PetscCall(VecRestoreArrayWrite(v, &va));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_MatSync(A));
  PetscCall(PB_VecWrite(v, &dv));
  PetscCallB200(b200CsrGetDiagonal(PB_h, (int)A->rmap->n, m->d_i, m->d_j, m->d_a, dv, NULL));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatMultTranspose[Add]_SeqAIJ (aij.c:1383-1440) on the explicit transposed pattern: bit-identical accumulation order */
static PetscErrorCode PB_MatSyncTranspose(Mat A)
{
  Mat_B200 *m = (Mat_B200 *)A->spptr;
  PetscFunctionBegin;
  PetscCall(PB_MatSync(A)); /* may drop m->T when the pattern changed */
  if (!m->T) {
    PetscCallB200(b200CsrTransposeCreate(PB_h, (int)A->rmap->n, (int)A->cmap->n, (int64_t)((Mat_SeqAIJ *)A->data)->nz, m->d_i, m->d_j, &m->T));
    m->T_valstate = (PetscObjectState)-1;
    {
      b200CsrPlan tp;
      PetscCallB200(b200CsrTransposeGetPlan(m->T, &tp));
      PetscCall(PB_PlanSetFromOptions(A, tp));
    }
  }
  if (m->T_valstate != m->valstate) {
    PetscCallB200(b200CsrTransposeSetValues(PB_h, m->T, m->d_a));
    m->T_valstate = m->valstate;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultTranspose_SeqAIJB200(Mat A, Vec x, Vec y)
{
  Mat_B200     *m = (Mat_B200 *)A->spptr;
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*m->multtranspose_seqaij)(A, x, y));
  PetscCall(PB_MatSyncTranspose(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(y, &dy));
  PetscCallB200(b200CsrTransposeSpMV(PB_h, m->T, dx, NULL, dy));
  PetscCall(PetscLogFlops(2.0 * ((Mat_SeqAIJ *)A->data)->nz)); /* aij.c:1427 */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultTransposeAdd_SeqAIJB200(Mat A, Vec x, Vec z, Vec y)
{
  Mat_B200     *m = (Mat_B200 *)A->spptr;
  const double *dx, *dz;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y) || !PB_IsB200(z)) PetscFunctionReturn((*m->multtransposeadd_seqaij)(A, x, z, y));
  PetscCall(PB_MatSyncTranspose(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(z, &dz));
  if (y == z) PetscCall(PB_VecRW(y, &dy));
  else PetscCall(PB_VecWrite(y, &dy));
  PetscCallB200(b200CsrTransposeSpMV(PB_h, m->T, dx, dz, dy));
  PetscCall(PetscLogFlops(2.0 * ((Mat_SeqAIJ *)A->data)->nz));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatBindToCPU (matrix.c sets PB_BoundToCPU(A) before calling this): the host CSR is the master copy, so binding is a flag that
   every device op checks (cf. MatBindToCPU_SeqAIJCUSPARSE, aijcusparse.cu:2741-2805, which swaps the ops tables) */
static PetscErrorCode MatBindToCPU_SeqAIJB200(Mat A, PetscBool flg)
{
  PetscFunctionBegin;
#if PetscDefined(HAVE_DEVICE)
  A->boundtocpu = flg;
#else
  (void)A;
  (void)flg; /* a PETSc configured without a device has no binding state: MatBindToCPU() is a no-op there */
#endif
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* where MatMult runs (MatGetCurrentMemType_SeqAIJCUSPARSE analogue) */
static PetscErrorCode MatGetCurrentMemType_SeqAIJB200(Mat A, PetscMemType *mtype)
{
  PetscFunctionBegin;
  *mtype = PB_BoundToCPU(A) ? PETSC_MEMTYPE_HOST : PETSC_MEMTYPE_CUDA;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* COO assembly (aij.c:4524-4732; device analogue MatSetPreallocationCOO_SeqAIJCUSPARSE / MatSetValuesCOO_SeqAIJCUSPARSE).
   The pattern is built by the parent on the host (the host CSR is this type's master copy); index arrays that live on the
   device are brought down first.  The device plan adopts the parent's jmap/perm, so values assembled on the device from a
   device-resident v[] are bit-identical to the reference's, then mirrored into the host array. */
static PetscErrorCode MatSetPreallocationCOO_SeqAIJB200(Mat A, PetscCount n, PetscInt coo_i[], PetscInt coo_j[])
{
  Mat_B200            *m = (Mat_B200 *)A->spptr;
  int                  dev_i = 0, dev_j = 0;
  PetscInt            *hi = coo_i, *hj = coo_j;
  PetscContainer       container;
  MatCOOStruct_SeqAIJ *coo;
  PetscFunctionBegin;
  PetscCallB200(b200PointerIsDevice(coo_i, &dev_i));
  PetscCallB200(b200PointerIsDevice(coo_j, &dev_j));
  if (dev_i) {
    PetscCall(PetscMalloc1(n, &hi));
    PetscCallB200(b200MemcpyDtoH(PB_h, hi, coo_i, sizeof(PetscInt) * (size_t)n));
  }
  if (dev_j) {
    PetscCall(PetscMalloc1(n, &hj));
    PetscCallB200(b200MemcpyDtoH(PB_h, hj, coo_j, sizeof(PetscInt) * (size_t)n));
  }
  PetscCall((*m->coo_prealloc_seqaij)(A, n, hi, hj)); /* MatSetPreallocationCOO_SeqAIJ: replaces i/j/a, keeps ops and spptr */
  if (dev_i) PetscCall(PetscFree(hi));
  if (dev_j) PetscCall(PetscFree(hj));
  if (m->coo) b200CooPlanDestroy(m->coo);
  m->coo = NULL;
  PetscCall(PetscObjectQuery((PetscObject)A, "__PETSc_MatCOOStruct_Host", (PetscObject *)&container));
  PetscCheck(container, PETSC_COMM_SELF, PETSC_ERR_PLIB, "Not found MatCOOStruct on this matrix");
  PetscCall(PetscContainerGetPointer(container, (void **)&coo));
  PetscCallB200(b200CooPlanCreateFromMaps(PB_h, (int64_t)coo->nz, (int64_t)coo->Atot, (const int64_t *)coo->jmap, (const int64_t *)coo->perm, &m->coo));
  m->valid = PETSC_FALSE; /* new pattern: rebuild the mirror on next use */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatSetValuesCOO_SeqAIJB200(Mat A, const PetscScalar v[], InsertMode imode)
{
  Mat_B200        *m = (Mat_B200 *)A->spptr;
  Mat_SeqAIJ      *a = (Mat_SeqAIJ *)A->data;
  int              dev = 0;
  PetscObjectState st;
  PetscFunctionBegin;
  PetscCallB200(b200PointerIsDevice(v, &dev));
  if (!dev || !m->coo || PB_BoundToCPU(A)) { /* host values: the parent's loop; the mirror follows through the object state */
    PetscCheck(!dev, PETSC_COMM_SELF, PETSC_ERR_SUP, "device COO values on a matrix bound to the CPU");
    PetscCall((*m->coo_setvalues_seqaij)(A, v, imode));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_MatSyncEx(A, PETSC_FALSE)); /* MatSetValuesCOO() assembles AFTER this method (gcreate.c MatSetValuesCOO) */
  PetscCallB200(b200CooSetValues(PB_h, m->coo, v, imode == INSERT_VALUES, m->d_a));
  PetscCallB200(b200MemcpyDtoH(PB_h, a->a, m->d_a, sizeof(double) * (size_t)a->nz)); /* host master copy follows */
  PetscCall(PetscObjectStateIncrease((PetscObject)A));
  PetscCall(PetscObjectStateGet((PetscObject)A, &st));
  m->valstate = st; /* device and host agree */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode MatDestroy_SeqAIJB200(Mat A)
{
  Mat_B200 *m = (Mat_B200 *)A->spptr;
  PetscErrorCode (*destroy)(Mat) = m->destroy_seqaij;
  PetscFunctionBegin;
  PetscCall(PB_MatFreeDevice(m));
  if (m->coo) b200CooPlanDestroy(m->coo);
  PetscCall(PetscFree(A->spptr));
  PetscCall(PetscObjectComposeFunction((PetscObject)A, "MatConvert_seqaij_seqaijb200_C", NULL));
  PetscCall((*destroy)(A));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode MatConvert_SeqAIJ_SeqAIJB200(Mat A, MatType type, MatReuse reuse, Mat *newmat);
static PetscErrorCode       MatDuplicate_SeqAIJB200(Mat A, MatDuplicateOption op, Mat *B)
{
  Mat_B200 *m = (Mat_B200 *)A->spptr;
  PetscFunctionBegin;
  PetscCall((*m->duplicate_seqaij)(A, op, B));
  PetscCall(MatConvert_SeqAIJ_SeqAIJB200(*B, MATSEQAIJB200, MAT_INPLACE_MATRIX, B));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode MatConvert_SeqAIJ_SeqAIJB200(Mat A, MatType type, MatReuse reuse, Mat *newmat)
{
  Mat       B;
  Mat_B200 *m;

  PetscFunctionBegin;
  (void)type;
  PetscCall(PB_Init());
  if (reuse == MAT_INITIAL_MATRIX) PetscCall(MatDuplicate(A, MAT_COPY_VALUES, newmat));
  else if (reuse == MAT_REUSE_MATRIX) PetscCall(MatCopy(A, *newmat, SAME_NONZERO_PATTERN));
  B = *newmat;
  if (B->spptr && B->ops->mult == MatMult_SeqAIJB200) PetscFunctionReturn(PETSC_SUCCESS); /* already converted */
  PetscCall(PetscFree(B->defaultvectype));
  PetscCall(PetscStrallocpy(VECSEQB200, &B->defaultvectype)); /* MatCreateVecs hands out device vectors (matrix.c:10069) */
  PetscCall(PetscNew(&m));
  m->destroy_seqaij   = B->ops->destroy;
  m->duplicate_seqaij = B->ops->duplicate;
  m->mult_seqaij      = B->ops->mult;
  m->multadd_seqaij   = B->ops->multadd;
  m->multtranspose_seqaij    = B->ops->multtranspose;
  m->multtransposeadd_seqaij = B->ops->multtransposeadd;
  PetscCall(PetscObjectQueryFunction((PetscObject)B, "MatSetPreallocationCOO_C", &m->coo_prealloc_seqaij));
  PetscCall(PetscObjectQueryFunction((PetscObject)B, "MatSetValuesCOO_C", &m->coo_setvalues_seqaij));
  B->spptr            = m;
  B->ops->mult        = MatMult_SeqAIJB200;
  B->ops->multadd     = MatMultAdd_SeqAIJB200;
  B->ops->getdiagonal = MatGetDiagonal_SeqAIJB200;
  B->ops->destroy     = MatDestroy_SeqAIJB200;
  B->ops->duplicate   = MatDuplicate_SeqAIJB200;
  B->ops->multtranspose     = MatMultTranspose_SeqAIJB200;
  B->ops->multtransposeadd  = MatMultTransposeAdd_SeqAIJB200;
  B->ops->bindtocpu         = MatBindToCPU_SeqAIJB200;
  B->ops->getcurrentmemtype = MatGetCurrentMemType_SeqAIJB200;
  if (m->coo_prealloc_seqaij && m->coo_setvalues_seqaij) {
    PetscCall(PetscObjectComposeFunction((PetscObject)B, "MatSetPreallocationCOO_C", MatSetPreallocationCOO_SeqAIJB200));
    PetscCall(PetscObjectComposeFunction((PetscObject)B, "MatSetValuesCOO_C", MatSetValuesCOO_SeqAIJB200));
  }
  ((Mat_SeqAIJ *)B->data)->inode.use = PETSC_FALSE; /* the device kernel is the mult path */
  PetscCall(PetscObjectChangeTypeName((PetscObject)B, MATSEQAIJB200));
  PetscCall(PetscObjectComposeFunction((PetscObject)B, "MatConvert_seqaij_seqaijb200_C", MatConvert_SeqAIJ_SeqAIJB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode MatCreate_SeqAIJB200(Mat B)
{
  PetscFunctionBegin;
  PetscCall(MatCreate_SeqAIJ(B)); /* exported creator of the parent (aij.c:4745) */
  PetscCall(MatConvert_SeqAIJ_SeqAIJB200(B, MATSEQAIJB200, MAT_INPLACE_MATRIX, &B));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== MatSolverType "b200": ILU(0) on the device */
typedef struct {
  b200IluPlan plan;
  double     *d_aval_tmp; /* when A is a plain seqaij (no device mirror) */
  PetscInt    n;
  double      nz;
} MatFactor_B200;

static PetscErrorCode MatDestroy_FactorB200(Mat F)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  PetscFunctionBegin;
  if (f) {
    if (f->plan) PetscCallB200(b200Ilu0Destroy(f->plan));
    PetscCallB200(b200Free(PB_h, f->d_aval_tmp));
    PetscCall(PetscFree(F->data));
  }
  PetscCall(PetscObjectComposeFunction((PetscObject)F, "MatFactorGetSolverType_C", NULL));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatSolve_FactorB200(Mat F, Vec b, Vec x)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  const double   *db;
  double         *dx;
  PetscFunctionBegin;
  if (!PB_IsB200(b) || !PB_IsB200(x)) { /* host vectors: stage through temporaries of the device type */
    Vec tb, tx;
    PetscCall(VecCreateSeq(PETSC_COMM_SELF, F->rmap->n, &tb));
    PetscCall(VecSetType(tb, VECSEQB200));
    PetscCall(VecDuplicate(tb, &tx));
    PetscCall(VecCopy(b, tb));
    PetscCall(MatSolve_FactorB200(F, tb, tx));
    PetscCall(VecCopy(tx, x));
    PetscCall(VecDestroy(&tb));
    PetscCall(VecDestroy(&tx));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_VecRead(b, &db));
  PetscCall(PB_VecWrite(x, &dx));
  PetscCallB200(b200Ilu0Solve(PB_h, f->plan, db, dx)); /* MatSolve_SeqAIJ_NaturalOrdering, aijfact.c:2413 */
  PetscCall(PetscLogFlops(2.0 * f->nz - F->cmap->n));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatLUFactorNumeric_FactorB200(Mat F, Mat A, const MatFactorInfo *info)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  const double   *d_a;
  int             nshift = 0;
  PetscFunctionBegin;
  if (A->ops->mult == MatMult_SeqAIJB200) {
    PetscCall(PB_MatSync(A));
    d_a = ((Mat_B200 *)A->spptr)->d_a;
  } else {
    Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
    if (!f->d_aval_tmp) PetscCallB200(b200Malloc(PB_h, (void **)&f->d_aval_tmp, sizeof(double) * ((size_t)a->nz + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, f->d_aval_tmp, a->a, sizeof(double) * (size_t)a->nz));
    d_a = f->d_aval_tmp;
  }
  PetscCallB200(b200Ilu0Numeric(PB_h, f->plan, d_a, info->zeropivot, info->shiftamount > 0 ? info->shiftamount : 100.0 * PETSC_MACHINE_EPSILON, &nshift));
  F->ops->solve    = MatSolve_FactorB200;
  F->assembled     = PETSC_TRUE;
  F->preallocated  = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatILUFactorSymbolic_FactorB200(Mat F, Mat A, IS isrow, IS iscol, const MatFactorInfo *info)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  Mat_SeqAIJ     *a = (Mat_SeqAIJ *)A->data;
  PetscBool       idr = PETSC_TRUE, idc = PETSC_TRUE;
  PetscFunctionBegin;
  PetscCheck(info->levels == 0, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSolverType b200 implements ILU(0) only (got %g levels)", (double)info->levels);
  if (isrow) PetscCall(ISIdentity(isrow, &idr));
  if (iscol) PetscCall(ISIdentity(iscol, &idc));
  PetscCheck(idr && idc, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSolverType b200 requires the natural ordering (-pc_factor_mat_ordering_type natural)");
  if (f->plan) PetscCallB200(b200Ilu0Destroy(f->plan));
  f->plan = NULL;
  PetscCallB200(b200Ilu0Symbolic(PB_h, (int)A->rmap->n, a->i, a->j, &f->plan)); /* layout of aijfact.c:1471-1534 + level schedule */
  f->nz                        = (double)a->nz;
  F->ops->lufactornumeric      = MatLUFactorNumeric_FactorB200;
  F->info.fill_ratio_given     = info->fill;
  F->info.fill_ratio_needed    = 1.0;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetInfo_FactorB200(Mat F, MatInfoType flag, MatInfo *info)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  PetscFunctionBegin;
  (void)flag;
  PetscCall(PetscMemzero(info, sizeof(*info)));
  info->block_size        = 1.0;
  info->nz_allocated      = f->nz;
  info->nz_used           = f->nz;
  info->fill_ratio_given  = F->info.fill_ratio_given;
  info->fill_ratio_needed = 1.0;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatFactorGetSolverType_B200(Mat F, MatSolverType *type)
{
  PetscFunctionBegin;
  (void)F;
  *type = MATSOLVERB200;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetFactor_seqaijb200_b200(Mat A, MatFactorType ftype, Mat *F)
{
  MatFactor_B200 *f;
  PetscInt        n = A->rmap->n;
  PetscFunctionBegin;
  PetscCheck(ftype == MAT_FACTOR_ILU, PetscObjectComm((PetscObject)A), PETSC_ERR_SUP, "MatSolverType b200 provides MAT_FACTOR_ILU only");
  PetscCall(PB_Init());
  PetscCall(MatCreate(PetscObjectComm((PetscObject)A), F));
  PetscCall(MatSetSizes(*F, n, n, n, n));
  PetscCall(PetscLayoutSetUp((*F)->rmap));
  PetscCall(PetscLayoutSetUp((*F)->cmap));
  PetscCall(PetscObjectChangeTypeName((PetscObject)*F, "seqaijb200factor"));
  PetscCall(PetscNew(&f));
  f->n                          = n;
  (*F)->data                    = f;
  (*F)->factortype              = ftype;
  (*F)->canuseordering          = PETSC_FALSE; /* natural ordering: PCSetUp_ILU then skips MatGetOrdering (ilu.c:127) */
  (*F)->ops->ilufactorsymbolic  = MatILUFactorSymbolic_FactorB200;
  (*F)->ops->destroy            = MatDestroy_FactorB200;
  (*F)->ops->getinfo            = MatGetInfo_FactorB200;
  (*F)->preallocated            = PETSC_TRUE;
  PetscCall(PetscFree((*F)->solvertype));
  PetscCall(PetscStrallocpy(MATSOLVERB200, &(*F)->solvertype));
  PetscCall(PetscFree((*F)->defaultvectype));
  PetscCall(PetscStrallocpy(VECSEQB200, &(*F)->defaultvectype));
  PetscCall(PetscObjectComposeFunction((PetscObject)*F, "MatFactorGetSolverType_C", MatFactorGetSolverType_B200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== PC "jacobib200": PCJACOBI (diagonal) with a fused applyBA
   KSPGMRESCycle reaches the operator through KSP_PCApplyBAorAB -> PCApplyBAorAB, which calls ops->applyBA when the PC has one
   (precon.c:853-854): with left preconditioning y = D^-1 (A x) is then ONE kernel (b200CsrSpMVJacobi) instead of
   MatMult + VecPointwiseMult -- 24 B/row less traffic, same rounding (row sum first, then one multiply).
   Setup follows PCSetUp_Jacobi (jacobi.c:172-270): MatGetDiagonal, reciprocal, zero diagonal -> 1.0. */
typedef struct {
  Vec dinv;
} PC_JacobiB200;

static PetscErrorCode PCSetUp_JacobiB200(PC pc)
{
  PC_JacobiB200 *jac = (PC_JacobiB200 *)pc->data;
  PetscInt       n;
  PetscFunctionBegin;
  if (!jac->dinv) PetscCall(MatCreateVecs(pc->pmat, &jac->dinv, NULL));
  PetscCall(MatGetDiagonal(pc->pmat, jac->dinv));
  PetscCall(VecGetLocalSize(jac->dinv, &n));
  if (PB_IsB200(jac->dinv)) {
    double *d;
    int     nzero = 0;
    PetscCall(PB_VecRW(jac->dinv, &d));
    PetscCallB200(b200JacobiInvertDiagonal(PB_h, (int64_t)n, d, d, &nzero));
    if (nzero) PetscCall(PetscInfo(pc, "Zero detected in diagonal of matrix, using 1 at those locations\n"));
  } else {
    PetscScalar *x;
    PetscCall(VecReciprocal(jac->dinv));
    PetscCall(VecGetArray(jac->dinv, &x));
    for (PetscInt i = 0; i < n; i++)
      if (x[i] == 0.0) x[i] = 1.0;
    PetscCall(VecRestoreArray(jac->dinv, &x));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PCApply_JacobiB200(PC pc, Vec x, Vec y)
{
  PC_JacobiB200 *jac = (PC_JacobiB200 *)pc->data;
  PetscFunctionBegin;
  PetscCall(VecPointwiseMult(y, x, jac->dinv)); /* jacobi.c:354 */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PCApplyBA_JacobiB200(PC pc, PCSide side, Vec x, Vec y, Vec work)
{
  PC_JacobiB200 *jac = (PC_JacobiB200 *)pc->data;
  Mat            A   = pc->mat;
  PetscFunctionBegin;
  if (side == PC_LEFT && A->ops->mult == MatMult_SeqAIJB200 && !PB_BoundToCPU(A) && PB_IsB200(x) && PB_IsB200(y) && PB_IsB200(jac->dinv)) {
    Mat_B200     *m = (Mat_B200 *)A->spptr;
    Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
    const double *dx, *dd;
    double       *dy;
    PetscCall(PB_MatSync(A));
    PetscCall(PB_VecRead(x, &dx));
    PetscCall(PB_VecRead(jac->dinv, &dd));
    PetscCall(PB_VecWrite(y, &dy));
    PetscCallB200(b200CsrSpMVJacobi(PB_h, m->plan, m->d_a, dx, dd, dy, NULL));
    PetscCall(PetscLogFlops(2.0 * a->nz - a->nonzerorowcnt + A->rmap->n));
  } else if (side == PC_LEFT) {
    PetscCall(MatMult(A, x, work));
    PetscCall(PCApply_JacobiB200(pc, work, y));
  } else if (side == PC_RIGHT) {
    PetscCall(PCApply_JacobiB200(pc, x, work));
    PetscCall(MatMult(A, work, y));
  } else SETERRQ(PetscObjectComm((PetscObject)pc), PETSC_ERR_SUP, "jacobib200 has no symmetric application; use -pc_type jacobi");
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PCReset_JacobiB200(PC pc)
{
  PC_JacobiB200 *jac = (PC_JacobiB200 *)pc->data;
  PetscFunctionBegin;
  PetscCall(VecDestroy(&jac->dinv));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PCDestroy_JacobiB200(PC pc)
{
  PetscFunctionBegin;
  PetscCall(PCReset_JacobiB200(pc));
  PetscCall(PetscFree(pc->data));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PETSC_EXTERN PetscErrorCode PCCreate_JacobiB200(PC pc)
{
  PC_JacobiB200 *jac;
  PetscFunctionBegin;
  PetscCall(PB_Init());
  PetscCall(PetscNew(&jac));
  pc->data                = jac;
  pc->ops->setup          = PCSetUp_JacobiB200;
  pc->ops->apply          = PCApply_JacobiB200;
  pc->ops->applytranspose = PCApply_JacobiB200;
  pc->ops->applyBA        = PCApplyBA_JacobiB200;
  pc->ops->reset          = PCReset_JacobiB200;
  pc->ops->destroy        = PCDestroy_JacobiB200;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== registration (src/sys/dll/reg.c:79,150; dl.c:178-199) */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_petscb200plugin(void)
{
  PetscFunctionBegin;
  PetscCall(VecRegister(VECSEQB200, VecCreate_SeqB200));
  PetscCall(VecRegister(VECB200, VecCreate_SeqB200));
  PetscCall(MatRegisterRootName(MATAIJB200, MATSEQAIJB200, "mpiaijb200"));
  PetscCall(MatRegister(MATSEQAIJB200, MatCreate_SeqAIJB200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJB200, MAT_FACTOR_ILU, MatGetFactor_seqaijb200_b200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJ, MAT_FACTOR_ILU, MatGetFactor_seqaijb200_b200));
  PetscCall(PCRegister(PCJACOBIB200, PCCreate_JacobiB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}
