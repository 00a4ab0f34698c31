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
 *   MatSolverTypeRegister  "b200" for seqaijb200, MAT_FACTOR_ILU and MAT_FACTOR_ICC   (src/mat/interface/matrix.c:4720)
 *   PCRegister             "jacobib200" and (unless -b200_keep_pcjacobi) "jacobi": the reference's PCJACOBI sub-classed with a
 *                          fused ops->applyBA (src/ksp/pc/interface/pcregis.c, precon.c:810-865)
 *   KSPRegister            "pipecgb200": single-reduction CG, one reduction kernel + one recurrence kernel per iteration
 *                          "pgmresb200": pipelined GMRES, the one reduction of an iteration is read one iteration after its launch
 *   VecRegister            "mpib200";  MatRegister "mpiaijb200": the row-partitioned types over NCCL ranks (one process per GPU)
 *   PetscSFRegister        "b200" and (unless -b200_keep_sfbasic) "basic": PETSCSFBASIC sub-classed so that VecScatter / PetscSF
 *                          broadcasts and reductions on device data run as device kernels (src/vec/is/sf/interface/sfregi.c:78)
 *
 * Structure mirrors the reference's own device subclassing (aijcusparse.cu:2807-2868, veccupmimpl.h:994-1047): create the
 * parent (MATSEQAIJ / VECSEQ), keep its host data structures, overwrite the ops of the hot path with functions that run
 * on a device mirror, and keep host and device coherent with an offload mask.  Everything the plugin does not override
 * falls back to the parent's host implementation, which reaches the data through VecGetArray*() -> our hooks.
 *
 * The PETSc this is built against is MPIUNI (no MPI in the image): every process is a one-rank PETSc.  The row-partitioned
 * types mpiaijb200 / mpib200 therefore span the NCCL ranks the launcher joins (see PB_Init), not an MPI communicator.
 * Compile: see petsc_plugin/Makefile (needs PETSC_DIR/PETSC_ARCH of the PETSc the application links).
 */
#include <petsc/private/vecimpl.h>
#include <petsc/private/matimpl.h>
#include <petsc/private/pcimpl.h>
#include <../src/vec/vec/impls/dvecimpl.h>
#include <../src/mat/impls/aij/seq/aij.h>
#include <petsc/private/kspimpl.h>
#include <petscksp.h>
#include "petscb200.h"

/* the C ABI moves int32 indices and real double scalars (include/petscb200.h "Conventions"): refuse other PETSc builds at
   compile time instead of uploading garbage */
#if defined(PETSC_USE_64BIT_INDICES) || defined(PETSC_USE_COMPLEX) || !defined(PETSC_USE_REAL_DOUBLE)
  #error "libpetscb200plugin needs a PETSc configured with 32-bit PetscInt and real double-precision PetscScalar"
#endif

#define VECSEQB200    "seqb200"
#define VECB200       "b200"
#define MATSEQAIJB200 "seqaijb200"
#define MATAIJB200    "aijb200"
#define MATSOLVERB200 "b200"
#define PCJACOBIB200  "jacobib200"
#define KSPPIPECGB200 "pipecgb200"
#define KSPPGMRESB200 "pgmresb200"

#define VECMPIB200    "mpib200"
#define MATMPIAIJB200 "mpiaijb200"

static b200Handle PB_h    = NULL;
static int        PB_rank = 0, PB_size = 1; /* NCCL ranks (one process per GPU); 0/1 without a communicator */

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

/* Device and communicator.  One process drives one GPU.  The device is -b200_device <i>, else $PETSCB200_DEVICE, else
   $LOCAL_RANK (torchrun), else the current device.  With a real MPI underneath PETSc the NCCL id would travel over
   MPI_Bcast on PETSC_COMM_WORLD; this PETSc is MPIUNI (one rank per process, no MPI in the image), so N processes are
   joined by the launcher instead: $PETSCB200_NRANKS, $PETSCB200_RANK and $PETSCB200_NCCL_ID (the 128-byte ncclUniqueId as
   256 hex digits, made by b200CommGetUniqueId on rank 0 and handed to every process).  The row-partitioned types
   mpiaijb200 / mpib200 then exchange halos and reduce over NCCL/NVLink. */
static PetscErrorCode PB_Init(void)
{
  PetscFunctionBegin;
  if (!PB_h) {
    PetscInt    dev = -1;
    PetscBool   set = PETSC_FALSE;
    const char *e;
    PetscCall(PetscOptionsGetInt(NULL, NULL, "-b200_device", &dev, &set));
    if (!set && (e = getenv("PETSCB200_DEVICE"))) dev = (PetscInt)atoi(e);
    else if (!set && getenv("PETSCB200_NRANKS") && (e = getenv("LOCAL_RANK"))) dev = (PetscInt)atoi(e);
    PetscCallB200(b200Create(&PB_h, (int)dev));
#if defined(PETSC_HAVE_CUDA)
    PetscCallB200(b200SetStream(PB_h, (void *)PetscDefaultCudaStream)); /* include/petscdevice_cuda.h:180 */
#endif
    if ((e = getenv("PETSCB200_NRANKS")) && atoi(e) > 1) {
      const char   *id = getenv("PETSCB200_NCCL_ID"), *r = getenv("PETSCB200_RANK");
      unsigned char uid[B200_UNIQUE_ID_BYTES];
      PetscCheck(id && r && strlen(id) == 2 * B200_UNIQUE_ID_BYTES, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "PETSCB200_NRANKS > 1 needs PETSCB200_RANK and PETSCB200_NCCL_ID (256 hex digits)");
      for (int i = 0; i < B200_UNIQUE_ID_BYTES; i++) {
        unsigned int byte;
        PetscCheck(sscanf(id + 2 * i, "%2x", &byte) == 1, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "PETSCB200_NCCL_ID is not hexadecimal");
        uid[i] = (unsigned char)byte;
      }
      PB_size = atoi(e);
      PB_rank = atoi(r);
      PetscCheck(PB_rank >= 0 && PB_rank < PB_size, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "PETSCB200_RANK %d outside [0,%d)", PB_rank, PB_size);
      PetscCallB200(b200CommInitRank(PB_h, PB_size, PB_rank, uid));
    }
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* for PETSc programs that link the plugin: the library handle (CUDA-event timing, generators) and the NCCL rank/size */
PETSC_EXTERN PetscErrorCode PetscB200GetHandle(void **handle, int *rank, int *size)
{
  PetscFunctionBegin;
  PetscCall(PB_Init());
  if (handle) *handle = (void *)PB_h;
  if (rank) *rank = PB_rank;
  if (size) *size = PB_size;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* PETSc-side profiling of device work (aijcusparse.cu:2463,2536,2565-2566): -log_view books these flops and this time in
   the GPU columns when PETSc has a device back end; in a host-only PETSc PetscLogGpuFlops is a no-op macro and
   PetscLogGpuTime{Begin,End} do nothing, so the flops are additionally booked with PetscLogFlops as before */
#if PetscDefined(HAVE_DEVICE)
  #define PB_LogFlops(f)    PetscLogGpuFlops(f)
  #define PB_LogTimeBegin() PetscLogGpuTimeBegin()
  #define PB_LogTimeEnd()   PetscLogGpuTimeEnd()
#else
  #define PB_LogFlops(f)    PetscLogFlops(f)
  #define PB_LogTimeBegin() PETSC_SUCCESS
  #define PB_LogTimeEnd()   PETSC_SUCCESS
#endif

/* ================================================================== Vec: seqb200 */
enum { PB_UNALLOCATED = 0, PB_CPU = 1, PB_GPU = 2, PB_BOTH = 3 }; /* PetscOffloadMask, include/petscdevicetypes.h:240 */

typedef struct {
  Vec_Seq seq;    /* MUST be first: the parent's data, untouched (host array, VECHEADER) */
  double *d;      /* device mirror */
  int     mask;
  double *d_sumsq; /* fused MAXPY+norm */
  double *h_sumsq; /* non-NULL: d_sumsq is the device alias of this mapped pinned host scalar (one rank: no all-reduce on it) */
  PetscObjectState sumsq_state;
  int              host_writers; /* VecGetArray / VecGetArrayWrite handed the host array out for writing and it is not restored yet */
  struct PB_Slab  *slab;       /* VecDuplicateVecs: d points into one shared device allocation (bvec2.c:670-691) */
  double          *lv_saved_d; /* VecGetLocalVector*: this vector's own device array while it aliases another's */
  int              lv_saved_mask, lv_active;
  PetscScalar     *host_owned; /* lazily allocated host array that is not registered with the parent (see PB_VecHostAlloc) */
} Vec_SeqB200;
struct PB_Slab {
  double *base;
  int     refs;
};

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
/* Host storage is LAZY: VecCreate_Seq allocates and zero-fills n scalars for every vector (bvec3.c:33-34), which for the 34
   vectors of a GMRES(30) basis at 512^3 is 36 GB of page-faulting memset (14 s measured) that a device-resident solve
   never reads.  The host array is created on first host access; a vector nobody has written yet reads as zeros on either
   side (VecCreate_Seq's guarantee). */
static PetscErrorCode PB_VecHostAlloc(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  if (!b->seq.array && v->map->n) {
    PetscScalar *a;
    PetscCall(PetscMalloc1(v->map->n, &a));
    if (b->mask == PB_UNALLOCATED || b->mask == PB_CPU) PetscCall(PetscArrayzero(a, v->map->n));
    b->seq.array = a;
    if (!b->seq.unplacedarray) b->seq.array_allocated = a; /* VecDestroy_Seq frees it */
    else b->host_owned = a;                                 /* a user array is placed on top: remember ours for the destroy */
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecToDevice(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecAlloc(v));
  if (b->mask == PB_CPU || b->mask == PB_UNALLOCATED) {
    if (v->map->n) {
      if (b->seq.array) {
        PetscCallB200(b200MemcpyHtoDAsync(PB_h, b->d, b->seq.array, sizeof(double) * (size_t)v->map->n));
        PetscCall(PetscLogCpuToGpu((PetscLogDouble)(sizeof(double) * (size_t)v->map->n)));
      } else PetscCallB200(b200Memset(PB_h, b->d, 0, sizeof(double) * (size_t)v->map->n)); /* never written: zeros */
    }
    b->mask = b->seq.array ? PB_BOTH : PB_GPU;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecToHost(Vec v)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecHostAlloc(v));
  if (b->mask == PB_GPU) {
    if (v->map->n) PetscCallB200(b200MemcpyDtoH(PB_h, b->seq.array, b->d, sizeof(double) * (size_t)v->map->n));
    PetscCall(PetscLogGpuToCpu((PetscLogDouble)(sizeof(double) * (size_t)v->map->n)));
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
/* Every device WRITE access bumps the vector's object state, as VecRestoreArray[Write]() / VecCUDARestoreArrayWrite() do for the
   reference's implementations (rvector.c:2085, veccupmimpl.h): PETSc keys its norm cache on that state (VecNormAvailable), and
   e.g. VecSet(x, 0) returns early when the cached 2-norm is 0 (rvector.c:506-512).  MatMult / PCApply / PCApplyBAorAB do NOT bump
   the state of their output themselves -- the implementation's array access does.  Found by the reference's ex9 (BiCGStab's
   VecSet(V,0) skipped on the second solve) and mat/tests/ex254 (MatMultEqual reading a stale cached norm). */
static PetscErrorCode PB_VecRW(Vec v, double **p)
{
  PetscFunctionBegin;
  PetscCall(PB_VecToDevice(v));
  ((Vec_SeqB200 *)v->data)->mask = PB_GPU;
  *p = ((Vec_SeqB200 *)v->data)->d;
  PetscCall(PetscObjectStateIncrease((PetscObject)v));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_VecWrite(Vec v, double **p)
{
  PetscFunctionBegin;
  PetscCall(PB_VecAlloc(v));
  ((Vec_SeqB200 *)v->data)->mask = PB_GPU;
  *p = ((Vec_SeqB200 *)v->data)->d;
  PetscCall(PetscObjectStateIncrease((PetscObject)v));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* host access hooks (rvector.c:2070-2160 dispatch to these because ops->getarray is set) */
static PetscErrorCode VecGetArray_SeqB200(Vec v, PetscScalar **a)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  PetscCall(PB_VecToHost(v));
  b->mask = PB_CPU;
  b->host_writers++;
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
  PetscCall(PB_VecHostAlloc(v));
  b->mask = PB_CPU;
  b->host_writers++;
  *a      = b->seq.array;
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* the end of a host WRITE access: the host array is the valid copy from here on, whatever happened in between (the reference's
   device vectors set PETSC_OFFLOAD_CPU in their restore as well).  Programs do call Vec operations on a vector whose array they
   hold -- ts/tutorials/ex10.c: DMDAVecGetArray(F); VecZeroEntries(F); fill; Restore -- which on VECSEQ act on the very memory the
   caller is filling; VecSet therefore works on the host array while it is handed out (VecSet_SeqB200). */
static PetscErrorCode VecRestoreArray_SeqB200(Vec v, PetscScalar **a)
{
  Vec_SeqB200 *b = (Vec_SeqB200 *)v->data;
  PetscFunctionBegin;
  if (b->host_writers > 0) b->host_writers--;
  b->mask = PB_CPU;
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
  if (!x->map->n) PetscFunctionReturn(PETSC_SUCCESS);
  if (((Vec_SeqB200 *)x->data)->host_writers > 0) { /* somebody holds the host array for writing: set THAT memory, as VECSEQ would */
    PetscCall((*PB_VecSeqOps.set)(x, a));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
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
  if (!b->d_sumsq) {
    /* one rank: the kernel writes |x|^2 straight into mapped pinned host memory, VecNorm then only synchronises the stream (no
       cudaMemcpy on the critical path of every iteration: config 1 is latency bound); several ranks: device memory, because the
       value is all-reduced by NCCL first */
    if (PB_size == 1) PetscCallB200(b200MallocMapped((void **)&b->h_sumsq, (void **)&b->d_sumsq, sizeof(double)));
    else PetscCallB200(b200Malloc(PB_h, (void **)&b->d_sumsq, sizeof(double)));
  }
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
    if (b->h_sumsq) {
      PetscCallB200(b200Synchronize(PB_h));
      ss = *(volatile double *)b->h_sumsq;
    } else PetscCallB200(b200MemcpyDtoH(PB_h, &ss, b->d_sumsq, sizeof(double)));
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
    if (b->lv_active) b->d = b->lv_saved_d; /* destroyed while aliasing another vector: that array is not ours */
    if (b->slab) {
      if (--b->slab->refs == 0) {
        PetscCallB200(b200Free(PB_h, b->slab->base));
        PetscCall(PetscFree(b->slab));
      }
    } else if (b->d) PetscCallB200(b200Free(PB_h, b->d));
    if (b->h_sumsq) PetscCallB200(b200FreeHost(b->h_sumsq));
    else if (b->d_sumsq) PetscCallB200(b200Free(PB_h, b->d_sumsq));
    b->h_sumsq = NULL;
    if (b->host_owned) {
      if (b->seq.array == b->host_owned) b->seq.array = NULL;
      PetscCall(PetscFree(b->host_owned));
    }
    b->d = b->d_sumsq = NULL;
    b->slab           = NULL;
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
  if (((Vec_SeqB200 *)v->data)->mask == PB_GPU) PetscCall(PB_VecToHost(v)); /* keep the current values in the array being set aside */
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

/* remaining reductions / element-wise ops of the VECSEQ table, on the device (no PCIe round trip of the whole vector) */
static PetscErrorCode VecSum_SeqB200(Vec x, PetscScalar *s)
{
  const double *dx;
  PetscFunctionBegin;
  PetscCall(PB_VecRead(x, &dx));
  PetscCallB200(b200VecSum(PB_h, N_(x), dx, s));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMax_SeqB200(Vec x, PetscInt *p, PetscReal *v)
{
  const double *dx;
  int64_t       idx = -1;
  PetscFunctionBegin;
  PetscCall(PB_VecRead(x, &dx));
  PetscCallB200(b200VecMax(PB_h, N_(x), dx, &idx, v));
  if (p) *p = (PetscInt)idx;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMin_SeqB200(Vec x, PetscInt *p, PetscReal *v)
{
  const double *dx;
  int64_t       idx = -1;
  PetscFunctionBegin;
  PetscCall(PB_VecRead(x, &dx));
  PetscCallB200(b200VecMin(PB_h, N_(x), dx, &idx, v));
  if (p) *p = (PetscInt)idx;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecShift_SeqB200(Vec x, PetscScalar a)
{
  double *d;
  PetscFunctionBegin;
  PetscCall(PB_VecRW(x, &d));
  PetscCallB200(b200VecShift(PB_h, N_(x), a, d));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* VecGetLocalVector[Read] / VecRestoreLocalVector[Read] (rvector.c:1905-2060; device analogue veccupmimpl.h:1040-1043):
   w aliases v's DEVICE array -- PCApply_BJacobi_Singleblock does this on every application (bjacobi.c:579-598), and the
   default implementation goes through VecGetArray, i.e. two PCIe round trips of the vector per iteration */
static PetscErrorCode PB_LocalVectorBegin(Vec v, Vec w, PetscBool write)
{
  Vec_SeqB200 *bv = (Vec_SeqB200 *)v->data, *bw;
  PetscFunctionBegin;
  if (!PB_IsB200(w)) { /* a host vector as the window: the default implementation's path */
    PetscScalar *a;
    if (write) PetscCall(VecGetArray(v, &a));
    else PetscCall(VecGetArrayRead(v, (const PetscScalar **)&a));
    PetscCall(VecPlaceArray(w, a));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  bw = (Vec_SeqB200 *)w->data;
  PetscCheck(!bw->lv_active, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "the local vector already maps another vector");
  PetscCall(PB_VecToDevice(v));
  bw->lv_saved_d    = bw->d;
  bw->lv_saved_mask = bw->mask;
  bw->lv_active     = 1;
  bw->d             = bv->d;
  bw->mask          = PB_GPU;
  if (write) bv->mask = PB_GPU;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_LocalVectorEnd(Vec v, Vec w, PetscBool write)
{
  Vec_SeqB200 *bw;
  PetscFunctionBegin;
  if (!PB_IsB200(w)) {
    const PetscScalar *a;
    PetscCall(VecGetArrayRead(w, &a));
    if (write) PetscCall(VecRestoreArray(v, (PetscScalar **)&a));
    else PetscCall(VecRestoreArrayRead(v, &a));
    PetscCall(VecResetArray(w));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  bw = (Vec_SeqB200 *)w->data;
  PetscCheck(bw->lv_active, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "the local vector maps nothing");
  if (bw->mask != PB_GPU) { /* somebody read or wrote w on the host meanwhile: bring that back into the shared device array */
    if (bw->mask == PB_CPU && w->map->n) PetscCallB200(b200MemcpyHtoD(PB_h, bw->d, bw->seq.array, sizeof(double) * (size_t)w->map->n));
  }
  bw->d         = bw->lv_saved_d;
  bw->mask      = bw->lv_saved_mask;
  bw->lv_active = 0;
  if (write) ((Vec_SeqB200 *)v->data)->mask = PB_GPU;
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* VecGetSubVector / VecRestoreSubVector (MatMult_Nest, PCFIELDSPLIT, VecNest ...).  The default implementation
   (rvector.c:1699-1707) hands out, for a contiguous index set, a vector of X's type PLACED on X's host array; what the caller writes
   into it -- on the device, if it is a b200 vector -- reaches that host array when the default VecRestoreSubVector resets the placed
   array (our resetarray copies the device values down first).  X's offload mask has to follow: after such a restore the HOST array
   is the valid copy.  The default code does this bookkeeping only for the device types it knows by name (rvector.c:1766-1795),
   so the two ops wrap it.  Non-contiguous index sets take the default's scatter path, which goes through VecScatter on X itself. */
static PetscErrorCode VecGetSubVector_SeqB200(Vec X, IS is, Vec *Y)
{
  PetscErrorCode (*self)(Vec, IS, Vec *) = X->ops->getsubvector;
  PetscErrorCode ierr;
  PetscFunctionBegin;
  X->ops->getsubvector = NULL;
  ierr                 = VecGetSubVector(X, is, Y);
  X->ops->getsubvector = self;
  PetscCall(ierr);
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecRestoreSubVector_SeqB200(Vec X, IS is, Vec *Y)
{
  PetscErrorCode (*self)(Vec, IS, Vec *) = X->ops->restoresubvector;
  PetscErrorCode   ierr;
  PetscObjectState before, after;
  PetscObject      scatter = NULL;
  PetscFunctionBegin;
  PetscCall(PetscObjectQuery((PetscObject)*Y, "VecGetSubVector_Scatter", &scatter));
  PetscCall(PetscObjectStateGet((PetscObject)X, &before));
  X->ops->restoresubvector = NULL;
  ierr                     = VecRestoreSubVector(X, is, Y);
  X->ops->restoresubvector = self;
  PetscCall(ierr);
  PetscCall(PetscObjectStateGet((PetscObject)X, &after));
  if (!scatter && after != before && PB_IsB200(X)) ((Vec_SeqB200 *)X->data)->mask = PB_CPU; /* the sub-vector was written: X's host array holds it */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecGetLocalVector_B200(Vec v, Vec w) { return PB_LocalVectorBegin(v, w, PETSC_TRUE); }
static PetscErrorCode VecRestoreLocalVector_B200(Vec v, Vec w) { return PB_LocalVectorEnd(v, w, PETSC_TRUE); }
static PetscErrorCode VecGetLocalVectorRead_B200(Vec v, Vec w) { return PB_LocalVectorBegin(v, w, PETSC_FALSE); }
static PetscErrorCode VecRestoreLocalVectorRead_B200(Vec v, Vec w) { return PB_LocalVectorEnd(v, w, PETSC_FALSE); }

/* VecDuplicateVecs as ONE device slab with the leading dimension rounded up (VecDuplicateVecs_Seq_GEMV, bvec2.c:670-691:
   the Krylov basis is a strided matrix); the slab is released when its last vector is destroyed */
static PetscErrorCode VecDuplicateVecs_B200(Vec w, PetscInt m, Vec *V[])
{
  struct PB_Slab *slab;
  size_t          lda = ((size_t)w->map->n + 31) & ~(size_t)31; /* 256-byte multiples: 128-bit loads and TMA stay aligned */
  PetscFunctionBegin;
  PetscCheck(m > 0, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "m must be > 0: m = %" PetscInt_FMT, m);
  PetscCall(PetscMalloc1(m, V));
  PetscCall(PetscNew(&slab));
  if (lda) PetscCallB200(b200Malloc(PB_h, (void **)&slab->base, sizeof(double) * lda * (size_t)m));
  for (PetscInt i = 0; i < m; i++) {
    Vec_SeqB200 *b;
    PetscCall(VecDuplicate(w, *V + i));
    b       = (Vec_SeqB200 *)(*V)[i]->data;
    b->d    = slab->base ? slab->base + lda * (size_t)i : NULL;
    b->slab = slab;
    slab->refs++;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecDestroyVecs_B200(PetscInt m, Vec v[])
{
  PetscFunctionBegin;
  for (PetscInt i = 0; i < m; i++) PetscCall(VecDestroy(&v[i]));
  PetscCall(PetscFree(v));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- mpib200: the row-partitioned vector.  Local kernels + NCCL all-reduce of the DEVICE results, then one copy to the host
   (VecXDot_MPI_Default / VecMXDot_MPI_Default / VecNorm_MPI_Default, pvecimpl.h:97-172, with ncclAllReduce for
   MPIU_Allreduce).  The *_local ops stay the sequential ones. */
static double *PB_dred = NULL; /* persistent device scratch for the reductions */
#define PB_DRED_MAX 4096
static PetscErrorCode PB_RedScratch(void)
{
  PetscFunctionBegin;
  if (!PB_dred) PetscCallB200(b200Malloc(PB_h, (void **)&PB_dred, sizeof(double) * PB_DRED_MAX));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_AllreduceHost(double *v, int n, int max) /* host scalars (rare paths: 1-norm, max, sum) */
{
  PetscFunctionBegin;
  if (PB_size == 1 || !n) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCheck(n <= PB_DRED_MAX, PETSC_COMM_SELF, PETSC_ERR_SUP, "too many values");
  PetscCall(PB_RedScratch());
  PetscCallB200(b200MemcpyHtoDAsync(PB_h, PB_dred, v, sizeof(double) * (size_t)n));
  if (max) PetscCallB200(b200CommAllreduceMax(PB_h, PB_dred, n));
  else PetscCallB200(b200CommAllreduceSum(PB_h, PB_dred, n));
  PetscCallB200(b200MemcpyDtoH(PB_h, v, PB_dred, sizeof(double) * (size_t)n));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMDot_MPIB200(Vec x, PetscInt nv, const Vec y[], PetscScalar *z)
{
  const double *dx, *yp[PB_DRED_MAX];
  PetscFunctionBegin;
  PetscCheck(nv <= PB_DRED_MAX, PETSC_COMM_SELF, PETSC_ERR_SUP, "nv too large");
  for (PetscInt j = 0; j < nv; j++) PetscCheck(PB_IsB200(y[j]), PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "mpib200 reductions need b200 vectors");
  PetscCall(PB_RedScratch());
  PetscCall(PB_VecRead(x, &dx));
  for (PetscInt j = 0; j < nv; j++) PetscCall(PB_VecRead(y[j], &yp[j]));
  PetscCallB200(b200VecMDotAsync(PB_h, N_(x), (int)nv, dx, yp, PB_dred));
  PetscCallB200(b200CommAllreduceSum(PB_h, PB_dred, (int)nv));
  PetscCallB200(b200MemcpyDtoH(PB_h, z, PB_dred, sizeof(double) * (size_t)nv));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecDot_MPIB200(Vec x, Vec y, PetscScalar *z) { return VecMDot_MPIB200(x, 1, &y, z); }
static PetscErrorCode VecNorm_MPIB200(Vec x, NormType type, PetscReal *z)
{
  Vec_SeqB200     *b = (Vec_SeqB200 *)x->data;
  PetscObjectState st;
  PetscFunctionBegin;
  if (type == NORM_2 || type == NORM_FROBENIUS) {
    double ss;
    PetscCall(PetscObjectStateGet((PetscObject)x, &st));
    if (b->d_sumsq && b->sumsq_state == st) { /* |x_local|^2 left on the device by the fused MAXPY */
      PetscCallB200(b200CommAllreduceSum(PB_h, b->d_sumsq, 1));
      PetscCallB200(b200MemcpyDtoH(PB_h, &ss, b->d_sumsq, sizeof(double)));
      b->sumsq_state = (PetscObjectState)-1; /* the buffer now holds the global value: never reduce it twice */
    } else PetscCall(VecMDot_MPIB200(x, 1, &x, &ss));
    *z = PetscSqrtReal(ss);
  } else if (type == NORM_1_AND_2) {
    PetscCall(VecNorm_SeqB200(x, NORM_1, &z[0]));
    PetscCall(PB_AllreduceHost(&z[0], 1, 0));
    PetscCall(VecNorm_MPIB200(x, NORM_2, &z[1]));
  } else {
    PetscCall(VecNorm_SeqB200(x, type, z));
    PetscCall(PB_AllreduceHost(z, 1, type == NORM_INFINITY));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecSum_MPIB200(Vec x, PetscScalar *s)
{
  PetscFunctionBegin;
  PetscCall(VecSum_SeqB200(x, s));
  PetscCall(PB_AllreduceHost(s, 1, 0));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMax_MPIB200(Vec x, PetscInt *p, PetscReal *v)
{
  PetscFunctionBegin;
  PetscCheck(!p || PB_size == 1, PETSC_COMM_SELF, PETSC_ERR_SUP, "mpib200: location of the maximum across ranks is not provided");
  PetscCall(VecMax_SeqB200(x, p, v));
  PetscCall(PB_AllreduceHost(v, 1, 1));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode VecMin_MPIB200(Vec x, PetscInt *p, PetscReal *v)
{
  double neg;
  PetscFunctionBegin;
  PetscCheck(!p || PB_size == 1, PETSC_COMM_SELF, PETSC_ERR_SUP, "mpib200: location of the minimum across ranks is not provided");
  PetscCall(VecMin_SeqB200(x, p, v));
  neg = -*v;
  PetscCall(PB_AllreduceHost(&neg, 1, 1));
  *v = -neg;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode VecCreate_SeqB200(Vec v);
static PetscErrorCode       VecDuplicate_SeqB200(Vec win, Vec *V)
{
  PetscFunctionBegin;
  PetscCall(VecCreate(PetscObjectComm((PetscObject)win), V));
  PetscCall(PetscLayoutReference(win->map, &(*V)->map));
  PetscCall(VecSetType(*V, ((PetscObject)win)->type_name)); /* seqb200 or mpib200 */
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
  /* parent = VECSEQ WITHOUT a host array.  VecCreate_Seq_Private(v, NULL) is what we want but it is not exported
     (SURVEY 7 hard part 7); VecCreateSeqWithArray(..., NULL, &tmp) runs it on a temporary, whose implementation (ops table
     + Vec_Seq) is then moved into v.  VecSetType(v, VECSEQ) would allocate and zero n scalars per vector (see PB_VecHostAlloc). */
  {
    Vec tmp;
    PetscCall(PetscLayoutSetUp(v->map));
    PetscCall(VecCreateSeqWithArray(PETSC_COMM_SELF, PetscMax(1, v->map->bs), v->map->n, NULL, &tmp));
    v->ops[0]          = tmp->ops[0];
    s                  = (Vec_Seq *)tmp->data;
    tmp->data          = NULL;
    tmp->ops->destroy  = NULL;
    v->petscnative     = PETSC_TRUE;
    PetscCall(VecDestroy(&tmp));
  }
  if (!PB_VecSeqOpsSet) {
    PB_VecSeqOps    = *v->ops;
    PB_VecSeqOpsSet = PETSC_TRUE;
  }
  /* grow the parent's data structure: Vec_SeqB200 starts with a Vec_Seq */
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
  v->ops->sum                        = VecSum_SeqB200;
  v->ops->max                        = VecMax_SeqB200;
  v->ops->min                        = VecMin_SeqB200;
  v->ops->shift                      = VecShift_SeqB200;
  v->ops->duplicatevecs              = VecDuplicateVecs_B200;
  v->ops->destroyvecs                = VecDestroyVecs_B200;
  v->ops->getsubvector               = VecGetSubVector_SeqB200;
  v->ops->restoresubvector           = VecRestoreSubVector_SeqB200;
  v->ops->getlocalvector             = VecGetLocalVector_B200;
  v->ops->restorelocalvector         = VecRestoreLocalVector_B200;
  v->ops->getlocalvectorread         = VecGetLocalVectorRead_B200;
  v->ops->restorelocalvectorread     = VecRestoreLocalVectorRead_B200;
  PetscCall(PetscObjectChangeTypeName((PetscObject)v, VECSEQB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* mpib200 = seqb200's storage and element-wise kernels + all-reduced reductions.  In this MPIUNI PETSc the object lives on
   a one-rank communicator and its PETSc layout is the LOCAL slice; the ranks are the NCCL ranks of PB_Init. */
PETSC_EXTERN PetscErrorCode VecCreate_MPIB200(Vec v)
{
  PetscFunctionBegin;
  PetscCall(VecCreate_SeqB200(v));
  v->ops->dot   = VecDot_MPIB200;
  v->ops->tdot  = VecDot_MPIB200;
  v->ops->mdot  = VecMDot_MPIB200;
  v->ops->mtdot = VecMDot_MPIB200;
  v->ops->norm  = VecNorm_MPIB200;
  v->ops->sum   = VecSum_MPIB200;
  v->ops->max   = VecMax_MPIB200;
  v->ops->min   = VecMin_MPIB200;
  PetscCall(PetscObjectChangeTypeName((PetscObject)v, VECMPIB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* "b200": by communicator size, as VecCreate_CUDA picks seqcuda/mpicuda (vecreg.c:87-93 family rule) -- here the size of
   the NCCL communicator */
PETSC_EXTERN PetscErrorCode VecCreate_B200(Vec v)
{
  PetscFunctionBegin;
  PetscCall(PB_Init());
  if (PB_size > 1) PetscCall(VecCreate_MPIB200(v));
  else PetscCall(VecCreate_SeqB200(v));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== Mat: seqaijb200 */
typedef struct {
  int             *d_i, *d_j;
  double          *d_a;
  b200CsrPlan      plan;
  int             *d_cr_i, *d_cr_rindex; /* compressed-row view (Mat_CompressedRow, aij.h:152-172) for MatMultAdd on mostly empty blocks */
  PetscInt         cr_nrows;
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

/* The context of a seqaijb200 matrix.  It can be absent: MatHeaderMerge() keeps A's ops but takes the other matrix's data AND spptr
   (gcreate.c:449-486) -- e.g. the in-place MatLUFactor() / MatPermute() of a seqaijb200 matrix end with our ops on top of a plain
   MATSEQAIJ's data.  The context is then rebuilt on first use: the parent's entry points are the same for every MATSEQAIJ (captured
   at the first conversion), the device mirror is simply not there yet (valid = false). */
static Mat_B200  PB_MatParent;
static PetscBool PB_MatParentSet = PETSC_FALSE;
static Mat_B200 *PB_M(Mat A)
{
  if (!A->spptr && PB_MatParentSet) {
    Mat_B200 *m = NULL;
    if (PetscNew(&m) != PETSC_SUCCESS) return NULL;
    *m       = PB_MatParent;
    A->spptr = m;
  }
  return (Mat_B200 *)A->spptr;
}

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
  PetscCallB200(b200Free(PB_h, m->d_cr_i));
  PetscCallB200(b200Free(PB_h, m->d_cr_rindex));
  m->d_i = m->d_j = m->d_cr_i = m->d_cr_rindex = NULL;
  m->d_a      = NULL;
  m->cr_nrows = 0;
  m->valid    = PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* host -> device mirror, keyed on nonzerostate (pattern) and the object state (values): the protocol of
   MatSeqAIJCUSPARSECopyToGPU (aijcusparse.cu:1477-1590) */
/* -mat_b200_spmv_lanes <0|1|2|4|8|16|32>: lanes per row of the SpMV kernels; 1 = the parity mode that reproduces
   MatMult_SeqAIJ's left-to-right row sums bit for bit (default 0 = chosen from the row-length statistics) */
static PetscErrorCode PB_PlanSetFromOptions(Mat A, b200CsrPlan plan)
{
  PetscInt  lanes = 0, nb = -1;
  PetscBool set   = PETSC_FALSE;
  PetscFunctionBegin;
  /* -mat_b200_spmv_column_blocks <n>: column-blocked SpMV passes for gathers that exceed the L2 (0 = off; default -1 = chosen
     from n, the row length and the measured column span: only scattered matrices with n*8 beyond the L2 are blocked) */
  PetscCall(PetscOptionsGetInt(((PetscObject)A)->options, ((PetscObject)A)->prefix, "-mat_b200_spmv_column_blocks", &nb, NULL));
  if (nb < 0) PetscCallB200(b200CsrPlanAutoColumnBlocks(PB_h, plan, NULL));
  else PetscCallB200(b200CsrPlanSetColumnBlocks(PB_h, plan, (int)nb));
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
  Mat_B200        *m = PB_M(A);
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
    PetscCallB200(b200CsrPlanPackValues(PB_h, m->plan, m->d_a)); /* no-op unless the plan is column-blocked */
    PetscCall(PetscLogCpuToGpu((PetscLogDouble)(sizeof(int) * ((size_t)nr + 1 + nz) + sizeof(double) * nz)));
    if (a->compressedrow.use && a->compressedrow.nrows > 0) { /* MatCheckCompressedRow found mostly empty rows (aij.c:1141) */
      const size_t ncr = (size_t)a->compressedrow.nrows;
      PetscCallB200(b200Malloc(PB_h, (void **)&m->d_cr_i, sizeof(int) * (ncr + 1)));
      PetscCallB200(b200Malloc(PB_h, (void **)&m->d_cr_rindex, sizeof(int) * ncr));
      PetscCallB200(b200MemcpyHtoD(PB_h, m->d_cr_i, a->compressedrow.i, sizeof(int) * (ncr + 1)));
      PetscCallB200(b200MemcpyHtoD(PB_h, m->d_cr_rindex, a->compressedrow.rindex, sizeof(int) * ncr));
      m->cr_nrows = a->compressedrow.nrows;
    }
    m->nonzerostate = A->nonzerostate;
    m->valstate     = st;
    m->valid        = PETSC_TRUE;
  } else if (m->valstate != st) {
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_a, a->a, sizeof(double) * nz));
    PetscCallB200(b200CsrPlanPackValues(PB_h, m->plan, m->d_a));
    PetscCall(PetscLogCpuToGpu((PetscLogDouble)(sizeof(double) * nz)));
    m->valstate = st;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the device mirror of A already exists (the matrix was generated or split on the device and copied DOWN to become the host
   master copy): adopt the arrays instead of uploading the same data again */
static PetscErrorCode PB_MatAdoptDevice(Mat A, int *d_i, int *d_j, double *d_a)
{
  Mat_B200        *m = PB_M(A);
  Mat_SeqAIJ      *a = (Mat_SeqAIJ *)A->data;
  PetscObjectState st;
  PetscFunctionBegin;
  PetscCall(PB_MatFreeDevice(m));
  m->d_i = d_i;
  m->d_j = d_j;
  m->d_a = d_a;
  PetscCallB200(b200CsrPlanCreate(PB_h, (int)A->rmap->n, (int)A->cmap->n, (int64_t)a->nz, m->d_i, m->d_j, &m->plan));
  PetscCall(PB_PlanSetFromOptions(A, m->plan));
  PetscCallB200(b200CsrPlanPackValues(PB_h, m->plan, m->d_a));
  if (a->compressedrow.use && a->compressedrow.nrows > 0) {
    const size_t ncr = (size_t)a->compressedrow.nrows;
    PetscCallB200(b200Malloc(PB_h, (void **)&m->d_cr_i, sizeof(int) * (ncr + 1)));
    PetscCallB200(b200Malloc(PB_h, (void **)&m->d_cr_rindex, sizeof(int) * ncr));
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_cr_i, a->compressedrow.i, sizeof(int) * (ncr + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, m->d_cr_rindex, a->compressedrow.rindex, sizeof(int) * ncr));
    m->cr_nrows = a->compressedrow.nrows;
  }
  PetscCall(PetscObjectStateGet((PetscObject)A, &st));
  m->nonzerostate = A->nonzerostate;
  m->valstate     = st;
  m->valid        = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode MatMult_SeqAIJB200(Mat A, Vec x, Vec y)
{
  Mat_B200     *m = PB_M(A);
  Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*m->mult_seqaij)(A, x, y)); /* host vectors: parent */
  PetscCall(PB_MatSync(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(y, &dy));
  PetscCall(PB_LogTimeBegin());
  PetscCallB200(b200CsrSpMV(PB_h, m->plan, m->d_a, dx, dy));
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(2.0 * a->nz - a->nonzerorowcnt)); /* aij.c:1497, booked as device flops (aijcusparse.cu:2565) */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultAdd_SeqAIJB200(Mat A, Vec x, Vec y, Vec z)
{
  Mat_B200     *m = PB_M(A);
  Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
  const double *dx, *dy;
  double       *dz;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y) || !PB_IsB200(z)) PetscFunctionReturn((*m->multadd_seqaij)(A, x, y, z));
  if (!a->nz) { /* no entries (e.g. the off-diagonal block of a one-rank mpiaijb200): z = y */
    if (z != y) PetscCall(VecCopy(y, z));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_MatSync(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_LogTimeBegin());
  if (m->cr_nrows) {
    /* compressed-row branch (aij.c:1626-1640): z = y on the empty rows, then only the rows that own entries are updated */
    if (z != y) PetscCall(VecCopy(y, z));
    PetscCall(PB_VecRW(z, &dz));
    PetscCallB200(b200CsrSpMVAddCompressed(PB_h, (int)m->cr_nrows, m->d_cr_i, m->d_cr_rindex, m->d_j, m->d_a, dx, dz, dz));
  } else {
    PetscCall(PB_VecRead(y, &dy));
    if (z == y) PetscCall(PB_VecRW(z, &dz));
    else PetscCall(PB_VecWrite(z, &dz));
    PetscCallB200(b200CsrSpMVAdd(PB_h, m->plan, m->d_a, dx, dy, dz));
  }
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(2.0 * a->nz));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetDiagonal_SeqAIJB200(Mat A, Vec v)
{
  Mat_B200 *m = PB_M(A);
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
  Mat_B200 *m = PB_M(A);
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
  Mat_B200     *m = PB_M(A);
  const double *dx;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y)) PetscFunctionReturn((*m->multtranspose_seqaij)(A, x, y));
  if (!((Mat_SeqAIJ *)A->data)->nz || !A->cmap->n) {
    PetscCall(VecSet(y, 0.0));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_MatSyncTranspose(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(y, &dy));
  PetscCall(PB_LogTimeBegin());
  PetscCallB200(b200CsrTransposeSpMV(PB_h, m->T, dx, NULL, dy));
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(2.0 * ((Mat_SeqAIJ *)A->data)->nz)); /* aij.c:1427 */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultTransposeAdd_SeqAIJB200(Mat A, Vec x, Vec z, Vec y)
{
  Mat_B200     *m = PB_M(A);
  const double *dx, *dz;
  double       *dy;
  PetscFunctionBegin;
  if (PB_BoundToCPU(A) || !PB_IsB200(x) || !PB_IsB200(y) || !PB_IsB200(z)) PetscFunctionReturn((*m->multtransposeadd_seqaij)(A, x, z, y));
  PetscCall(PB_MatSyncTranspose(A));
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecRead(z, &dz));
  if (y == z) PetscCall(PB_VecRW(y, &dy));
  else PetscCall(PB_VecWrite(y, &dy));
  PetscCall(PB_LogTimeBegin());
  PetscCallB200(b200CsrTransposeSpMV(PB_h, m->T, dx, dz, dy));
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(2.0 * ((Mat_SeqAIJ *)A->data)->nz));
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
  Mat_B200            *m = PB_M(A);
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
  Mat_B200        *m = PB_M(A);
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
  PetscCallB200(b200CsrPlanPackValues(PB_h, m->plan, m->d_a));
  PetscCallB200(b200MemcpyDtoH(PB_h, a->a, m->d_a, sizeof(double) * (size_t)a->nz)); /* host master copy follows */
  PetscCall(PetscObjectStateIncrease((PetscObject)A));
  PetscCall(PetscObjectStateGet((PetscObject)A, &st));
  m->valstate = st; /* device and host agree */
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* Products with a dense matrix (MatMatMult / MatTransposeMatMult / MatMatTransposeMult with MATSEQDENSE): MatProductSetFromOptions
   looks the implementation up under a name built from BOTH type names (matproduct.c:445-471), so a sub-class of MATSEQAIJ has to
   route its own name to the parent's entry, which MATSEQDENSE composes on the dense matrix ("..._seqaij_seqdense_C",
   "..._seqdense_seqaij_C").  The parent's kernels read the host CSR, which is always current (host master copy). */
static PetscErrorCode PB_ProductForward(Mat C, const char *parentname)
{
  Mat_Product *product = C->product;
  PetscErrorCode (*f)(Mat) = NULL;
  PetscFunctionBegin;
  PetscCall(PetscObjectQueryFunction((PetscObject)product->B, parentname, &f));
  if (!f) PetscCall(PetscObjectQueryFunction((PetscObject)product->A, parentname, &f));
  if (f) PetscCall((*f)(C));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatProductSetFromOptions_SeqAIJB200_SeqDense(Mat C)
{
  return PB_ProductForward(C, "MatProductSetFromOptions_seqaij_seqdense_C");
}
static PetscErrorCode MatProductSetFromOptions_SeqDense_SeqAIJB200(Mat C)
{
  return PB_ProductForward(C, "MatProductSetFromOptions_seqdense_seqaij_C");
}

static PetscErrorCode MatDestroy_SeqAIJB200(Mat A)
{
  Mat_B200 *m = PB_M(A);
  PetscErrorCode (*destroy)(Mat) = m->destroy_seqaij;
  PetscFunctionBegin;
  PetscCall(PB_MatFreeDevice(m));
  if (m->coo) b200CooPlanDestroy(m->coo);
  PetscCall(PetscFree(A->spptr));
  PetscCall(PetscObjectComposeFunction((PetscObject)A, "MatConvert_seqaij_seqaijb200_C", NULL));
  PetscCall(PetscObjectComposeFunction((PetscObject)A, "MatProductSetFromOptions_seqaijb200_seqdense_C", NULL));
  PetscCall(PetscObjectComposeFunction((PetscObject)A, "MatProductSetFromOptions_seqdense_seqaijb200_C", NULL));
  PetscCall((*destroy)(A));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode MatConvert_SeqAIJ_SeqAIJB200(Mat A, MatType type, MatReuse reuse, Mat *newmat);
static PetscErrorCode       MatDuplicate_SeqAIJB200(Mat A, MatDuplicateOption op, Mat *B)
{
  Mat_B200 *m = PB_M(A);
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
  if (!PB_MatParentSet) {
    PB_MatParent    = *m; /* entry points only: the device fields are still zero */
    PB_MatParentSet = PETSC_TRUE;
  }
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
  PetscCall(PetscObjectComposeFunction((PetscObject)B, "MatProductSetFromOptions_seqaijb200_seqdense_C", MatProductSetFromOptions_SeqAIJB200_SeqDense));
  PetscCall(PetscObjectComposeFunction((PetscObject)B, "MatProductSetFromOptions_seqdense_seqaijb200_C", MatProductSetFromOptions_SeqDense_SeqAIJB200));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode MatCreate_SeqAIJB200(Mat B)
{
  PetscFunctionBegin;
  PetscCall(MatCreate_SeqAIJ(B)); /* exported creator of the parent (aij.c:4745) */
  PetscCall(MatConvert_SeqAIJ_SeqAIJB200(B, MATSEQAIJB200, MAT_INPLACE_MATRIX, &B));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* MatCreateSeqAIJWithArrays (aij.c:4876) for CSR arrays that live on the DEVICE (b200Malloc'ed; e.g. produced by a device
   assembly or generator): they are copied down once into PETSc-owned host arrays (the host master copy every inherited
   MATSEQAIJ method relies on) and then ADOPTED as the device mirror -- the Mat frees them */
PETSC_EXTERN PetscErrorCode MatCreateSeqAIJB200WithDeviceArrays(PetscInt m, PetscInt n, PetscInt *d_i, PetscInt *d_j, PetscScalar *d_a, Mat *mat)
{
  PetscInt    *hi, *hj, nz = 0;
  PetscScalar *ha;
  Mat_SeqAIJ  *sa;
  PetscFunctionBegin;
  PetscCall(PB_Init());
  PetscCall(PetscMalloc1(m + 1, &hi));
  PetscCallB200(b200MemcpyDtoH(PB_h, hi, d_i, sizeof(int) * ((size_t)m + 1)));
  nz = hi[m];
  PetscCall(PetscMalloc1(nz + 1, &hj));
  PetscCall(PetscMalloc1(nz + 1, &ha));
  PetscCallB200(b200MemcpyDtoH(PB_h, hj, d_j, sizeof(int) * (size_t)nz));
  PetscCallB200(b200MemcpyDtoH(PB_h, ha, d_a, sizeof(double) * (size_t)nz));
  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, n, hi, hj, ha, mat));
  sa         = (Mat_SeqAIJ *)(*mat)->data;
  sa->free_a = sa->free_ij = PETSC_TRUE;
  PetscCall(MatSetType(*mat, MATSEQAIJB200));
  PetscCall(PB_MatAdoptDevice(*mat, d_i, d_j, d_a));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== MatSolverType "b200": ILU(0) on the device */
typedef struct {
  b200IluPlan plan;
  b200IccPlan icc;        /* MAT_FACTOR_ICC */
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
    if (f->icc) PetscCallB200(b200Icc0Destroy(f->icc));
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
  PetscCall(PB_LogTimeBegin());
  PetscCallB200(b200Ilu0Solve(PB_h, f->plan, db, dx)); /* MatSolve_SeqAIJ_NaturalOrdering, aijfact.c:2413 */
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(2.0 * f->nz - F->cmap->n));
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
    d_a = (PB_M(A))->d_a;
  } else {
    Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
    if (!f->d_aval_tmp) PetscCallB200(b200Malloc(PB_h, (void **)&f->d_aval_tmp, sizeof(double) * ((size_t)a->nz + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, f->d_aval_tmp, a->a, sizeof(double) * (size_t)a->nz));
    d_a = f->d_aval_tmp;
  }
  /* -pc_factor_shift_type: NONZERO is the device loop (MatPivotCheck_nz, matimpl.h:795-811); NONE must fail on a zero pivot
     like MatPivotCheck_none; the positive-definite and in-blocks strategies are not provided */
  PetscCheck((MatFactorShiftType)info->shifttype == MAT_SHIFT_NONE || (MatFactorShiftType)info->shifttype == MAT_SHIFT_NONZERO, PetscObjectComm((PetscObject)A), PETSC_ERR_SUP, "MatSolverType b200 supports -pc_factor_shift_type none|nonzero only");
  PetscCallB200(b200Ilu0Numeric(PB_h, f->plan, d_a, info->zeropivot, info->shiftamount > 0 ? info->shiftamount : 100.0 * PETSC_MACHINE_EPSILON, &nshift));
  if ((MatFactorShiftType)info->shifttype == MAT_SHIFT_NONE && nshift) {
    PetscCheck(!F->erroriffailure, PetscObjectComm((PetscObject)A), PETSC_ERR_MAT_LU_ZRPVT, "Zero pivot in ILU(0) (tolerance %g) with -pc_factor_shift_type none", (double)info->zeropivot);
    F->factorerrortype = MAT_FACTOR_NUMERIC_ZEROPIVOT; /* what MatPivotCheck_none records when errors are deferred (matimpl.h:771-787) */
    PetscCall(PetscInfo(F, "Zero pivot in ILU(0): factorisation flagged as failed (-pc_factor_shift_type none)\n"));
  }
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
/* ---- ICC(0): MatICCFactorSymbolic_SeqAIJ / MatCholeskyFactorNumeric_SeqAIJ / MatSolve_SeqSBAIJ_1_NaturalOrdering on the device
   (PCICC is PETSc's default PC for a sequential matrix flagged symmetric, e.g. ex2) */
static PetscErrorCode MatSolve_IccB200(Mat F, Vec b, Vec x)
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
    PetscCall(MatSolve_IccB200(F, tb, tx));
    PetscCall(VecCopy(tx, x));
    PetscCall(VecDestroy(&tb));
    PetscCall(VecDestroy(&tx));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_VecRead(b, &db));
  PetscCall(PB_VecWrite(x, &dx));
  PetscCall(PB_LogTimeBegin());
  PetscCallB200(b200Icc0Solve(PB_h, f->icc, db, dx)); /* sbaijfact2.c:2030-2065 */
  PetscCall(PB_LogTimeEnd());
  PetscCall(PB_LogFlops(4.0 * f->nz - 3.0 * F->rmap->n));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatCholeskyFactorNumeric_FactorB200(Mat F, Mat A, const MatFactorInfo *info)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  const double   *d_a;
  int             bad = 0;
  PetscFunctionBegin;
  if (A->ops->mult == MatMult_SeqAIJB200) {
    PetscCall(PB_MatSync(A));
    d_a = (PB_M(A))->d_a;
  } else {
    Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
    if (!f->d_aval_tmp) PetscCallB200(b200Malloc(PB_h, (void **)&f->d_aval_tmp, sizeof(double) * ((size_t)a->nz + 1)));
    PetscCallB200(b200MemcpyHtoD(PB_h, f->d_aval_tmp, a->a, sizeof(double) * (size_t)a->nz));
    d_a = f->d_aval_tmp;
  }
  PetscCallB200(b200Icc0Numeric(PB_h, f->icc, d_a, info->zeropivot, &bad));
  if (bad) {
    /* the reference's MatPivotCheck_pd would shift the diagonal and refactor (matimpl.h:813-833); this solver reports the
       indefinite pivot instead -- choose -pc_factor_mat_solver_type petsc for such matrices */
    PetscCheck(!F->erroriffailure, PetscObjectComm((PetscObject)A), PETSC_ERR_MAT_CH_ZRPVT, "Zero or negative pivot in ICC(0) at row %d (MatSolverType b200 does not shift)", bad - 1);
    F->factorerrortype = MAT_FACTOR_NUMERIC_ZEROPIVOT;
    PetscCall(PetscInfo(F, "Zero or negative pivot in ICC(0) at row %d: factorisation flagged as failed\n", bad - 1));
  }
  F->ops->solve          = MatSolve_IccB200;
  F->ops->solvetranspose = MatSolve_IccB200; /* symmetric factor */
  F->assembled           = PETSC_TRUE;
  F->preallocated        = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatICCFactorSymbolic_FactorB200(Mat F, Mat A, IS perm, const MatFactorInfo *info)
{
  MatFactor_B200 *f = (MatFactor_B200 *)F->data;
  Mat_SeqAIJ     *a = (Mat_SeqAIJ *)A->data;
  PetscBool       id = PETSC_TRUE;
  int64_t         nzu = 0;
  PetscFunctionBegin;
  PetscCheck(info->levels == 0, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSolverType b200 implements ICC(0) only (got %g levels)", (double)info->levels);
  if (perm) PetscCall(ISIdentity(perm, &id));
  PetscCheck(id, PETSC_COMM_SELF, PETSC_ERR_SUP, "MatSolverType b200 requires the natural ordering (-pc_factor_mat_ordering_type natural)");
  PetscCheck(A->rmap->n == A->cmap->n, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "Must be square matrix, rows %" PetscInt_FMT " columns %" PetscInt_FMT, A->rmap->n, A->cmap->n); /* aijfact.c:2064 */
  if (f->icc) PetscCallB200(b200Icc0Destroy(f->icc));
  f->icc = NULL;
  PetscCallB200(b200Icc0Symbolic(PB_h, (int)A->rmap->n, a->i, a->j, &f->icc));
  PetscCallB200(b200Icc0GetInfo(f->icc, &nzu, NULL, NULL, NULL));
  f->nz                          = (double)nzu;
  F->ops->choleskyfactornumeric  = MatCholeskyFactorNumeric_FactorB200;
  F->info.fill_ratio_given       = info->fill;
  F->info.fill_ratio_needed      = 1.0;
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
  PetscCheck(ftype == MAT_FACTOR_ILU || ftype == MAT_FACTOR_ICC, PetscObjectComm((PetscObject)A), PETSC_ERR_SUP, "MatSolverType b200 provides MAT_FACTOR_ILU and MAT_FACTOR_ICC");
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
  (*F)->ops->iccfactorsymbolic  = MatICCFactorSymbolic_FactorB200;
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

/* ================================================================== Mat: mpiaijb200 (row-partitioned over the NCCL ranks)
   The data structure and MatMult of Mat_MPIAIJ (mpiaij.h:41-76, mpiaij.c:1047-1061): diagonal block A (seqaijb200, local
   columns), off-diagonal block B (seqaijb200, m x ec, columns renumbered into the sorted garray, compressed rows), lvec,
   and the Mvctx scatter -- here a b200Halo (pack kernel + grouped ncclSend/ncclRecv on a second stream, overlapped with the
   diagonal-block SpMV).  garray / B's numbering come from b200MpiaijBuildGarray, the restatement of mmaij.c:25-61.
   With a real MPI under PETSc this would sub-class MATMPIAIJ (mpiaijcusparse.cu:195-247); this PETSc is MPIUNI, so the
   object lives on the process's one-rank communicator with the LOCAL sizes as its PETSc layout, and the global column
   space exists only inside this structure.  Vectors are mpib200 (reductions all-reduced); MatGetDiagonalBlock hands
   PCBJACOBI the sequential block, exactly as MatGetDiagonalBlock_MPIAIJ does. */
typedef struct {
  Mat       A, B;
  PetscInt *garray, ec, N, rstart;
  Vec       lvec;
  b200Halo  Mvctx;
  int64_t  *ranges;
} Mat_MPIAIJB200;

static PetscErrorCode MatMult_MPIAIJB200(Mat mat, Vec x, Vec y)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  const double   *dx;
  double         *dl;
  PetscFunctionBegin;
  PetscCheck(PB_IsB200(x) && PB_IsB200(y), PetscObjectComm((PetscObject)mat), PETSC_ERR_ARG_WRONG, "mpiaijb200 needs b200 vectors");
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(a->lvec, &dl));
  PetscCallB200(b200HaloBegin(PB_h, a->Mvctx, dx, dl)); /* VecScatterBegin (mpiaij.c:1055) */
  PetscCall(MatMult(a->A, x, y));                       /* overlaps the exchange */
  PetscCallB200(b200HaloEnd(PB_h, a->Mvctx));           /* VecScatterEnd */
  PetscCall(MatMultAdd(a->B, a->lvec, y, y));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatMultAdd_MPIAIJB200(Mat mat, Vec x, Vec y, Vec z)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data; /* mpiaij.c:1072-1084 */
  const double   *dx;
  double         *dl;
  PetscFunctionBegin;
  PetscCall(PB_VecRead(x, &dx));
  PetscCall(PB_VecWrite(a->lvec, &dl));
  PetscCallB200(b200HaloBegin(PB_h, a->Mvctx, dx, dl));
  PetscCall(MatMultAdd(a->A, x, y, z));
  PetscCallB200(b200HaloEnd(PB_h, a->Mvctx));
  PetscCall(MatMultAdd(a->B, a->lvec, z, z));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatMultTranspose_MPIAIJ (mpiaij.c:1086-1097): lvec = B^T x, y = A^T x, then the reverse scatter adds lvec into the owners' y */
static PetscErrorCode MatMultTranspose_MPIAIJB200(Mat mat, Vec x, Vec y)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  const double   *dl;
  double         *dy;
  PetscFunctionBegin;
  PetscCall(MatMultTranspose(a->B, x, a->lvec));
  PetscCall(PB_VecRead(a->lvec, &dl));
  PetscCallB200(b200HaloReduceBegin(PB_h, a->Mvctx, dl));
  PetscCall(MatMultTranspose(a->A, x, y));
  PetscCall(PB_VecRW(y, &dy));
  PetscCallB200(b200HaloReduceEnd(PB_h, a->Mvctx, dy));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetDiagonal_MPIAIJB200(Mat mat, Vec v)
{
  PetscFunctionBegin;
  PetscCall(MatGetDiagonal(((Mat_MPIAIJB200 *)mat->data)->A, v)); /* mpiaij.c:1158 */
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatGetDiagonalBlock_MPIAIJB200(Mat mat, Mat *blk)
{
  PetscFunctionBegin;
  *blk = ((Mat_MPIAIJB200 *)mat->data)->A;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatDestroy_MPIAIJB200(Mat mat)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  PetscFunctionBegin;
  if (a) {
    PetscCall(MatDestroy(&a->A));
    PetscCall(MatDestroy(&a->B));
    PetscCall(VecDestroy(&a->lvec));
    if (a->Mvctx) PetscCallB200(b200HaloDestroy(a->Mvctx));
    if (a->garray) PetscCallB200(b200HostFree(a->garray));
    PetscCall(PetscFree(a->ranges));
    PetscCall(PetscFree(mat->data));
  }
  PetscCall(PetscObjectComposeFunction((PetscObject)mat, "MatMPIAIJB200GetSeqAIJ_C", NULL));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode MatView_MPIAIJB200(Mat mat, PetscViewer viewer)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  PetscBool       ascii;
  PetscFunctionBegin;
  PetscCall(PetscObjectTypeCompare((PetscObject)viewer, PETSCVIEWERASCII, &ascii));
  if (ascii) PetscCall(PetscViewerASCIIPrintf(viewer, "mpiaijb200: NCCL rank %d of %d, rows [%" PetscInt_FMT ",%" PetscInt_FMT ") of %" PetscInt_FMT ", %" PetscInt_FMT " ghost columns\n", PB_rank, PB_size, a->rstart, a->rstart + mat->rmap->n, a->N, a->ec));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatMPIAIJGetSeqAIJ analogue (mpiaij.c:5800): the two blocks and garray */
static PetscErrorCode MatMPIAIJB200GetSeqAIJ_MPIAIJB200(Mat mat, Mat *Ad, Mat *Ao, const PetscInt **colmap)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  PetscFunctionBegin;
  if (Ad) *Ad = a->A;
  if (Ao) *Ao = a->B;
  if (colmap) *colmap = a->garray;
  PetscFunctionReturn(PETSC_SUCCESS);
}
PETSC_EXTERN PetscErrorCode MatMPIAIJB200GetSeqAIJ(Mat mat, Mat *Ad, Mat *Ao, const PetscInt **colmap)
{
  PetscFunctionBegin;
  PetscUseMethod(mat, "MatMPIAIJB200GetSeqAIJ_C", (Mat, Mat *, Mat *, const PetscInt **), (mat, Ad, Ao, colmap));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the shell of the type: blocks are attached by one of the creators below */
static PetscErrorCode PB_MatCreateMPIAIJB200Shell(PetscInt m, PetscInt N, PetscInt rstart, Mat *mat, Mat_MPIAIJB200 **data)
{
  Mat             M;
  Mat_MPIAIJB200 *a;
  PetscFunctionBegin;
  PetscCall(PB_Init());
  PetscCall(MatCreate(PETSC_COMM_SELF, &M));
  PetscCall(MatSetSizes(M, m, m, m, m)); /* the PETSc layout of this process is its local slice (see the header comment) */
  PetscCall(PetscLayoutSetUp(M->rmap));
  PetscCall(PetscLayoutSetUp(M->cmap));
  PetscCall(PetscNew(&a));
  a->N      = N;
  a->rstart = rstart;
  M->data   = a;
  M->ops->mult             = MatMult_MPIAIJB200;
  M->ops->multadd          = MatMultAdd_MPIAIJB200;
  M->ops->multtranspose    = MatMultTranspose_MPIAIJB200;
  M->ops->getdiagonal      = MatGetDiagonal_MPIAIJB200;
  M->ops->getdiagonalblock = MatGetDiagonalBlock_MPIAIJB200;
  M->ops->destroy          = MatDestroy_MPIAIJB200;
  M->ops->view             = MatView_MPIAIJB200;
  M->assembled             = PETSC_TRUE;
  M->preallocated          = PETSC_TRUE;
  PetscCall(PetscFree(M->defaultvectype));
  PetscCall(PetscStrallocpy(PB_size > 1 ? VECMPIB200 : VECSEQB200, &M->defaultvectype));
  PetscCall(PetscObjectChangeTypeName((PetscObject)M, MATMPIAIJB200));
  PetscCall(PetscObjectComposeFunction((PetscObject)M, "MatMPIAIJB200GetSeqAIJ_C", MatMPIAIJB200GetSeqAIJ_MPIAIJB200));
  *mat  = M;
  *data = a;
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatSetUpMultiply_MPIAIJ (mmaij.c:8-126): garray, B renumbered and shrunk to ec columns, lvec, the scatter.
   oj holds GLOBAL columns on entry and is renumbered in place.  copy: duplicate the off-diagonal arrays (the caller's are
   not adopted). */
static PetscErrorCode PB_MatSetUpMultiply_MPIAIJB200(Mat mat, PetscInt *oi, PetscInt *oj, PetscScalar *oa, PetscBool copy)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  const PetscInt  m = mat->rmap->n;
  int             ec = 0;
  PetscFunctionBegin;
  for (PetscInt k = 0; k < oi[m]; k++) PetscCheck(oj[k] >= 0 && oj[k] < a->N, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Column %" PetscInt_FMT " out of range [0,%" PetscInt_FMT ")", oj[k], a->N);
  PetscCallB200(b200MpiaijBuildGarray((int64_t)oi[m], oj, &a->garray, &ec));
  a->ec = ec;
  if (copy) {
    PetscCall(MatCreate(PETSC_COMM_SELF, &a->B));
    PetscCall(MatSetSizes(a->B, m, ec, m, ec));
    PetscCall(MatSetType(a->B, MATSEQAIJB200));
    PetscCall(MatSeqAIJSetPreallocationCSR(a->B, oi, oj, oa));
  } else {
    PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, ec, oi, oj, oa, &a->B));
    PetscCall(MatSetType(a->B, MATSEQAIJB200));
  }
  PetscCall(VecCreateSeq(PETSC_COMM_SELF, ec, &a->lvec)); /* mmaij.c:103 */
  PetscCall(VecSetType(a->lvec, VECSEQB200));
  PetscCall(PetscMalloc1(PB_size + 1, &a->ranges));
  PetscCallB200(b200HaloCreateFromGarray(PB_h, (int)m, ec, a->garray, a->ranges, &a->Mvctx));
  PetscCheck(a->ranges[PB_rank] == a->rstart && a->ranges[PB_size] == a->N, PETSC_COMM_SELF, PETSC_ERR_ARG_INCOMP, "row ranges of the ranks do not tile [0,N): rank %d starts at %" PetscInt_FMT ", expected %lld; N %" PetscInt_FMT " vs %lld", PB_rank, a->rstart, (long long)a->ranges[PB_rank], a->N, (long long)a->ranges[PB_size]);
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* MatCreateMPIAIJWithSplitArrays (mpiaij.c:6977) for mpiaijb200: the caller's HOST arrays are adopted without a copy --
   diagonal block (i,j,a) with LOCAL columns, off-diagonal block (oi,oj,oa) with GLOBAL columns (renumbered in place, as
   the reference does).  This process owns rows [rstart, rstart+m) of the N x N operator.  Collective over the NCCL ranks. */
PETSC_EXTERN PetscErrorCode MatCreateMPIAIJB200WithSplitArrays(PetscInt m, PetscInt N, PetscInt rstart, PetscInt i[], PetscInt j[], PetscScalar a[], PetscInt oi[], PetscInt oj[], PetscScalar oa[], Mat *mat)
{
  Mat_MPIAIJB200 *d;
  PetscFunctionBegin;
  PetscCall(PB_MatCreateMPIAIJB200Shell(m, N, rstart, mat, &d));
  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, m, i, j, a, &d->A));
  PetscCall(MatSetType(d->A, MATSEQAIJB200));
  PetscCall(PB_MatSetUpMultiply_MPIAIJB200(*mat, oi, oj, oa, PETSC_FALSE));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatMPIAIJSetPreallocationCSR / MatCreateMPIAIJWithArrays (mpiaij.c:4130,4330) for mpiaijb200: local rows with GLOBAL
   columns, host or device pointers (detected); the arrays are copied.  Device input is split on the device
   (b200CsrSplitColumns) and the diagonal block is copied down once to become the host master copy of its seqaijb200. */
PETSC_EXTERN PetscErrorCode MatCreateMPIAIJB200WithArrays(PetscInt m, PetscInt N, PetscInt rstart, const PetscInt i[], const PetscInt j[], const PetscScalar a[], Mat *mat)
{
  Mat_MPIAIJB200 *d;
  int             dev = 0;
  PetscInt       *Ai, *Aj, *Bi, *Bj;
  PetscScalar    *Aa, *Ba;
  int64_t         nzA = 0, nzB = 0;
  int            *adopt_i = NULL, *adopt_j = NULL;
  double         *adopt_a = NULL;
  PetscFunctionBegin;
  PetscCall(PB_MatCreateMPIAIJB200Shell(m, N, rstart, mat, &d));
  PetscCallB200(b200PointerIsDevice(i, &dev));
  if (dev) {
    int    *dAi, *dAj, *dBi, *dBj;
    double *dAa, *dBa;
    PetscCallB200(b200CsrSplitColumns(PB_h, (int)m, i, j, a, (int)rstart, (int)(rstart + m), &dAi, &dAj, &dAa, &nzA, &dBi, &dBj, &dBa, &nzB));
    PetscCall(PetscMalloc1(m + 1, &Ai));
    PetscCall(PetscMalloc1(nzA + 1, &Aj));
    PetscCall(PetscMalloc1(nzA + 1, &Aa));
    PetscCall(PetscMalloc1(m + 1, &Bi));
    PetscCall(PetscMalloc1(nzB + 1, &Bj));
    PetscCall(PetscMalloc1(nzB + 1, &Ba));
    PetscCallB200(b200MemcpyDtoH(PB_h, Ai, dAi, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyDtoH(PB_h, Aj, dAj, sizeof(int) * (size_t)nzA));
    PetscCallB200(b200MemcpyDtoH(PB_h, Aa, dAa, sizeof(double) * (size_t)nzA));
    PetscCallB200(b200MemcpyDtoH(PB_h, Bi, dBi, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyDtoH(PB_h, Bj, dBj, sizeof(int) * (size_t)nzB));
    PetscCallB200(b200MemcpyDtoH(PB_h, Ba, dBa, sizeof(double) * (size_t)nzB));
    PetscCallB200(b200Free(PB_h, dBi)); PetscCallB200(b200Free(PB_h, dBj)); PetscCallB200(b200Free(PB_h, dBa));
    adopt_i = dAi; adopt_j = dAj; adopt_a = dAa;
  } else {
    PetscCallB200(b200MpiaijSplitHost((int)m, (int)rstart, (int)(rstart + m), i, j, a, &nzA, &nzB, NULL, NULL, NULL, NULL, NULL, NULL));
    PetscCall(PetscMalloc1(m + 1, &Ai));
    PetscCall(PetscMalloc1(nzA + 1, &Aj));
    PetscCall(PetscMalloc1(nzA + 1, &Aa));
    PetscCall(PetscMalloc1(m + 1, &Bi));
    PetscCall(PetscMalloc1(nzB + 1, &Bj));
    PetscCall(PetscMalloc1(nzB + 1, &Ba));
    PetscCallB200(b200MpiaijSplitHost((int)m, (int)rstart, (int)(rstart + m), i, j, a, &nzA, &nzB, Ai, Aj, Aa, Bi, Bj, Ba));
  }
  /* the diagonal block adopts the split arrays (freed with the Mat, like MatSeqAIJSetPreallocationCSR's copies) */
  PetscCall(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, m, m, Ai, Aj, Aa, &d->A));
  {
    Mat_SeqAIJ *sa = (Mat_SeqAIJ *)d->A->data;
    sa->free_a = sa->free_ij = PETSC_TRUE; /* aij.h:60-61: PETSc now owns Ai/Aj/Aa */
  }
  PetscCall(MatSetType(d->A, MATSEQAIJB200));
  if (adopt_i) PetscCall(PB_MatAdoptDevice(d->A, adopt_i, adopt_j, adopt_a));
  PetscCall(PB_MatSetUpMultiply_MPIAIJB200(*mat, Bi, Bj, Ba, PETSC_TRUE));
  PetscCall(PetscFree(Bi));
  PetscCall(PetscFree(Bj));
  PetscCall(PetscFree(Ba));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* MatCreate + MatSetType("mpiaijb200"|"aijb200" on >1 NCCL ranks) gives an empty shell that only these creators can fill */
PETSC_EXTERN PetscErrorCode MatCreate_MPIAIJB200(Mat B)
{
  PetscFunctionBegin;
  SETERRQ(PetscObjectComm((PetscObject)B), PETSC_ERR_SUP, "mpiaijb200 matrices are built with MatCreateMPIAIJB200WithArrays() / MatCreateMPIAIJB200WithSplitArrays() (this PETSc is MPIUNI: MatSetValues() cannot route off-process entries)");
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== PC "jacobib200" (and, by default, "jacobi" itself):
   PCJACOBI with a fused ops->applyBA.  KSPGMRESCycle reaches the operator through KSP_PCApplyBAorAB -> PCApplyBAorAB, which
   calls ops->applyBA when the PC has one (precon.c:848-849): with left preconditioning y = D^-1 (A x) is then ONE kernel
   (b200CsrSpMVJacobi; on mpiaijb200 the diagonal-block kernel + a halo-row epilogue) instead of MatMult +
   VecPointwiseMult -- 24 B/row less traffic, same rounding (row sum first, then one multiply).
   This is a sub-class of the reference's own PCJACOBI (PCCreate_Jacobi is exported): every option, type (rowmax, rowsum,
   rowl1), symmetric application, view ... is the parent's; only applyBA is added, and it fuses only the plain diagonal
   variant on b200 matrices, falling back to the generic composition of precon.c:850-862 otherwise.  The fused diagonal is
   computed as PCSetUp_Jacobi does (jacobi.c:172-270): MatGetDiagonal, reciprocal, zero diagonal -> 1.0.
   The plugin re-registers the name "jacobi" with this creator, so -pc_type jacobi is fused with no change to the command
   line; -b200_keep_pcjacobi leaves the stock type alone (then -pc_type jacobib200 selects this one). */
PETSC_EXTERN PetscErrorCode PCCreate_Jacobi(PC);
typedef struct {
  Vec              dinv;
  PetscObjectState matstate;
  Mat              mat;
  PetscBool        fuse, usable;
  PetscErrorCode (*apply_parent)(PC, Vec, Vec);
  PetscErrorCode (*getdiagonal_parent)(PC, Vec, Vec);
} PC_JacobiB200;

static PetscErrorCode PB_JacobiCtx(PC pc, PC_JacobiB200 **jac)
{
  PetscContainer c;
  PetscFunctionBegin;
  PetscCall(PetscObjectQuery((PetscObject)pc, "PCJacobiB200_ctx", (PetscObject *)&c));
  PetscCheck(c, PetscObjectComm((PetscObject)pc), PETSC_ERR_PLIB, "PCJACOBIB200 context missing");
  PetscCall(PetscContainerGetPointer(c, (void **)jac));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_JacobiCtxDestroy(PetscCtxRt ctx)
{
  PC_JacobiB200 *jac = *(PC_JacobiB200 **)ctx;
  PetscFunctionBegin;
  PetscCall(VecDestroy(&jac->dinv));
  PetscCall(PetscFree(jac));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* (re)build the fused diagonal when the operator changed (object state): PCSetUp's own refresh rule (precon.c:1080-1110) */
static PetscErrorCode PB_JacobiRefresh(PC pc, PC_JacobiB200 *jac)
{
  PetscObjectState st;
  PCJacobiType     type;
  PetscBool        useabs, b200mat;
  PetscInt         n;
  PetscFunctionBegin;
  PetscCall(PetscObjectStateGet((PetscObject)pc->pmat, &st));
  if (jac->mat == pc->pmat && jac->matstate == st && jac->dinv) PetscFunctionReturn(PETSC_SUCCESS);
  jac->mat      = pc->pmat;
  jac->matstate = st;
  jac->usable   = PETSC_FALSE;
  PetscCall(PCJacobiGetType(pc, &type));
  PetscCall(PCJacobiGetUseAbs(pc, &useabs));
  PetscCall(PetscObjectTypeCompareAny((PetscObject)pc->mat, &b200mat, MATSEQAIJB200, MATMPIAIJB200, ""));
  if (!jac->fuse || type != PC_JACOBI_DIAGONAL || useabs || !b200mat || pc->mat != pc->pmat) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(PB_Init());
  if (!jac->dinv) PetscCall(MatCreateVecs(pc->pmat, &jac->dinv, NULL));
  if (!PB_IsB200(jac->dinv)) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(MatGetDiagonal(pc->pmat, jac->dinv));
  PetscCall(VecGetLocalSize(jac->dinv, &n));
  {
    double *d;
    int     nzero = 0;
    PetscCall(PB_VecRW(jac->dinv, &d));
    PetscCallB200(b200JacobiInvertDiagonal(PB_h, (int64_t)n, d, d, &nzero));
    /* a zero diagonal: the parent decides between 1.0 and an error/inf depending on the SPD flag (jacobi.c:253-266) --
       leave that case to the parent's own apply */
    if (nzero) PetscFunctionReturn(PETSC_SUCCESS);
  }
  jac->usable = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* PCApply: with the device diagonal at hand it is one VecPointwiseMult (jacobi.c:354).  The parent's lazy set-up
   (PCSetUp_Jacobi_NonSymmetric -> PCSetUp_Jacobi, jacobi.c:172-270) scans the diagonal for zeros in a HOST loop -- a device ->
   host -> device round trip of the vector plus n host iterations -- which the device inversion already did (nzero == 0). */
static PetscErrorCode PCApply_JacobiB200(PC pc, Vec x, Vec y)
{
  PC_JacobiB200 *jac;
  PetscFunctionBegin;
  PetscCall(PB_JacobiCtx(pc, &jac));
  PetscCall(PB_JacobiRefresh(pc, jac));
  if (jac->usable && PB_IsB200(x) && PB_IsB200(y)) PetscCall(VecPointwiseMult(y, x, jac->dinv));
  else PetscCall((*jac->apply_parent)(pc, x, y));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PCApplyBA_JacobiB200(PC pc, PCSide side, Vec x, Vec y, Vec work)
{
  PC_JacobiB200 *jac;
  Mat            A = pc->mat;
  PetscFunctionBegin;
  PetscCall(PB_JacobiCtx(pc, &jac));
  if (side == PC_LEFT) PetscCall(PB_JacobiRefresh(pc, jac));
  if (side == PC_LEFT && jac->usable && PB_IsB200(x) && PB_IsB200(y) && A->ops->mult == MatMult_SeqAIJB200 && !PB_BoundToCPU(A)) {
    Mat_B200     *m = PB_M(A);
    Mat_SeqAIJ   *a = (Mat_SeqAIJ *)A->data;
    const double *dx, *dd;
    double       *dy;
    PetscCall(PB_MatSync(A));
    PetscCall(PB_VecRead(x, &dx));
    PetscCall(PB_VecRead(jac->dinv, &dd));
    PetscCall(PB_VecWrite(y, &dy));
    PetscCall(PB_LogTimeBegin());
    PetscCallB200(b200CsrSpMVJacobi(PB_h, m->plan, m->d_a, dx, dd, dy, NULL));
    PetscCall(PB_LogTimeEnd());
    PetscCall(PB_LogFlops(2.0 * a->nz - a->nonzerorowcnt + A->rmap->n));
  } else if (side == PC_LEFT && jac->usable && PB_IsB200(x) && PB_IsB200(y) && A->ops->mult == MatMult_MPIAIJB200) {
    /* mpiaij.c:1047-1061 + jacobi.c:354 fused: diagonal block writes w = dinv.*(A_d x) while the halo travels; the rows
       that own off-diagonal entries are then redone with both blocks in the reference's order */
    Mat_MPIAIJB200 *mp = (Mat_MPIAIJB200 *)A->data;
    Mat_B200       *mA = PB_M(mp->A), *mB = PB_M(mp->B);
    Mat_SeqAIJ     *sA = (Mat_SeqAIJ *)mp->A->data, *sB = (Mat_SeqAIJ *)mp->B->data;
    const double   *dx, *dd, *dlr;
    double         *dy, *dl;
    PetscCall(PB_MatSync(mp->A));
    if (sB->nz) PetscCall(PB_MatSync(mp->B));
    if (sB->nz && !mB->cr_nrows) { /* off-diagonal block not in compressed-row form: unfused */
      PetscCall(MatMult(A, x, work));
      PetscCall(VecPointwiseMult(y, work, jac->dinv));
      PetscFunctionReturn(PETSC_SUCCESS);
    }
    PetscCall(PB_VecRead(x, &dx));
    PetscCall(PB_VecRead(jac->dinv, &dd));
    PetscCall(PB_VecWrite(mp->lvec, &dl));
    PetscCall(PB_VecWrite(y, &dy));
    PetscCall(PB_LogTimeBegin());
    PetscCallB200(b200HaloBegin(PB_h, mp->Mvctx, dx, dl));
    PetscCallB200(b200CsrSpMVJacobi(PB_h, mA->plan, mA->d_a, dx, dd, dy, NULL));
    PetscCallB200(b200HaloEnd(PB_h, mp->Mvctx));
    if (sB->nz) {
      PetscCall(PB_VecRead(mp->lvec, &dlr));
      PetscCallB200(b200CsrSpMVAddCompressedJacobi(PB_h, (int)mB->cr_nrows, mB->d_cr_i, mB->d_cr_rindex, mB->d_j, mB->d_a, dlr, mA->d_i, mA->d_j, mA->d_a, dx, dd, dy));
    }
    PetscCall(PB_LogTimeEnd());
    PetscCall(PB_LogFlops(2.0 * sA->nz + 2.0 * sB->nz - sA->nonzerorowcnt + A->rmap->n));
  } else if (side == PC_LEFT) { /* the generic composition of PCApplyBAorAB (precon.c:850-862) */
    PetscCall(MatMult(A, x, work));
    PetscCall(PCApply(pc, work, y));
  } else if (side == PC_RIGHT) {
    PetscCall(PCApply(pc, x, work));
    PetscCall(MatMult(A, work, y));
  } else {
    PetscCall(PCApplySymmetricRight(pc, x, work));
    PetscCall(MatMult(A, work, y));
    PetscCall(VecCopy(y, work));
    PetscCall(PCApplySymmetricLeft(pc, work, y));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* PCJacobiGetDiagonal(): the parent copies out the (inverted) diagonal its own apply created lazily (jacobi.c:144-157: "Use PCApply
   to force creation"); when the sub-class applied the preconditioner from its device copy the parent never made one */
static PetscErrorCode PCJacobiGetDiagonal_JacobiB200(PC pc, Vec diag, Vec diagsqrt)
{
  PC_JacobiB200 *jac;
  PetscFunctionBegin;
  PetscCall(PB_JacobiCtx(pc, &jac));
  if (diag && !diagsqrt && jac->usable && jac->dinv) {
    PetscCall(PB_JacobiRefresh(pc, jac));
    if (jac->usable) {
      PetscCall(VecCopy(jac->dinv, diag));
      PetscFunctionReturn(PETSC_SUCCESS);
    }
  }
  PetscCheck(jac->getdiagonal_parent, PetscObjectComm((PetscObject)pc), PETSC_ERR_SUP, "PCJacobiGetDiagonal is not available");
  PetscCall((*jac->getdiagonal_parent)(pc, diag, diagsqrt));
  PetscFunctionReturn(PETSC_SUCCESS);
}
PETSC_EXTERN PetscErrorCode PCCreate_JacobiB200(PC pc)
{
  PC_JacobiB200 *jac;
  PetscContainer c;
  PetscFunctionBegin;
  PetscCall(PCCreate_Jacobi(pc)); /* the reference's PCJACOBI: data, options, setup, apply, view, destroy */
  PetscCall(PetscNew(&jac));
  jac->fuse = PETSC_TRUE;
  PetscCall(PetscOptionsGetBool(((PetscObject)pc)->options, ((PetscObject)pc)->prefix, "-pc_jacobi_b200_fuse", &jac->fuse, NULL));
  PetscCall(PetscContainerCreate(PetscObjectComm((PetscObject)pc), &c));
  PetscCall(PetscContainerSetPointer(c, jac));
  PetscCall(PetscContainerSetCtxDestroy(c, PB_JacobiCtxDestroy));
  PetscCall(PetscObjectCompose((PetscObject)pc, "PCJacobiB200_ctx", (PetscObject)c));
  PetscCall(PetscContainerDestroy(&c));
  PetscCall(PetscObjectQueryFunction((PetscObject)pc, "PCJacobiGetDiagonal_C", &jac->getdiagonal_parent));
  PetscCall(PetscObjectComposeFunction((PetscObject)pc, "PCJacobiGetDiagonal_C", PCJacobiGetDiagonal_JacobiB200));
  jac->apply_parent       = pc->ops->apply;
  pc->ops->apply          = PCApply_JacobiB200;
  pc->ops->applytranspose = PCApply_JacobiB200;
  pc->ops->applyBA        = PCApplyBA_JacobiB200;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== KSP "pipecgb200": single-reduction CG with fused recurrences
   (a fused-reduction KSP registered through KSPRegister).  The method is Ghysels & Vanroose's pipelined CG, the
   one behind the reference's KSPPIPECG (src/ksp/ksp/impls/cg/pipecg/pipecg.c); what this type adds on the device is
     * ONE reduction kernel per iteration -- (u,u), (r,u), (w,u) as a 3-vector VecMDot: one kernel, one host synchronisation
       and, on several GPUs, one ncclAllReduce of three numbers, where KSPCG needs three of each (cg.c:220-330);
     * ONE kernel for the eight vector recurrences of an iteration (b200VecPipeCGUpdate: z,q,p,s <- n,m,u,w + beta(z,q,p,s);
       x,u,w,r <- ... -/+ alpha(p,q,z,s)): 10 vector reads + 8 writes instead of the 24 + 8 of eight separate AYPX/AXPY kernels.
   Per-entry arithmetic is that of the separate VecAYPX/VecAXPY calls (tested bit for bit against them), so the iterates are those
   of KSPPIPECG: the residual histories of the reference's -ksp_type pipecg fixtures are reproduced to 1e-12 * r0.
   Preconditioned residual norm, left preconditioning, zero or non-zero initial guess. */
typedef struct {
  Vec r, u, w, m, n, z, q, p, s; /* work vectors (KSPSetWorkVecs) */
  PetscBool fuse;
} PipeCGB200_Vecs;

static PetscErrorCode KSPSetUp_PipeCGB200(KSP ksp)
{
  PetscFunctionBegin;
  PetscCall(KSPSetWorkVecs(ksp, 9));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* the eight recurrences: one kernel when every vector is a device vector, else the eight BLAS-1 calls */
static PetscErrorCode PipeCGB200_Advance(KSP ksp, PipeCGB200_Vecs *v, PetscScalar alpha, PetscScalar beta, PetscBool first)
{
  Vec       x = ksp->vec_sol, all[10] = {v->n, v->m, v->u, v->w, v->z, v->q, v->p, v->s, x, v->r};
  PetscBool dev = v->fuse;
  PetscFunctionBegin;
  for (int k = 0; k < 10 && dev; k++) dev = PB_IsB200(all[k]) ? PETSC_TRUE : PETSC_FALSE;
  if (dev) {
    const double *dn, *dm;
    double       *d[8];
    PetscCall(PB_VecRead(v->n, &dn));
    PetscCall(PB_VecRead(v->m, &dm));
    for (int k = 0; k < 8; k++) {
      if (first && k >= 2 && k < 6) PetscCall(PB_VecWrite(all[2 + k], &d[k])); /* z, q, p, s are overwritten in the first step */
      else PetscCall(PB_VecRW(all[2 + k], &d[k]));
      PetscCall(PetscObjectStateIncrease((PetscObject)all[2 + k]));
    }
    PetscCall(PB_LogTimeBegin());
    PetscCallB200(b200VecPipeCGUpdate(PB_h, N_(x), alpha, beta, first ? 1 : 0, dn, dm, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]));
    PetscCall(PB_LogTimeEnd());
    PetscCall(PB_LogFlops((first ? 8.0 : 16.0) * x->map->n));
  } else {
    if (first) {
      PetscCall(VecCopy(v->n, v->z));
      PetscCall(VecCopy(v->m, v->q));
      PetscCall(VecCopy(v->u, v->p));
      PetscCall(VecCopy(v->w, v->s));
    } else {
      PetscCall(VecAYPX(v->z, beta, v->n));
      PetscCall(VecAYPX(v->q, beta, v->m));
      PetscCall(VecAYPX(v->p, beta, v->u));
      PetscCall(VecAYPX(v->s, beta, v->w));
    }
    PetscCall(VecAXPY(x, alpha, v->p));
    PetscCall(VecAXPY(v->u, -alpha, v->q));
    PetscCall(VecAXPY(v->w, -alpha, v->z));
    PetscCall(VecAXPY(v->r, -alpha, v->s));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* residual norm bookkeeping shared by the start-up and the loop: history, monitors, convergence test */
static PetscErrorCode PipeCGB200_Check(KSP ksp, PetscInt it, PetscReal rnorm)
{
  PetscFunctionBegin;
  KSPCheckNorm(ksp, rnorm);
  ksp->rnorm = rnorm;
  PetscCall(KSPLogResidualHistory(ksp, rnorm));
  PetscCall(KSPMonitor(ksp, it, rnorm));
  PetscCall((*ksp->converged)(ksp, it, rnorm, &ksp->reason, ksp->cnvP));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode KSPSolve_PipeCGB200(KSP ksp)
{
  PipeCGB200_Vecs v;
  Mat             A, P;
  Vec             b = ksp->vec_rhs, x = ksp->vec_sol, trio[3];
  PetscScalar     dots[3], alpha = 0.0, beta = 0.0, gamma_prev = 0.0;
  PetscReal       rnorm;

  PetscFunctionBegin;
  v.r = ksp->work[0]; v.z = ksp->work[1]; v.p = ksp->work[2]; v.n = ksp->work[3]; v.w = ksp->work[4];
  v.q = ksp->work[5]; v.u = ksp->work[6]; v.m = ksp->work[7]; v.s = ksp->work[8];
  v.fuse = PETSC_TRUE;
  PetscCall(PetscOptionsGetBool(((PetscObject)ksp)->options, ((PetscObject)ksp)->prefix, "-ksp_pipecgb200_fuse_update", &v.fuse, NULL));
  PetscCall(PCGetOperators(ksp->pc, &A, &P));
  ksp->its = 0;
  /* r = b - A x,  u = B r,  w = A u */
  if (ksp->guess_zero) PetscCall(VecCopy(b, v.r));
  else {
    PetscCall(KSP_MatMult(ksp, A, x, v.r));
    PetscCall(VecAYPX(v.r, -1.0, b));
  }
  PetscCall(KSP_PCApply(ksp, v.r, v.u));
  PetscCall(KSP_MatMult(ksp, A, v.u, v.w));
  PetscCall(VecNorm(v.u, NORM_2, &rnorm));
  PetscCall(PipeCGB200_Check(ksp, 0, rnorm));
  trio[0] = v.u; trio[1] = v.r; trio[2] = v.w;
  for (PetscInt it = 0; !ksp->reason; it++) {
    if (it >= ksp->max_it) {
      ksp->reason = KSP_DIVERGED_ITS;
      break;
    }
    /* m = B w and n = A m do not depend on the reduction: queued first, so the device keeps working while the host waits */
    PetscCall(KSP_PCApply(ksp, v.w, v.m));
    PetscCall(KSP_MatMult(ksp, A, v.m, v.n));
    PetscCall(VecMDot(v.u, 3, trio, dots)); /* (u,u), (r,u), (w,u) */
    if (it > 0) {
      PetscCall(PipeCGB200_Check(ksp, it, PetscSqrtReal(PetscAbsScalar(dots[0]))));
      if (ksp->reason) break;
    }
    {
      const PetscScalar gamma = dots[1], delta = dots[2];
      if (it == 0) alpha = gamma / delta;
      else {
        beta  = gamma / gamma_prev;
        alpha = gamma / (delta - beta / alpha * gamma);
      }
      gamma_prev = gamma;
    }
    PetscCall(PipeCGB200_Advance(ksp, &v, alpha, beta, it == 0 ? PETSC_TRUE : PETSC_FALSE));
    ksp->its = it + 1;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
PETSC_EXTERN PetscErrorCode KSPCreate_PipeCGB200(KSP ksp)
{
  PetscFunctionBegin;
  PetscCall(KSPSetSupportedNorm(ksp, KSP_NORM_PRECONDITIONED, PC_LEFT, 3));
  PetscCall(KSPSetSupportedNorm(ksp, KSP_NORM_NONE, PC_LEFT, 1));
  ksp->ops->setup          = KSPSetUp_PipeCGB200;
  ksp->ops->solve          = KSPSolve_PipeCGB200;
  ksp->ops->destroy        = KSPDestroyDefault;
  ksp->ops->view           = NULL;
  ksp->ops->setfromoptions = NULL;
  ksp->ops->buildsolution  = KSPBuildSolutionDefault;
  ksp->ops->buildresidual  = KSPBuildResidualDefault;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== KSP "pgmresb200": pipelined GMRES, one reduction per iteration,
   read one iteration after it was launched (the GMRES half).  The method is Ghysels, Ashby, Meerbergen & Vanroose's
   p(1)-GMRES, the one behind the reference's KSPPGMRES (src/ksp/ksp/impls/gmres/pgmres/pgmres.c): the iterates and the residual
   history are those of -ksp_type pgmres (reference-run fixtures, 1e-12 * r0).  The reference overlaps an MPI reduction with the
   next matrix product through VecMDotBegin/End; on one GPU its split phase computes the local part synchronously, so nothing
   overlaps.  Here the reduction of iteration j -- ONE VecMDot kernel writing into mapped pinned memory, the norm fused into the
   VecMAXPY before it -- is launched, an event recorded, and the host reads the results in iteration j+1 AFTER it has queued that
   iteration's fused B*A product: the host waits for the event only, the device keeps working on the product.  One host wait per
   iteration where KSPGMRES needs two (after VecMDot, after VecNorm).  Host vectors / several ranks: the same algorithm through the
   public Vec interface (every reduction then synchronises where it is issued).
   Left preconditioning, preconditioned residual norm, restart = -ksp_gmres_restart (30). */
typedef struct {
  PetscInt     max_k;
  Vec         *vv; /* max_k + 3 basis vectors; work[0], work[1] = temp, tmat */
  PetscScalar *hh, *hes, *grs, *cc, *ss, *nrs, *work;
  double      *h_red, *d_red; /* mapped pinned: results of the pending VecMDot (max_k + 3 doubles) */
  b200Event    ev;
  PetscBool    async;
  /* what is pending from the previous iteration */
  PetscInt     pend_mdot; /* number of dot products in flight (0 = none); they are column pend_col of H */
  PetscInt     pend_col;
  Vec          pend_norm_vec; /* the vector whose fused |.|^2 is in flight (NULL = none) */
  PetscReal    pend_norm;     /* generic path: the value itself */
} KSP_PGMRESB200;
#define PG_HH(a, b)  (pg->hh + (size_t)(b) * (pg->max_k + 2) + (a))
#define PG_HES(a, b) (pg->hes + (size_t)(b) * (pg->max_k + 2) + (a))

static PetscErrorCode KSPSetUp_PGMRESB200(KSP ksp)
{
  KSP_PGMRESB200 *pg = (KSP_PGMRESB200 *)ksp->data;
  const size_t    hs = (size_t)(pg->max_k + 2) * (pg->max_k + 2);
  PetscFunctionBegin;
  if (pg->vv) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(KSPSetWorkVecs(ksp, 2));
  PetscCall(KSPCreateVecs(ksp, pg->max_k + 3, &pg->vv, 0, NULL));
  PetscCall(PetscCalloc7(hs, &pg->hh, hs, &pg->hes, pg->max_k + 3, &pg->grs, pg->max_k + 3, &pg->cc, pg->max_k + 3, &pg->ss, pg->max_k + 3, &pg->nrs, pg->max_k + 3, &pg->work));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode KSPReset_PGMRESB200(KSP ksp)
{
  KSP_PGMRESB200 *pg = (KSP_PGMRESB200 *)ksp->data;
  PetscFunctionBegin;
  if (pg->vv) PetscCall(VecDestroyVecs(pg->max_k + 3, &pg->vv));
  PetscCall(PetscFree7(pg->hh, pg->hes, pg->grs, pg->cc, pg->ss, pg->nrs, pg->work));
  if (pg->h_red) PetscCallB200(b200FreeHost(pg->h_red));
  pg->h_red = pg->d_red = NULL;
  if (pg->ev) PetscCallB200(b200EventDestroy(pg->ev));
  pg->ev = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode KSPDestroy_PGMRESB200(KSP ksp)
{
  PetscFunctionBegin;
  PetscCall(KSPReset_PGMRESB200(ksp));
  PetscCall(KSPDestroyDefault(ksp));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode KSPSetFromOptions_PGMRESB200(KSP ksp, PetscOptionItems PetscOptionsObject)
{
  KSP_PGMRESB200 *pg = (KSP_PGMRESB200 *)ksp->data;
  PetscInt        restart = pg->max_k;
  PetscBool       flg;
  PetscFunctionBegin;
  PetscOptionsHeadBegin(PetscOptionsObject, "KSP pgmresb200 options");
  PetscCall(PetscOptionsInt("-ksp_gmres_restart", "Number of Krylov search directions", "KSPGMRESSetRestart", restart, &restart, &flg));
  if (flg) {
    PetscCheck(restart >= 1, PetscObjectComm((PetscObject)ksp), PETSC_ERR_ARG_OUTOFRANGE, "restart must be positive");
    if (restart != pg->max_k) {
      PetscCall(KSPReset_PGMRESB200(ksp));
      pg->max_k        = restart;
      ksp->setupstage  = KSP_SETUP_NEW;
    }
  }
  PetscCall(PetscOptionsBool("-ksp_pgmresb200_async", "Launch the reduction and read it one iteration later (device vectors, one rank)", "", pg->async, &pg->async, NULL));
  PetscOptionsHeadEnd();
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* "VecMDotBegin(z, nv, VV, H(:,col))" */
static PetscErrorCode PG_MDotBegin(KSP_PGMRESB200 *pg, PetscBool dev, Vec z, PetscInt nv, PetscInt col)
{
  PetscFunctionBegin;
  if (dev) {
    const double  *dz;
    const double **yp;
    PetscCall(PB_VecRead(z, &dz));
    PetscCall(PetscMalloc1(nv, &yp));
    for (PetscInt j = 0; j < nv; j++) PetscCall(PB_VecRead(pg->vv[j], &yp[j]));
    PetscCallB200(b200VecMDotAsync(PB_h, N_(z), (int)nv, dz, yp, pg->d_red));
    PetscCallB200(b200EventRecord(PB_h, pg->ev)); /* everything the next iteration has to wait for is before this point */
    PetscCall(PetscFree(yp));
    PetscCall(PetscLogFlops(2.0 * nv * z->map->n));
    pg->pend_mdot = nv;
    pg->pend_col  = col;
  } else {
    PetscCall(VecMDot(z, nv, pg->vv, PG_HH(0, col)));
    pg->pend_mdot = 0;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* "VecNormBegin(v)" right after the VecMAXPY that produced v: on the device the squared norm is already in flight (fused) */
static PetscErrorCode PG_NormBegin(KSP_PGMRESB200 *pg, PetscBool dev, Vec v)
{
  PetscFunctionBegin;
  pg->pend_norm_vec = NULL;
  if (dev) {
    Vec_SeqB200     *b = (Vec_SeqB200 *)v->data;
    PetscObjectState st;
    PetscCall(PetscObjectStateGet((PetscObject)v, &st));
    if (b->h_sumsq && b->sumsq_state == st) {
      pg->pend_norm_vec = v;
      PetscFunctionReturn(PETSC_SUCCESS);
    }
  }
  PetscCall(VecNorm(v, NORM_2, &pg->pend_norm));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* the End of both: one wait for the event recorded after the VecMDot launch */
static PetscErrorCode PG_ReductionsEnd(KSP_PGMRESB200 *pg)
{
  PetscFunctionBegin;
  if (pg->pend_mdot || pg->pend_norm_vec) PetscCallB200(b200EventSynchronize(pg->ev));
  if (pg->pend_mdot) {
    for (PetscInt j = 0; j < pg->pend_mdot; j++) *PG_HH(j, pg->pend_col) = ((volatile double *)pg->h_red)[j];
    pg->pend_mdot = 0;
  }
  if (pg->pend_norm_vec) {
    pg->pend_norm     = PetscSqrtReal(*(volatile double *)((Vec_SeqB200 *)pg->pend_norm_vec->data)->h_sumsq);
    pg->pend_norm_vec = NULL;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PG_Log(KSP ksp, PetscReal rnorm)
{
  PetscFunctionBegin;
  PetscCall(KSPLogResidualHistory(ksp, rnorm));
  PetscCall(KSPMonitor(ksp, ksp->its, rnorm));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode KSPSolve_PGMRESB200(KSP ksp)
{
  KSP_PGMRESB200 *pg = (KSP_PGMRESB200 *)ksp->data;
  const PetscInt  max_k = pg->max_k;
  Vec             x = ksp->vec_sol, b = ksp->vec_rhs, temp = ksp->work[0], tmat = ksp->work[1], *vv = pg->vv;
  Mat             A, P;
  PetscBool       dev = pg->async, guess_zero = ksp->guess_zero;
  PetscInt        itcount = 0;
  PetscReal       rnorm = -1.0;

  PetscFunctionBegin;
  PetscCall(PCGetOperators(ksp->pc, &A, &P));
  if (PB_size > 1) dev = PETSC_FALSE; /* several ranks: the reductions are all-reduced where they are issued */
  for (PetscInt j = 0; j < max_k + 3 && dev; j++) dev = PB_IsB200(vv[j]) ? PETSC_TRUE : PETSC_FALSE;
  if (dev && !pg->h_red) {
    PetscCall(PB_Init());
    PetscCallB200(b200MallocMapped((void **)&pg->h_red, (void **)&pg->d_red, sizeof(double) * (size_t)(max_k + 3)));
    PetscCallB200(b200EventCreate(&pg->ev));
  }
  ksp->its    = 0;
  ksp->reason = KSP_CONVERGED_ITERATING;
  while (!ksp->reason) {
    PetscInt  it     = 0;
    PetscBool hapend = PETSC_FALSE;
    PetscReal resn;
    /* KSPInitialResidual, left preconditioning (itres.c) */
    if (!guess_zero) {
      PetscCall(KSP_MatMult(ksp, A, x, temp));
      PetscCall(VecCopy(b, tmat));
      PetscCall(VecAXPY(tmat, -1.0, temp));
      PetscCall(KSP_PCApply(ksp, tmat, vv[0]));
    } else PetscCall(KSP_PCApply(ksp, b, vv[0]));
    /* ---- one cycle (pgmres.c:17-174) ---- */
    pg->pend_mdot     = 0;
    pg->pend_norm_vec = NULL;
    PetscCall(VecNorm(vv[0], NORM_2, &resn));
    KSPCheckNorm(ksp, resn);
    if (resn != 0.0) PetscCall(VecScale(vv[0], 1.0 / resn));
    pg->grs[0] = resn;
    rnorm      = resn;
    ksp->rnorm = rnorm;
    PetscCall(PG_Log(ksp, rnorm));
    if (!resn) {
      ksp->reason = KSP_CONVERGED_ATOL;
      break;
    }
    PetscCall((*ksp->converged)(ksp, ksp->its, rnorm, &ksp->reason, ksp->cnvP));
    for (; !ksp->reason; it++) {
      Vec Zcur = vv[it], Znext = vv[it + 1];
      if (it < max_k + 1 && ksp->its + 1 < PetscMax(2, ksp->max_it)) PetscCall(KSP_PCApplyBAorAB(ksp, Zcur, Znext, tmat)); /* Znext <- B A Zcur, queued BEFORE the wait */
      PetscCall(PG_ReductionsEnd(pg));
      if (it > 1) *PG_HH(it - 1, it - 2) = pg->pend_norm;
      if (it > 1) {
        PetscCall(VecScale(vv[it - 1], 1.0 / *PG_HH(it - 1, it - 2)));
        { /* Hessenberg update of column it-2: copy, apply the previous rotations, new rotation (pgmres.c:226-301) */
          const PetscInt c = it - 2;
          PetscScalar   *h = PG_HH(0, c);
          PetscReal      hapbnd;
          for (PetscInt j = 0; j <= c + 1; j++) *PG_HES(j, c) = h[j];
          hapbnd = PetscAbsScalar(h[c + 1] / pg->grs[c]);
          if (hapbnd > 1.0e-30) hapbnd = 1.0e-30; /* ksp->haptol-like bound of the reference (gmres haptol default) */
          if (PetscAbsScalar(h[c + 1]) < hapbnd) hapend = PETSC_TRUE;
          for (PetscInt j = 0; j < c; j++) {
            const PetscScalar hhj = h[j];
            h[j]     = pg->cc[j] * hhj + pg->ss[j] * h[j + 1];
            h[j + 1] = -pg->ss[j] * hhj + pg->cc[j] * h[j + 1];
          }
          if (!hapend) {
            const PetscReal delta = PetscSqrtReal(h[c] * h[c] + h[c + 1] * h[c + 1]);
            if (delta == 0.0) {
              ksp->reason = KSP_DIVERGED_NULL;
              break;
            }
            pg->cc[c]      = h[c] / delta;
            pg->ss[c]      = h[c + 1] / delta;
            h[c]           = pg->cc[c] * h[c] + pg->ss[c] * h[c + 1];
            pg->grs[c + 1] = -pg->ss[c] * pg->grs[c];
            pg->grs[c]     = pg->cc[c] * pg->grs[c];
            resn           = PetscAbsScalar(pg->grs[c + 1]);
          } else resn = 0.0;
        }
        ksp->its++;
        rnorm      = resn;
        ksp->rnorm = rnorm;
        PetscCall((*ksp->converged)(ksp, ksp->its, rnorm, &ksp->reason, ksp->cnvP));
        if (ksp->reason) break;
        if (it < max_k + 1) PetscCall(PG_Log(ksp, rnorm));
        if (hapend) {
          ksp->reason = KSP_DIVERGED_BREAKDOWN;
          break;
        }
        if (!(it < max_k + 1 && ksp->its < ksp->max_it)) break;
        {
          const PetscScalar sc = *PG_HH(it - 1, it - 2);
          PetscCall(VecScale(Zcur, 1.0 / sc));
          PetscCall(VecScale(Znext, 1.0 / sc));
          for (PetscInt k = 0; k < it; k++) *PG_HH(k, it - 1) /= sc;
          *PG_HH(it - 1, it - 1) /= sc;
        }
      }
      if (it > 0) {
        for (PetscInt k = 0; k < it + 1; k++) {
          pg->work[k] = 0;
          for (PetscInt j = PetscMax(k - 1, 0); j < it - 1; j++) pg->work[k] -= *PG_HES(k, j) * *PG_HH(j, it - 1);
        }
        PetscCall(VecMAXPY(Znext, it + 1, pg->work, vv));
        PetscCall(VecAXPY(Znext, -*PG_HH(it - 1, it - 1), Zcur));
        for (PetscInt k = 0; k < it; k++) pg->work[k] = -*PG_HH(k, it - 1);
        PetscCall(VecMAXPY(Zcur, it, pg->work, vv));
        PetscCall(PG_NormBegin(pg, dev, vv[it]));
      }
      PetscCall(PG_MDotBegin(pg, dev, Znext, it + 1, it));
    }
    PetscCall(PG_ReductionsEnd(pg)); /* nothing may stay in flight across a restart */
    itcount += PetscMax(it - 1, 0);
    { /* the correction: back substitution on the rotated Hessenberg, x += VV * y (pgmres.c:176-206) */
      const PetscInt k = it - 2;
      if (k >= 0) {
        pg->nrs[k] = (*PG_HH(k, k) != 0.0) ? pg->grs[k] / *PG_HH(k, k) : 0.0;
        for (PetscInt kk = k - 1; kk >= 0; kk--) {
          PetscScalar tt = pg->grs[kk];
          for (PetscInt j = kk + 1; j <= k; j++) tt -= *PG_HH(kk, j) * pg->nrs[j];
          pg->nrs[kk] = tt / *PG_HH(kk, kk);
        }
        PetscCall(VecSet(temp, 0.0));
        PetscCall(VecMAXPY(temp, k + 1, pg->nrs, vv));
        PetscCall(VecAXPY(x, 1.0, temp));
      }
    }
    if (!ksp->reason && ksp->its == ksp->max_it) ksp->reason = KSP_DIVERGED_ITS;
    if (ksp->reason) PetscCall(PG_Log(ksp, rnorm));
    if (itcount >= ksp->max_it) {
      if (!ksp->reason) ksp->reason = KSP_DIVERGED_ITS;
      break;
    }
    guess_zero = PETSC_FALSE;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}
PETSC_EXTERN PetscErrorCode KSPCreate_PGMRESB200(KSP ksp)
{
  KSP_PGMRESB200 *pg;
  PetscFunctionBegin;
  PetscCall(PetscNew(&pg));
  pg->max_k = 30; /* GMRES_DEFAULT_MAXK (gmresimpl.h) */
  pg->async = PETSC_TRUE;
  ksp->data = (void *)pg;
  PetscCall(KSPSetSupportedNorm(ksp, KSP_NORM_PRECONDITIONED, PC_LEFT, 3));
  ksp->ops->setup          = KSPSetUp_PGMRESB200;
  ksp->ops->solve          = KSPSolve_PGMRESB200;
  ksp->ops->reset          = KSPReset_PGMRESB200;
  ksp->ops->destroy        = KSPDestroy_PGMRESB200;
  ksp->ops->view           = NULL;
  ksp->ops->setfromoptions = KSPSetFromOptions_PGMRESB200;
  ksp->ops->buildsolution  = KSPBuildSolutionDefault;
  ksp->ops->buildresidual  = KSPBuildResidualDefault;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ================================================================== PetscSF "b200": PETSCSFBASIC sub-classed for device data
   (SURVEY 8f.4).  VecScatterBegin hands the SF whatever VecGetArray[Read]AndMemType returns (vscat.c:50-51,70-73): for b200 vectors
   that is a device pointer tagged PETSC_MEMTYPE_CUDA.  A PETSc configured with a device back end runs its d_ScatterAnd<Op> kernels
   on such pointers (sfpack.c:759-775); a host-only PETSc has none and would dereference them on the host.  This type keeps every
   host-memory operation with the parent (PETSCSFBASIC, sfbasic.c:607) and runs the device ones itself:
     * roots and leaves both on the device, graph local to the process, unit = n x PetscScalar or n x PetscInt, op one of
       REPLACE/SUM/PROD/MAX/MIN: one b200IndexedOp kernel in Begin (leaf <- leaf op root for a broadcast, root <- root op leaf for a
       reduction), entries sharing a destination applied in graph order like PetscSFLinkScatterLocal (sfpack.c:1082) -- results
       equal the reference's host results bit for bit;
     * anything else that involves device memory (mixed host/device, other units or ops, an in-place scatter, a graph that spans
       MPI ranks): the device side is staged through a host buffer and the PARENT does the operation -- correct, not fast.
   Registered as "b200" and, unless -b200_keep_sfbasic, under the name "basic" too so that every VecScatter of a program that
   runs with -vec_type b200 is covered without an -sf_type option. */
#include <petsc/private/sfimpl.h>
#define PETSCSFB200 "b200"
#define PB_SF_MAXINFLIGHT 8
static PetscErrorCode (*PB_SFCreate_Basic)(PetscSF) = NULL; /* the reference's creator, looked up at registration */

typedef struct {
  PetscErrorCode (*bcastbegin)(PetscSF, MPI_Datatype, PetscMemType, const void *, PetscMemType, void *, MPI_Op);
  PetscErrorCode (*bcastend)(PetscSF, MPI_Datatype, const void *, void *, MPI_Op);
  PetscErrorCode (*reducebegin)(PetscSF, MPI_Datatype, PetscMemType, const void *, PetscMemType, void *, MPI_Op);
  PetscErrorCode (*reduceend)(PetscSF, MPI_Datatype, const void *, void *, MPI_Op);
  PetscErrorCode (*fetchbegin)(PetscSF, MPI_Datatype, PetscMemType, void *, PetscMemType, const void *, void *, MPI_Op);
  PetscErrorCode (*reset)(PetscSF);
  b200IndexedPlan bcast, reduce; /* root -> leaf, leaf -> root; built on first device use of the current graph */
  PetscBool       planned, local;
  PetscInt        nroots, leafextent;
  int             ninflight;
  struct {
    const void *root, *leaf;
  } inflight[PB_SF_MAXINFLIGHT]; /* Begin calls this type completed itself: the matching End is a no-op */
  PetscLogDouble nnative, nstaged; /* operations run by the device kernel / staged through the host (PetscSFView) */
} SF_B200;

static PetscErrorCode PB_SFCtx(PetscSF sf, SF_B200 **b)
{
  PetscContainer c;
  PetscFunctionBegin;
  PetscCall(PetscObjectQuery((PetscObject)sf, "PetscSFB200_ctx", (PetscObject *)&c));
  PetscCheck(c, PetscObjectComm((PetscObject)sf), PETSC_ERR_PLIB, "PetscSF of type b200 without its context");
  PetscCall(PetscContainerGetPointer(c, (void **)b));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_SFDropPlans(SF_B200 *b)
{
  PetscFunctionBegin;
  if (b->bcast) PetscCallB200(b200IndexedPlanDestroy(PB_h, b->bcast));
  if (b->reduce) PetscCallB200(b200IndexedPlanDestroy(PB_h, b->reduce));
  b->bcast = b->reduce = NULL;
  b->planned   = PETSC_FALSE;
  b->ninflight = 0;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PB_SFCtxDestroy(PetscCtxRt ctx)
{
  SF_B200 *b = *(SF_B200 **)ctx;
  PetscFunctionBegin;
  PetscCall(PB_SFDropPlans(b));
  PetscCall(PetscFree(b));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PetscSFReset_B200(PetscSF sf) /* PetscSFSetGraph / PetscSFDestroy: the graph is about to change */
{
  SF_B200 *b;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  PetscCall(PB_SFDropPlans(b));
  if (b->reset) PetscCall((*b->reset)(sf));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* extents of the two data buffers and, for a process-local graph, the two indexed plans */
static PetscErrorCode PB_SFPlan(PetscSF sf, SF_B200 *b)
{
  PetscInt           nroots, nleaves, minleaf, maxleaf;
  const PetscInt    *ilocal;
  const PetscSFNode *iremote;
  PetscMPIInt        size;
  PetscFunctionBegin;
  if (b->planned) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(PetscSFGetGraph(sf, &nroots, &nleaves, &ilocal, &iremote));
  PetscCheck(nroots >= 0, PetscObjectComm((PetscObject)sf), PETSC_ERR_ARG_WRONGSTATE, "PetscSF graph has not been set");
  PetscCall(PetscSFGetLeafRange(sf, &minleaf, &maxleaf));
  PetscCallMPI(MPI_Comm_size(PetscObjectComm((PetscObject)sf), &size));
  b->nroots     = nroots;
  b->leafextent = nleaves ? maxleaf + 1 : 0;
  b->local      = (size == 1) ? PETSC_TRUE : PETSC_FALSE;
  if (b->local) {
    int *ridx;
    PetscCall(PetscMalloc1(nleaves + 1, &ridx));
    for (PetscInt k = 0; k < nleaves; k++) {
      PetscCheck(iremote[k].rank == 0 && iremote[k].index >= 0 && iremote[k].index < nroots, PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "leaf %" PetscInt_FMT " points at root (%" PetscInt_FMT ",%" PetscInt_FMT ") outside this process", k, (PetscInt)iremote[k].rank, iremote[k].index);
      ridx[k] = (int)iremote[k].index;
    }
    PetscCall(PB_Init());
    PetscCallB200(b200IndexedPlanCreate(PB_h, (int64_t)nleaves, ridx, 0, (const int *)ilocal, 0, &b->bcast));  /* leaf[ilocal[k]] <- root[ridx[k]] */
    PetscCallB200(b200IndexedPlanCreate(PB_h, (int64_t)nleaves, (const int *)ilocal, 0, ridx, 0, &b->reduce)); /* root[ridx[k]] <- leaf[ilocal[k]] */
    PetscCall(PetscFree(ridx));
  }
  b->planned = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* unit and op as the kernels know them; *native = PETSC_FALSE when they do not */
static PetscErrorCode PB_SFDecode(MPI_Datatype unit, MPI_Op op, int *dtype, int *bs, int *sfop, size_t *unitbytes, PetscBool *native)
{
  PetscMPIInt sz;
  PetscFunctionBegin;
  *native = PETSC_TRUE;
  *dtype  = -1;
  *bs     = 1;
  PetscCallMPI(MPI_Type_size(unit, &sz));
  *unitbytes = (size_t)sz;
#if defined(PETSC_HAVE_MPIUNI)
  { /* MPIUNI packs a datatype as [combiner:4 | type-index:8 | count:12 | base-bytes:8] (include/petsc/mpiuni/mpi.h:205) */
    const int idx = ((int)unit >> 20) & 0xff, cnt = ((int)unit >> 8) & 0xfff, esz = (int)unit & 0xff;
    if (idx == (((int)MPI_DOUBLE >> 20) & 0xff) && esz == (int)sizeof(double)) *dtype = B200_SF_F64;
    else if (idx == (((int)MPI_INT >> 20) & 0xff) && esz == (int)sizeof(int)) *dtype = B200_SF_I32;
    *bs = cnt;
  }
#else
  if (unit == MPIU_SCALAR || unit == MPIU_REAL) *dtype = B200_SF_F64;
  else if (unit == MPIU_INT) *dtype = B200_SF_I32;
#endif
  if (op == MPI_REPLACE) *sfop = B200_SF_REPLACE;
  else if (op == MPI_SUM || op == MPIU_SUM) *sfop = B200_SF_SUM;
  else if (op == MPI_PROD) *sfop = B200_SF_PROD;
  else if (op == MPI_MAX || op == MPIU_MAX) *sfop = B200_SF_MAX;
  else if (op == MPI_MIN || op == MPIU_MIN) *sfop = B200_SF_MIN;
  else *native = PETSC_FALSE;
  if (*dtype < 0 || *bs < 1) *native = PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PB_SFPush(SF_B200 *b, const void *root, const void *leaf)
{
  PetscFunctionBegin;
  PetscCheck(b->ninflight < PB_SF_MAXINFLIGHT, PETSC_COMM_SELF, PETSC_ERR_SUP, "more than %d PetscSF operations in flight on device data", PB_SF_MAXINFLIGHT);
  b->inflight[b->ninflight].root = root;
  b->inflight[b->ninflight].leaf = leaf;
  b->ninflight++;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscBool PB_SFPop(SF_B200 *b, const void *root, const void *leaf)
{
  for (int i = 0; i < b->ninflight; i++)
    if (b->inflight[i].root == root && b->inflight[i].leaf == leaf) {
      b->inflight[i] = b->inflight[--b->ninflight];
      return PETSC_TRUE;
    }
  return PETSC_FALSE;
}

/* direction 0: broadcast (src = root, dst = leaf); 1: reduction (src = leaf, dst = root) */
static PetscErrorCode PB_SFBegin(PetscSF sf, int direction, MPI_Datatype unit, PetscMemType rmtype, const void *rootdata, PetscMemType lmtype, const void *leafdata, MPI_Op op)
{
  SF_B200    *b;
  const PetscBool rdev = PetscMemTypeDevice(rmtype) ? PETSC_TRUE : PETSC_FALSE, ldev = PetscMemTypeDevice(lmtype) ? PETSC_TRUE : PETSC_FALSE;
  int         dtype, bs, sfop;
  size_t      ub;
  PetscBool   native;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  if (!rdev && !ldev) { /* host data: the parent, untouched */
    if (direction == 0) PetscCall((*b->bcastbegin)(sf, unit, rmtype, rootdata, lmtype, (void *)leafdata, op));
    else PetscCall((*b->reducebegin)(sf, unit, lmtype, leafdata, rmtype, (void *)rootdata, op));
    PetscFunctionReturn(PETSC_SUCCESS);
  }
  PetscCall(PB_Init());
  PetscCall(PB_SFPlan(sf, b));
  PetscCall(PB_SFDecode(unit, op, &dtype, &bs, &sfop, &ub, &native));
  if (rdev && ldev && native && b->local && rootdata != leafdata) {
    PetscCall(PB_LogTimeBegin());
    if (direction == 0) PetscCallB200(b200IndexedOp(PB_h, b->bcast, dtype, bs, sfop, rootdata, (void *)leafdata));
    else PetscCallB200(b200IndexedOp(PB_h, b->reduce, dtype, bs, sfop, leafdata, (void *)rootdata));
    PetscCall(PB_LogTimeEnd());
    b->nnative += 1;
  } else { /* stage the device side(s) through the host and let the parent do the operation there */
    const size_t rbytes = (size_t)b->nroots * ub, lbytes = (size_t)b->leafextent * ub;
    const PetscBool same = (rootdata == leafdata) ? PETSC_TRUE : PETSC_FALSE;
    char *hroot = (char *)rootdata, *hleaf = (char *)leafdata;
    if (rdev) {
      PetscCall(PetscMalloc1(PetscMax(rbytes, same ? lbytes : 0) + 1, &hroot));
      if (PetscMax(rbytes, same ? lbytes : 0)) PetscCallB200(b200MemcpyDtoH(PB_h, hroot, rootdata, PetscMax(rbytes, same ? lbytes : 0)));
    }
    if (same) hleaf = hroot;
    else if (ldev) {
      PetscCall(PetscMalloc1(lbytes + 1, &hleaf));
      if (lbytes) PetscCallB200(b200MemcpyDtoH(PB_h, hleaf, leafdata, lbytes));
    }
    if (direction == 0) {
      PetscCall((*b->bcastbegin)(sf, unit, PETSC_MEMTYPE_HOST, hroot, PETSC_MEMTYPE_HOST, hleaf, op));
      PetscCall((*b->bcastend)(sf, unit, hroot, hleaf, op));
      if (ldev && lbytes) PetscCallB200(b200MemcpyHtoD(PB_h, (void *)leafdata, hleaf, lbytes));
    } else {
      PetscCall((*b->reducebegin)(sf, unit, PETSC_MEMTYPE_HOST, hleaf, PETSC_MEMTYPE_HOST, hroot, op));
      PetscCall((*b->reduceend)(sf, unit, hleaf, hroot, op));
      if (rdev && rbytes) PetscCallB200(b200MemcpyHtoD(PB_h, (void *)rootdata, hroot, rbytes));
    }
    if (rdev) PetscCall(PetscFree(hroot));
    if (ldev && !same) PetscCall(PetscFree(hleaf));
    b->nstaged += 1;
  }
  PetscCall(PB_SFPush(b, rootdata, leafdata));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PetscSFBcastBegin_B200(PetscSF sf, MPI_Datatype unit, PetscMemType rmtype, const void *rootdata, PetscMemType lmtype, void *leafdata, MPI_Op op)
{
  return PB_SFBegin(sf, 0, unit, rmtype, rootdata, lmtype, leafdata, op);
}
static PetscErrorCode PetscSFReduceBegin_B200(PetscSF sf, MPI_Datatype unit, PetscMemType lmtype, const void *leafdata, PetscMemType rmtype, void *rootdata, MPI_Op op)
{
  return PB_SFBegin(sf, 1, unit, rmtype, rootdata, lmtype, leafdata, op);
}
static PetscErrorCode PetscSFBcastEnd_B200(PetscSF sf, MPI_Datatype unit, const void *rootdata, void *leafdata, MPI_Op op)
{
  SF_B200 *b;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  if (!PB_SFPop(b, rootdata, leafdata)) PetscCall((*b->bcastend)(sf, unit, rootdata, leafdata, op));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PetscSFReduceEnd_B200(PetscSF sf, MPI_Datatype unit, const void *leafdata, void *rootdata, MPI_Op op)
{
  SF_B200 *b;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  if (!PB_SFPop(b, rootdata, leafdata)) PetscCall((*b->reduceend)(sf, unit, leafdata, rootdata, op));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode PetscSFFetchAndOpBegin_B200(PetscSF sf, MPI_Datatype unit, PetscMemType rmtype, void *rootdata, PetscMemType lmtype, const void *leafdata, void *leafupdate, MPI_Op op)
{
  SF_B200 *b;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  PetscCheck(!PetscMemTypeDevice(rmtype) && !PetscMemTypeDevice(lmtype), PetscObjectComm((PetscObject)sf), PETSC_ERR_SUP, "PetscSFFetchAndOp on device data is not supported by PetscSF type b200");
  PetscCall((*b->fetchbegin)(sf, unit, rmtype, rootdata, lmtype, leafdata, leafupdate, op));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* for tests and PETSc programs that link the plugin: how many Begin calls of this SF ran on the device / were staged */
PETSC_EXTERN PetscErrorCode PetscSFB200GetCounts(PetscSF sf, PetscInt *native, PetscInt *staged)
{
  SF_B200 *b;
  PetscFunctionBegin;
  PetscCall(PB_SFCtx(sf, &b));
  if (native) *native = (PetscInt)b->nnative;
  if (staged) *staged = (PetscInt)b->nstaged;
  PetscFunctionReturn(PETSC_SUCCESS);
}

PETSC_EXTERN PetscErrorCode PetscSFCreate_B200(PetscSF sf)
{
  SF_B200       *b;
  PetscContainer c;
  PetscFunctionBegin;
  PetscCheck(PB_SFCreate_Basic, PetscObjectComm((PetscObject)sf), PETSC_ERR_ORDER, "PetscSF type b200 used before the plugin was registered");
  PetscCall((*PB_SFCreate_Basic)(sf)); /* the reference's PETSCSFBASIC: data, setup, host pack/unpack, view, destroy */
  PetscCall(PetscNew(&b));
  b->bcastbegin  = sf->ops->BcastBegin;
  b->bcastend    = sf->ops->BcastEnd;
  b->reducebegin = sf->ops->ReduceBegin;
  b->reduceend   = sf->ops->ReduceEnd;
  b->fetchbegin  = sf->ops->FetchAndOpBegin;
  b->reset       = sf->ops->Reset;
  PetscCall(PetscContainerCreate(PetscObjectComm((PetscObject)sf), &c));
  PetscCall(PetscContainerSetPointer(c, b));
  PetscCall(PetscContainerSetCtxDestroy(c, PB_SFCtxDestroy));
  PetscCall(PetscObjectCompose((PetscObject)sf, "PetscSFB200_ctx", (PetscObject)c));
  PetscCall(PetscContainerDestroy(&c));
  sf->ops->BcastBegin      = PetscSFBcastBegin_B200;
  sf->ops->BcastEnd        = PetscSFBcastEnd_B200;
  sf->ops->ReduceBegin     = PetscSFReduceBegin_B200;
  sf->ops->ReduceEnd       = PetscSFReduceEnd_B200;
  sf->ops->FetchAndOpBegin = PetscSFFetchAndOpBegin_B200;
  sf->ops->Reset           = PetscSFReset_B200;
  PetscCall(PetscObjectComposeFunction((PetscObject)sf, "PetscSFB200GetCounts_C", PetscSFB200GetCounts));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* ================================================================== registration (src/sys/dll/reg.c:79,150; dl.c:178-199) */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_petscb200plugin(void)
{
  PetscFunctionBegin;
  PetscBool keep = PETSC_FALSE;
  PetscCall(VecRegister(VECSEQB200, VecCreate_SeqB200));
  PetscCall(VecRegister(VECMPIB200, VecCreate_MPIB200));
  PetscCall(VecRegister(VECB200, VecCreate_B200));
  PetscCall(MatRegisterRootName(MATAIJB200, MATSEQAIJB200, MATMPIAIJB200));
  PetscCall(MatRegister(MATSEQAIJB200, MatCreate_SeqAIJB200));
  PetscCall(MatRegister(MATMPIAIJB200, MatCreate_MPIAIJB200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJB200, MAT_FACTOR_ILU, MatGetFactor_seqaijb200_b200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJ, MAT_FACTOR_ILU, MatGetFactor_seqaijb200_b200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJB200, MAT_FACTOR_ICC, MatGetFactor_seqaijb200_b200));
  PetscCall(MatSolverTypeRegister(MATSOLVERB200, MATSEQAIJ, MAT_FACTOR_ICC, MatGetFactor_seqaijb200_b200));
  PetscCall(PCRegister(PCJACOBIB200, PCCreate_JacobiB200));
  PetscCall(KSPRegister(KSPPIPECGB200, KSPCreate_PipeCGB200));
  PetscCall(KSPRegister(KSPPGMRESB200, KSPCreate_PGMRESB200));
  /* -pc_type jacobi is the fused sub-class unless -b200_keep_pcjacobi (PCRegister replaces an existing name; PCRegister
     itself runs PCRegisterAll first, so the stock entry is already there to be replaced) */
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-b200_keep_pcjacobi", &keep, NULL));
  if (!keep) PetscCall(PCRegister(PCJACOBI, PCCreate_JacobiB200));
  /* PetscSF: the parent's creator is not exported by name, it is found in the registry (sfregi.c:33); "basic" then becomes the
     sub-class (host operations are the parent's, unchanged) unless -b200_keep_sfbasic */
  PetscCall(PetscSFInitializePackage());
  PetscCall(PetscFunctionListFind(PetscSFList, PETSCSFBASIC, &PB_SFCreate_Basic));
  PetscCheck(PB_SFCreate_Basic, PETSC_COMM_SELF, PETSC_ERR_PLIB, "PETSCSFBASIC is not registered");
  PetscCall(PetscSFRegister(PETSCSFB200, PetscSFCreate_B200));
  keep = PETSC_FALSE;
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-b200_keep_sfbasic", &keep, NULL));
  if (!keep) PetscCall(PetscSFRegister(PETSCSFBASIC, PetscSFCreate_B200));
  PetscFunctionReturn(PETSC_SUCCESS);
}
