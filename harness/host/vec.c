/* vec.c -- Vec interface (src/vec/vec/interface/{vector.c,rvector.c,vecreg.c}) and the device vector types
   seqb200 / mpib200 (the analogues of VECSEQCUDA / VECMPICUDA: vecseqcupm_impl.hpp, pvecimpl.h:97-172).
   All arithmetic happens in libpetscb200's kernels; the host array is only a mirror governed by the offload mask
   (include/petscdevicetypes.h:240-246, veccupmimpl.h:391-440). */
#include "hostimpl.h"

static PetscFunctionList VecList = NULL;
static int               VecRegisterAllCalled = 0;
static int               fuse_maxpy_norm = 1; /* -vec_b200_fuse_maxpy_norm */

PetscErrorCode VecRegister(const char sname[], PetscErrorCode (*function)(Vec)) { return PetscFunctionListAdd(&VecList, sname, (void *)function); }
static PetscErrorCode VecRegisterAll(void)
{
  if (VecRegisterAllCalled) return PETSC_SUCCESS;
  VecRegisterAllCalled = 1;
  PetscCall(VecRegister(VECSEQB200, VecCreate_SeqB200));
  PetscCall(VecRegister(VECMPIB200, VecCreate_MPIB200));
  PetscCall(VecRegister(VECB200, VecCreate_B200));
  PetscCall(VecRegister(VECSTANDARD, VecCreate_B200));
  return PETSC_SUCCESS;
}

#define H (PetscB200.h)

/* ------------------------------------------------------------------ creation */
PetscErrorCode VecCreate(MPI_Comm comm, Vec *vec)
{
  PetscValidPointer(vec, 2);
  PetscCall(PetscB200EnsureInit());
  Vec v = (Vec)calloc(1, sizeof(*v));
  PetscCheck(v, comm, PETSC_ERR_MEM, "out of memory");
  v->hdr.comm  = comm;
  v->hdr.refct = 1;
  v->n = v->N = -1;
  for (int t = 0; t < 4; t++) v->norm_state[t] = -1;
  v->sumsq_state = -1;
  *vec           = v;
  return PETSC_SUCCESS;
}

PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N)
{
  PetscValidHeader(v, 1);
  PetscCheck(!(N >= 0 && n > N), v->hdr.comm, PETSC_ERR_ARG_INCOMP, "Local size %d cannot be larger than global size %d", n, N);
  PetscCheck(!v->sizes_set || (v->n == n || n < 0) , v->hdr.comm, PETSC_ERR_SUP, "Cannot change/reset vector sizes to %d local %d global after previously setting them to %d local %d global", n, N, v->n, v->N);
  v->n = n;
  v->N = N;
  return PETSC_SUCCESS;
}

static PetscErrorCode VecSetUpLayout(Vec v)
{
  if (v->sizes_set) return PETSC_SUCCESS;
  PetscCheck(v->n >= 0 || v->N >= 0, v->hdr.comm, PETSC_ERR_ORDER, "Must call VecSetSizes() first");
  PetscCall(PetscSplitOwnership(v->hdr.comm, &v->n, &v->N));
  int      size = PetscB200CommSize(v->hdr.comm), rank = PetscB200CommRank(v->hdr.comm);
  int64_t *all = (int64_t *)malloc(sizeof(int64_t) * (size_t)size);
  PetscCall(PetscB200AllgatherInt64(v->hdr.comm, v->n, all));
  int64_t s = 0;
  for (int r = 0; r < rank; r++) s += all[r];
  v->rstart = (PetscInt)s;
  v->rend   = (PetscInt)(s + v->n);
  free(all);
  v->sizes_set = 1;
  return PETSC_SUCCESS;
}

static PetscErrorCode VecAlloc(Vec v)
{
  if (v->d_array || v->n == 0) return PETSC_SUCCESS;
  PetscCallB200(b200Malloc(H, (void **)&v->d_array, sizeof(double) * (size_t)v->n));
  PetscCallB200(b200Memset(H, v->d_array, 0, sizeof(double) * (size_t)v->n));
  v->owns_device = 1;
  v->offloadmask = OFFLOAD_GPU;
  return PETSC_SUCCESS;
}

PetscErrorCode VecSetType(Vec v, VecType type)
{
  PetscErrorCode (*create)(Vec) = NULL;
  PetscValidHeader(v, 1);
  PetscCall(VecRegisterAll());
  if (!strcmp(v->hdr.type_name, type)) return PETSC_SUCCESS;
  PetscCall(PetscFunctionListFind(VecList, type, (void **)&create));
  PetscCheck(create, v->hdr.comm, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unknown vector type: %s", type);
  PetscCheck(!v->type_set || !v->d_array, v->hdr.comm, PETSC_ERR_SUP, "Cannot change the type of a vector that holds data");
  PetscCall((*create)(v));
  v->type_set = 1;
  return PETSC_SUCCESS;
}

PetscErrorCode VecSetFromOptions(Vec v)
{
  char      type[64] = VECB200;
  PetscBool b;
  PetscCall(PetscOptionsGetString(NULL, v->hdr.prefix, "-vec_type", type, sizeof type, NULL));
  if (!strcmp(type, "cuda") || !strcmp(type, "seq") || !strcmp(type, "mpi")) strcpy(type, VECB200);
  PetscCall(VecSetType(v, type));
  b = (PetscBool)fuse_maxpy_norm;
  PetscCall(PetscOptionsGetBool(NULL, NULL, "-vec_b200_fuse_maxpy_norm", &b, NULL));
  fuse_maxpy_norm = b;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetType(Vec v, VecType *t)
{
  *t = v->hdr.type_name;
  return PETSC_SUCCESS;
}

static PetscErrorCode VecSetUp_Private(Vec v)
{
  if (!v->type_set) PetscCall(VecSetType(v, VECB200));
  PetscCall(VecSetUpLayout(v));
  PetscCall(VecAlloc(v));
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ device/host access (offload protocol) */
static PetscErrorCode VecSyncToDevice(Vec v)
{
  PetscCall(VecSetUp_Private(v));
  PetscCheck(!v->array_gotten, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Vector is locked by an outstanding VecGetArray()");
  if (v->offloadmask == OFFLOAD_CPU) {
    PetscCallB200(b200MemcpyHtoDAsync(H, v->d_array, v->h_array, sizeof(double) * (size_t)v->n));
    v->offloadmask = OFFLOAD_BOTH;
  }
  return PETSC_SUCCESS;
}
PetscErrorCode VecB200GetArrayRead(Vec v, const double **d)
{
  PetscCall(VecSyncToDevice(v));
  *d = v->d_array;
  return PETSC_SUCCESS;
}
PetscErrorCode VecB200GetArray(Vec v, double **d)
{
  PetscCall(VecSyncToDevice(v));
  v->offloadmask = OFFLOAD_GPU;
  VecStateIncrease(v);
  *d = v->d_array;
  return PETSC_SUCCESS;
}
PetscErrorCode VecB200GetArrayWrite(Vec v, double **d)
{
  PetscCall(VecSetUp_Private(v));
  PetscCheck(!v->array_gotten, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Vector is locked by an outstanding VecGetArray()");
  v->offloadmask = OFFLOAD_GPU;
  VecStateIncrease(v);
  *d = v->d_array;
  return PETSC_SUCCESS;
}
static PetscErrorCode VecSyncToHost(Vec v, int need_values)
{
  PetscCall(VecSetUp_Private(v));
  if (!v->h_array && v->n) PetscCallB200(b200MallocHost((void **)&v->h_array, sizeof(double) * (size_t)v->n));
  if (need_values && v->offloadmask == OFFLOAD_GPU) {
    PetscCallB200(b200MemcpyDtoH(H, v->h_array, v->d_array, sizeof(double) * (size_t)v->n));
    v->offloadmask = OFFLOAD_BOTH;
  } else if (v->offloadmask == OFFLOAD_BOTH) {
    PetscCallB200(b200Synchronize(H)); /* an asynchronous host->device copy of this array may still be reading it */
  }
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar **a)
{
  PetscCall(VecSyncToHost(v, 1));
  v->offloadmask  = OFFLOAD_CPU;
  v->array_gotten = 2;
  VecStateIncrease(v);
  *a = v->h_array;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetArrayWrite(Vec v, PetscScalar **a)
{
  PetscCall(VecSyncToHost(v, 0));
  v->offloadmask  = OFFLOAD_CPU;
  v->array_gotten = 2;
  VecStateIncrease(v);
  *a = v->h_array;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a)
{
  PetscCall(VecSyncToHost(v, 1));
  v->array_gotten = 1;
  *a              = v->h_array;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a)
{
  v->array_gotten = 0;
  if (a) *a = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArrayWrite(Vec v, PetscScalar **a) { return VecRestoreArray(v, a); }
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a)
{
  v->array_gotten = 0;
  if (a) *a = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetArrayAndMemType(Vec v, PetscScalar **a, PetscMemType *mtype)
{
  PetscCall(VecB200GetArray(v, a));
  if (mtype) *mtype = PETSC_MEMTYPE_CUDA;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArrayAndMemType(Vec v, PetscScalar **a)
{
  (void)v;
  if (a) *a = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetArrayReadAndMemType(Vec v, const PetscScalar **a, PetscMemType *mtype)
{
  PetscCall(VecB200GetArrayRead(v, a));
  if (mtype) *mtype = PETSC_MEMTYPE_CUDA;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreArrayReadAndMemType(Vec v, const PetscScalar **a)
{
  (void)v;
  if (a) *a = NULL;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ local-vector aliases (PCApply_BJacobi_Singleblock, bjacobi.c:579-598) */
PetscErrorCode VecGetLocalVectorRead(Vec v, Vec w)
{
  PetscValidHeader(v, 1);
  PetscValidHeader(w, 2);
  PetscCall(VecSyncToDevice(v));
  PetscCall(VecSetUpLayout(w));
  PetscCheck(w->n == v->n, v->hdr.comm, PETSC_ERR_ARG_SIZ, "local vector has size %d, expected %d", w->n, v->n);
  if (w->owns_device && w->d_array) {
    PetscCallB200(b200Free(H, w->d_array));
    w->d_array = NULL;
  }
  w->owns_device    = 0;
  w->d_array        = v->d_array;
  w->offloadmask    = OFFLOAD_GPU;
  w->localrep_owner = v;
  VecStateIncrease(w);
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetLocalVector(Vec v, Vec w)
{
  PetscCall(VecGetLocalVectorRead(v, w));
  v->offloadmask = OFFLOAD_GPU;
  VecStateIncrease(v);
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreLocalVectorRead(Vec v, Vec w)
{
  (void)v;
  w->d_array        = NULL;
  w->offloadmask    = OFFLOAD_UNALLOCATED;
  w->localrep_owner = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecRestoreLocalVector(Vec v, Vec w)
{
  VecStateIncrease(v);
  return VecRestoreLocalVectorRead(v, w);
}

/* ------------------------------------------------------------------ interface wrappers (checks + state + norm cache) */
#define VecCheckSameLocalSize(x, y) PetscCheck((x)->n == (y)->n, (x)->hdr.comm, PETSC_ERR_ARG_INCOMP, "Incompatible vector local lengths %d != %d", (x)->n, (y)->n)

static PetscErrorCode VecPrep(Vec v)
{
  PetscValidHeader(v, 1);
  return VecSetUp_Private(v);
}

PetscErrorCode VecGetSize(Vec x, PetscInt *size)
{
  PetscCall(VecPrep(x));
  *size = x->N;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetLocalSize(Vec x, PetscInt *size)
{
  PetscCall(VecPrep(x));
  *size = x->n;
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetOwnershipRange(Vec x, PetscInt *low, PetscInt *high)
{
  PetscCall(VecPrep(x));
  if (low) *low = x->rstart;
  if (high) *high = x->rend;
  return PETSC_SUCCESS;
}

PetscErrorCode VecDuplicate(Vec v, Vec *newv)
{
  PetscCall(VecPrep(v));
  PetscCall((*v->ops.duplicate)(v, newv));
  return PETSC_SUCCESS;
}
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt m, Vec *V[])
{
  PetscCall(VecPrep(v));
  PetscCheck(m > 0, v->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "m must be > 0: m = %d", m);
  PetscCall((*v->ops.duplicatevecs)(v, m, V));
  return PETSC_SUCCESS;
}
PetscErrorCode VecDestroy(Vec *v)
{
  if (!v || !*v) return PETSC_SUCCESS;
  if (--(*v)->hdr.refct > 0) {
    *v = NULL;
    return PETSC_SUCCESS;
  }
  if ((*v)->ops.destroy) PetscCall((*(*v)->ops.destroy)(*v));
  free(*v);
  *v = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode VecDestroyVecs(PetscInt m, Vec *vv[])
{
  if (!vv || !*vv) return PETSC_SUCCESS;
  for (PetscInt i = 0; i < m; i++) PetscCall(VecDestroy(&(*vv)[i]));
  free(*vv);
  *vv = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode VecSet(Vec x, PetscScalar alpha)
{
  PetscCall(VecPrep(x));
  PetscCall((*x->ops.set)(x, alpha));
  /* rvector.c VecSet: norms of a constant vector are known */
  return PETSC_SUCCESS;
}
PetscErrorCode VecZeroEntries(Vec x) { return VecSet(x, 0.0); }
PetscErrorCode VecCopy(Vec x, Vec y)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  if (x == y) return PETSC_SUCCESS;
  VecCheckSameLocalSize(x, y);
  PetscCall((*x->ops.copy)(x, y));
  for (int t = 0; t < 4; t++) /* rvector.c VecCopy carries the cached norms over */
    if (x->norm_state[t] == x->hdr.state) {
      y->norm_state[t] = y->hdr.state;
      y->norm_val[t]   = x->norm_val[t];
    }
  return PETSC_SUCCESS;
}
PetscErrorCode VecSwap(Vec x, Vec y)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  PetscCall(VecSyncToDevice(x));
  PetscCall(VecSyncToDevice(y));
  PetscCheck(x->owns_device == y->owns_device || 1, 0, 0, " ");
  /* three copies through a temporary keeps slab layouts intact */
  Vec t;
  PetscCall(VecDuplicate(x, &t));
  PetscCall(VecCopy(x, t));
  PetscCall(VecCopy(y, x));
  PetscCall(VecCopy(t, y));
  PetscCall(VecDestroy(&t));
  return PETSC_SUCCESS;
}
PetscErrorCode VecScale(Vec x, PetscScalar alpha)
{
  PetscCall(VecPrep(x));
  if (alpha == 1.0) return PETSC_SUCCESS;
  /* rvector.c:1010-1040: cached norms scale with |alpha| */
  PetscReal nv[4];
  int       have[4];
  for (int t = 0; t < 4; t++) {
    have[t] = x->norm_state[t] == x->hdr.state;
    nv[t]   = x->norm_val[t];
  }
  PetscCall((*x->ops.scale)(x, alpha));
  for (int t = 0; t < 4; t++)
    if (have[t]) {
      x->norm_state[t] = x->hdr.state;
      x->norm_val[t]   = fabs(alpha) * nv[t];
    }
  return PETSC_SUCCESS;
}
PetscErrorCode VecShift(Vec v, PetscScalar shift)
{
  PetscCall(VecPrep(v));
  if (shift == 0.0) return PETSC_SUCCESS;
  return (*v->ops.shift)(v, shift);
}
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  PetscCheck(x != y, y->hdr.comm, PETSC_ERR_ARG_IDN, "x and y cannot be the same vector");
  if (alpha == 0.0) return PETSC_SUCCESS;
  return (*y->ops.axpy)(y, alpha, x);
}
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  PetscCheck(x != y, y->hdr.comm, PETSC_ERR_ARG_IDN, "x and y cannot be the same vector");
  return (*y->ops.aypx)(y, beta, x);
}
PetscErrorCode VecAXPBY(Vec y, PetscScalar alpha, PetscScalar beta, Vec x)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  PetscCheck(x != y, y->hdr.comm, PETSC_ERR_ARG_IDN, "x and y cannot be the same vector");
  if (alpha == 0.0 && beta == 1.0) return PETSC_SUCCESS;
  return (*y->ops.axpby)(y, alpha, beta, x);
}
PetscErrorCode VecWAXPY(Vec w, PetscScalar alpha, Vec x, Vec y)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  PetscCall(VecPrep(w));
  VecCheckSameLocalSize(x, y);
  VecCheckSameLocalSize(x, w);
  PetscCheck(w != y && w != x, w->hdr.comm, PETSC_ERR_SUP, "Result vector w cannot be same as input vectors, suggest VecAXPY()");
  return (*w->ops.waxpy)(w, alpha, x, y);
}
PetscErrorCode VecMAXPY(Vec y, PetscInt nv, const PetscScalar alpha[], Vec x[])
{
  PetscCall(VecPrep(y));
  PetscCheck(nv >= 0, y->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Number of vectors (given %d) cannot be negative", nv);
  if (!nv) return PETSC_SUCCESS;
  for (PetscInt i = 0; i < nv; i++) {
    PetscCall(VecPrep(x[i]));
    VecCheckSameLocalSize(y, x[i]);
    PetscCheck(x[i] != y, y->hdr.comm, PETSC_ERR_ARG_IDN, "x[%d] and y cannot be the same vector", i);
  }
  return (*y->ops.maxpy)(y, nv, alpha, x);
}
PetscErrorCode VecMAXPBY(Vec y, PetscInt nv, const PetscScalar alpha[], PetscScalar beta, Vec x[])
{
  /* rvector.c:1434-1438 (no maxpby op): scale or zero y, then VecMAXPY */
  if (beta == 0.0) PetscCall(VecSet(y, 0.0));
  else PetscCall(VecScale(y, beta));
  return VecMAXPY(y, nv, alpha, x);
}
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val)
{
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  return (*x->ops.dot)(x, y, val);
}
PetscErrorCode VecTDot(Vec x, Vec y, PetscScalar *val) { return VecDot(x, y, val); }
PetscErrorCode VecMDot(Vec x, PetscInt nv, const Vec y[], PetscScalar val[])
{
  PetscCall(VecPrep(x));
  PetscCheck(nv >= 0, x->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Number of vectors (given %d) cannot be negative", nv);
  if (!nv) return PETSC_SUCCESS;
  for (PetscInt i = 0; i < nv; i++) {
    PetscCall(VecPrep(y[i]));
    VecCheckSameLocalSize(x, y[i]);
  }
  return (*x->ops.mdot)(x, nv, y, val);
}
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val)
{
  PetscCall(VecPrep(x));
  PetscCheck(type == NORM_1 || type == NORM_2 || type == NORM_FROBENIUS || type == NORM_INFINITY, x->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Unknown norm type %d", (int)type);
  if (type == NORM_FROBENIUS) type = NORM_2;
  if (x->norm_state[type] == x->hdr.state) { /* rvector.c:211 */
    *val = x->norm_val[type];
    return PETSC_SUCCESS;
  }
  PetscCall((*x->ops.norm)(x, type, val));
  x->norm_state[type] = x->hdr.state; /* rvector.c:232 */
  x->norm_val[type]   = *val;
  return PETSC_SUCCESS;
}
PetscErrorCode VecNormalize(Vec x, PetscReal *val)
{
  PetscReal norm;
  PetscCall(VecNorm(x, NORM_2, &norm)); /* rvector.c:289-311 */
  if (norm == 0.0) {
    /* zero vector: left as is */
  } else if (isinf(norm) || isnan(norm)) {
    /* left as is */
  } else {
    PetscCall(VecScale(x, 1.0 / norm));
  }
  if (val) *val = norm;
  return PETSC_SUCCESS;
}
PetscErrorCode VecSum(Vec v, PetscScalar *sum)
{
  PetscCall(VecPrep(v));
  return (*v->ops.sum)(v, sum);
}
PetscErrorCode VecMax(Vec x, PetscInt *p, PetscReal *val)
{
  PetscCall(VecPrep(x));
  return (*x->ops.max)(x, p, val);
}
PetscErrorCode VecMin(Vec x, PetscInt *p, PetscReal *val)
{
  PetscCall(VecPrep(x));
  return (*x->ops.min)(x, p, val);
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
  PetscCall(VecPrep(w));
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  VecCheckSameLocalSize(x, w);
  return (*w->ops.pointwisemult)(w, x, y);
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y)
{
  PetscCall(VecPrep(w));
  PetscCall(VecPrep(x));
  PetscCall(VecPrep(y));
  VecCheckSameLocalSize(x, y);
  VecCheckSameLocalSize(x, w);
  return (*w->ops.pointwisedivide)(w, x, y);
}
PetscErrorCode VecReciprocal(Vec v)
{
  PetscCall(VecPrep(v));
  return (*v->ops.reciprocal)(v);
}
PetscErrorCode VecSetValues(Vec x, PetscInt ni, const PetscInt ix[], const PetscScalar y[], InsertMode iora)
{
  PetscScalar *a;
  PetscCall(VecPrep(x));
  PetscCall(VecGetArray(x, &a));
  for (PetscInt i = 0; i < ni; i++) {
    if (ix[i] < 0) continue;
    PetscCheck(ix[i] >= x->rstart && ix[i] < x->rend, x->hdr.comm, PETSC_ERR_SUP, "VecSetValues of off-process entry %d (owned range [%d,%d)) is not supported by the b200 vector types", ix[i], x->rstart, x->rend);
    if (iora == ADD_VALUES) a[ix[i] - x->rstart] += y[i];
    else a[ix[i] - x->rstart] = y[i];
  }
  PetscCall(VecRestoreArray(x, &a));
  return PETSC_SUCCESS;
}
PetscErrorCode VecAssemblyBegin(Vec vec)
{
  (void)vec;
  return PETSC_SUCCESS;
}
PetscErrorCode VecAssemblyEnd(Vec vec)
{
  (void)vec;
  return PETSC_SUCCESS;
}
PetscErrorCode VecCreateSeqWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, const PetscScalar array[], Vec *V)
{
  (void)bs;
  PetscCall(VecCreate(comm == PETSC_COMM_WORLD && PetscB200.size > 1 ? PETSC_COMM_SELF : comm, V));
  PetscCall(VecSetSizes(*V, n, n));
  PetscCall(VecSetType(*V, VECSEQB200));
  PetscCall(VecSetUp_Private(*V));
  if (array) PetscCall(VecPlaceArray(*V, array)); /* the user's array IS the host storage (bvec2.c VecCreateSeqWithArray) */
  return PETSC_SUCCESS;
}
PetscErrorCode VecCreateMPIWithArray(MPI_Comm comm, PetscInt bs, PetscInt n, PetscInt N, const PetscScalar array[], Vec *V)
{
  (void)bs;
  PetscCall(VecCreate(comm, V));
  PetscCall(VecSetSizes(*V, n, N));
  PetscCall(VecSetType(*V, PetscB200CommSize(comm) > 1 ? VECMPIB200 : VECSEQB200));
  PetscCall(VecSetUp_Private(*V));
  if (array) PetscCall(VecPlaceArray(*V, array));
  return PETSC_SUCCESS;
}
/* VecPlaceArray / VecResetArray (rvector.c:2593, bvec2.c VecPlaceArray_Seq): the vector uses the user's HOST array as its host
   storage from now on; its current contents are the vector's values (the device mirror is refreshed on the next device use,
   straight from that array: pinned user memory is copied at full PCIe speed) */
PetscErrorCode VecPlaceArray(Vec v, const PetscScalar array[])
{
  PetscValidHeader(v, 1);
  PetscCall(VecSetUp_Private(v));
  PetscCheck(!v->array_gotten, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Vector is locked by an outstanding VecGetArray()");
  PetscCheck(!v->h_saved && !v->h_saved_user, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "VecPlaceArray() was already called on this vector, without a call to VecResetArray()");
  PetscCheck(array || !v->n, v->hdr.comm, PETSC_ERR_ARG_NULL, "null array");
  v->h_saved      = v->h_array;
  v->h_saved_user = v->h_user | 2; /* bit 1: something is saved */
  v->h_array      = (double *)array;
  v->h_user       = 1;
  v->offloadmask  = OFFLOAD_CPU;
  VecStateIncrease(v);
  return PETSC_SUCCESS;
}
PetscErrorCode VecResetArray(Vec v)
{
  PetscValidHeader(v, 1);
  PetscCheck(v->h_saved_user & 2, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "VecResetArray() without VecPlaceArray()");
  PetscCheck(!v->array_gotten, v->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Vector is locked by an outstanding VecGetArray()");
  v->h_array      = v->h_saved;
  v->h_user       = v->h_saved_user & 1;
  v->h_saved      = NULL;
  v->h_saved_user = 0;
  /* the original array's contents become the values again (bvec2.c VecResetArray_Seq); without one the device copy stays */
  if (v->h_array) {
    v->offloadmask = OFFLOAD_CPU;
    VecStateIncrease(v);
  } else if (v->offloadmask == OFFLOAD_CPU) {
    PetscCallB200(b200Memset(H, v->d_array, 0, sizeof(double) * (size_t)v->n));
    v->offloadmask = OFFLOAD_GPU;
    VecStateIncrease(v);
  } else v->offloadmask = OFFLOAD_GPU; /* no host mirror left: the device copy is the only one */
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ seqb200 implementation */
static PetscErrorCode VecDestroy_B200(Vec v)
{
  if (v->slab_ref) {
    if (--(*v->slab_ref) == 0) {
      PetscCallB200(b200Free(H, v->slab));
      free(v->slab_ref);
    }
  } else if (v->owns_device && v->d_array) PetscCallB200(b200Free(H, v->d_array));
  if (v->h_array && !v->h_user) PetscCallB200(b200FreeHost(v->h_array));
  if (v->h_saved && !(v->h_saved_user & 1)) PetscCallB200(b200FreeHost(v->h_saved));
  if (v->d_sumsq) PetscCallB200(b200Free(H, v->d_sumsq));
  return PETSC_SUCCESS;
}

static PetscErrorCode VecDuplicate_B200(Vec v, Vec *newv)
{
  Vec w;
  PetscCall(VecCreate(v->hdr.comm, &w));
  w->n = v->n; w->N = v->N; w->rstart = v->rstart; w->rend = v->rend;
  w->sizes_set = 1;
  w->ops       = v->ops;
  w->type_set  = 1;
  strcpy(w->hdr.type_name, v->hdr.type_name);
  PetscCall(VecAlloc(w));
  *newv = w;
  return PETSC_SUCCESS;
}

/* VecDuplicateVecs_Seq_GEMV layout (bvec2.c:670-692): one slab, leading dimension rounded up to 64 bytes */
static PetscErrorCode VecDuplicateVecs_B200(Vec v, PetscInt m, Vec **V)
{
  size_t  lda = ((size_t)v->n + 7) & ~(size_t)7;
  double *slab = NULL;
  int    *ref  = (int *)malloc(sizeof(int));
  Vec    *vv   = (Vec *)calloc((size_t)m, sizeof(Vec));
  PetscCheck(ref && vv, v->hdr.comm, PETSC_ERR_MEM, "out of memory");
  if (lda) {
    PetscCallB200(b200Malloc(H, (void **)&slab, sizeof(double) * lda * (size_t)m));
    PetscCallB200(b200Memset(H, slab, 0, sizeof(double) * lda * (size_t)m));
  }
  *ref = m;
  for (PetscInt i = 0; i < m; i++) {
    Vec w;
    PetscCall(VecCreate(v->hdr.comm, &w));
    w->n = v->n; w->N = v->N; w->rstart = v->rstart; w->rend = v->rend;
    w->sizes_set = 1;
    w->ops       = v->ops;
    w->type_set  = 1;
    strcpy(w->hdr.type_name, v->hdr.type_name);
    w->d_array     = slab ? slab + lda * (size_t)i : NULL;
    w->owns_device = 0;
    w->slab        = slab;
    w->slab_ref    = ref;
    w->offloadmask = OFFLOAD_GPU;
    vv[i]          = w;
  }
  *V = vv;
  return PETSC_SUCCESS;
}

#define RD(v, p) \
  const double *p; \
  PetscCall(VecB200GetArrayRead(v, &p))
#define RW(v, p) \
  double *p; \
  PetscCall(VecB200GetArray(v, &p))
#define WR(v, p) \
  double *p; \
  PetscCall(VecB200GetArrayWrite(v, &p))

static PetscErrorCode VecSet_B200(Vec x, PetscScalar a)
{
  WR(x, d);
  PetscCallB200(b200VecSet(H, x->n, a, d));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecCopy_B200(Vec x, Vec y)
{
  RD(x, dx);
  WR(y, dy);
  PetscCallB200(b200VecCopy(H, x->n, dx, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecScale_B200(Vec x, PetscScalar a)
{
  RW(x, d);
  if (a == 0.0) PetscCallB200(b200VecSet(H, x->n, 0.0, d));
  else PetscCallB200(b200VecScale(H, x->n, a, d));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecShift_B200(Vec x, PetscScalar s)
{
  RW(x, d);
  PetscCallB200(b200VecShift(H, x->n, s, d));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecAXPY_B200(Vec y, PetscScalar a, Vec x)
{
  RD(x, dx);
  RW(y, dy);
  PetscCallB200(b200VecAXPY(H, y->n, a, dx, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecAYPX_B200(Vec y, PetscScalar b, Vec x)
{
  RD(x, dx);
  RW(y, dy);
  PetscCallB200(b200VecAYPX(H, y->n, b, dx, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecAXPBY_B200(Vec y, PetscScalar a, PetscScalar b, Vec x)
{
  RD(x, dx);
  RW(y, dy);
  PetscCallB200(b200VecAXPBY(H, y->n, a, b, dx, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecWAXPY_B200(Vec w, PetscScalar a, Vec x, Vec y)
{
  RD(x, dx);
  RD(y, dy);
  WR(w, dw);
  PetscCallB200(b200VecWAXPY(H, w->n, a, dx, dy, dw));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecPointwiseMult_B200(Vec w, Vec x, Vec y)
{
  RD(x, dx);
  RD(y, dy);
  double *dw;
  if (w == x || w == y) PetscCall(VecB200GetArray(w, &dw));
  else PetscCall(VecB200GetArrayWrite(w, &dw));
  PetscCallB200(b200VecPointwiseMult(H, w->n, dx, dy, dw));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecPointwiseDivide_B200(Vec w, Vec x, Vec y)
{
  RD(x, dx);
  RD(y, dy);
  double *dw;
  if (w == x || w == y) PetscCall(VecB200GetArray(w, &dw));
  else PetscCall(VecB200GetArrayWrite(w, &dw));
  PetscCallB200(b200VecPointwiseDivide(H, w->n, dx, dy, dw));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecReciprocal_B200(Vec x)
{
  RW(x, d);
  PetscCallB200(b200VecReciprocal(H, x->n, d));
  return PETSC_SUCCESS;
}

/* reductions: *_local leave the per-rank value on the host; the MPI type adds the all-reduce */
static PetscErrorCode VecDot_Local(Vec x, Vec y, PetscScalar *val)
{
  RD(x, dx);
  RD(y, dy);
  PetscCallB200(b200VecDot(H, x->n, dx, dy, val));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecMDot_Local(Vec x, PetscInt nv, const Vec y[], PetscScalar *val)
{
  RD(x, dx);
  const double **yp = (const double **)malloc(sizeof(double *) * (size_t)nv);
  for (PetscInt j = 0; j < nv; j++) PetscCall(VecB200GetArrayRead(y[j], &yp[j]));
  PetscErrorCode rc = b200VecMDot(H, x->n, nv, dx, yp, val);
  free(yp);
  PetscCallB200(rc);
  return PETSC_SUCCESS;
}
static PetscErrorCode VecSumsqFetch(Vec x, double *ss)
{
  PetscCallB200(b200MemcpyDtoH(H, ss, x->d_sumsq, sizeof(double)));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecNorm_Local(Vec x, NormType type, PetscReal *val)
{
  if (type == NORM_2 && x->sumsq_state == x->hdr.state && x->d_sumsq) { /* left behind by the fused VecMAXPY */
    double ss;
    PetscCall(VecSumsqFetch(x, &ss));
    *val = sqrt(ss);
    return PETSC_SUCCESS;
  }
  RD(x, dx);
  PetscCallB200(b200VecNorm(H, x->n, dx, type == NORM_1 ? 0 : (type == NORM_INFINITY ? 3 : 1), val));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecMAXPY_B200(Vec y, PetscInt nv, const PetscScalar *alpha, Vec *x)
{
  const double **xp = (const double **)malloc(sizeof(double *) * (size_t)nv);
  for (PetscInt j = 0; j < nv; j++) PetscCall(VecB200GetArrayRead(x[j], &xp[j]));
  RW(y, dy);
  if (fuse_maxpy_norm && !y->d_sumsq) PetscCallB200(b200Malloc(H, (void **)&y->d_sumsq, sizeof(double)));
  PetscErrorCode rc = b200VecMAXPYAsync(H, y->n, nv, alpha, xp, dy, fuse_maxpy_norm ? y->d_sumsq : NULL);
  free(xp);
  PetscCallB200(rc);
  if (fuse_maxpy_norm) y->sumsq_state = y->hdr.state;
  return PETSC_SUCCESS;
}
static PetscErrorCode VecSum_B200(Vec x, PetscScalar *s)
{
  RD(x, dx);
  PetscCallB200(b200VecSum(H, x->n, dx, s));
  if (PetscB200CommSize(x->hdr.comm) > 1) PetscCall(PetscB200AllreduceHost(x->hdr.comm, s, 1, 0));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecMax_B200(Vec x, PetscInt *p, PetscReal *val)
{
  RD(x, dx);
  int64_t idx = -1;
  PetscCallB200(b200VecMax(H, x->n, dx, p ? &idx : NULL, val));
  if (PetscB200CommSize(x->hdr.comm) > 1) {
    PetscCheck(!p, x->hdr.comm, PETSC_ERR_SUP, "VecMax location is not supported in parallel by mpib200");
    PetscCall(PetscB200AllreduceHost(x->hdr.comm, val, 1, 1));
  }
  if (p) *p = (PetscInt)idx;
  return PETSC_SUCCESS;
}
static PetscErrorCode VecMin_B200(Vec x, PetscInt *p, PetscReal *val)
{
  RD(x, dx);
  int64_t idx = -1;
  PetscCallB200(b200VecMin(H, x->n, dx, p ? &idx : NULL, val));
  if (PetscB200CommSize(x->hdr.comm) > 1) {
    PetscCheck(!p, x->hdr.comm, PETSC_ERR_SUP, "VecMin location is not supported in parallel by mpib200");
    double neg = -*val;
    PetscCall(PetscB200AllreduceHost(x->hdr.comm, &neg, 1, 1));
    *val = -neg;
  }
  if (p) *p = (PetscInt)idx;
  return PETSC_SUCCESS;
}

static void VecSetOps_Common(Vec v)
{
  v->ops.duplicate       = VecDuplicate_B200;
  v->ops.duplicatevecs   = VecDuplicateVecs_B200;
  v->ops.destroy         = VecDestroy_B200;
  v->ops.scale           = VecScale_B200;
  v->ops.copy            = VecCopy_B200;
  v->ops.set             = VecSet_B200;
  v->ops.axpy            = VecAXPY_B200;
  v->ops.aypx            = VecAYPX_B200;
  v->ops.axpby           = VecAXPBY_B200;
  v->ops.waxpy           = VecWAXPY_B200;
  v->ops.maxpy           = VecMAXPY_B200;
  v->ops.pointwisemult   = VecPointwiseMult_B200;
  v->ops.pointwisedivide = VecPointwiseDivide_B200;
  v->ops.reciprocal      = VecReciprocal_B200;
  v->ops.shift           = VecShift_B200;
  v->ops.sum             = VecSum_B200;
  v->ops.max             = VecMax_B200;
  v->ops.min             = VecMin_B200;
  v->ops.dot_local       = VecDot_Local;
  v->ops.mdot_local      = VecMDot_Local;
  v->ops.norm_local      = VecNorm_Local;
}

PetscErrorCode VecCreate_SeqB200(Vec v)
{
  PetscCheck(PetscB200CommSize(v->hdr.comm) == 1, v->hdr.comm, PETSC_ERR_ARG_WRONG, "Cannot create VECSEQB200 on more than one process");
  VecSetOps_Common(v);
  v->ops.dot  = VecDot_Local;
  v->ops.mdot = VecMDot_Local;
  v->ops.norm = VecNorm_Local;
  strcpy(v->hdr.type_name, VECSEQB200);
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ mpib200: local kernels + NCCL all-reduce (pvecimpl.h:97-172) */
static double *d_red = NULL; /* device scratch for the fused local-reduce + all-reduce */
static PetscErrorCode VecMDot_MPIB200(Vec x, PetscInt nv, const Vec y[], PetscScalar *val);
static PetscErrorCode VecDot_MPIB200(Vec x, Vec y, PetscScalar *val)
{
  /* VecXDot_MPI_Default (pvecimpl.h:97-121): the local dot is the nv = 1 MDot kernel, its device result is all-reduced in
     place on persistent scratch, one device-to-host copy */
  return VecMDot_MPIB200(x, 1, &y, val);
}
static PetscErrorCode VecMDot_MPIB200(Vec x, PetscInt nv, const Vec y[], PetscScalar *val)
{
  /* VecMXDot_MPI_Default: local mdot then MPIU_Allreduce(nv). Here the local results never leave the device before the
     all-reduce: MDot kernel -> ncclAllReduce(nv doubles) on the same stream -> one device-to-host copy */
  RD(x, dx);
  if (!d_red) PetscCallB200(b200Malloc(H, (void **)&d_red, sizeof(double) * 4096));
  PetscCheck(nv <= 4096, x->hdr.comm, PETSC_ERR_SUP, "nv too large");
  const double *yp[4096];
  for (PetscInt j = 0; j < nv; j++) PetscCall(VecB200GetArrayRead(y[j], &yp[j]));
  PetscCallB200(b200VecMDotAsync(H, x->n, nv, dx, yp, d_red));
  PetscCallB200(b200CommAllreduceSum(H, d_red, nv));
  PetscCallB200(b200MemcpyDtoH(H, val, d_red, sizeof(double) * (size_t)nv));
  return PETSC_SUCCESS;
}
static PetscErrorCode VecNorm_MPIB200(Vec x, NormType type, PetscReal *val)
{
  /* VecNorm_MPI_Default (pvecimpl.h:150-172): square / all-reduce / sqrt for NORM_2, sum for NORM_1, max for NORM_INFINITY */
  if (type == NORM_2 && x->sumsq_state == x->hdr.state && x->d_sumsq) {
    PetscCallB200(b200CommAllreduceSum(H, x->d_sumsq, 1));
    double ss;
    PetscCall(VecSumsqFetch(x, &ss));
    x->sumsq_state = -1; /* the buffer now holds the global value: do not reduce it twice */
    *val           = sqrt(ss);
    return PETSC_SUCCESS;
  }
  PetscCall(VecNorm_Local(x, type, val));
  if (type == NORM_2) {
    *val = *val * *val;
    PetscCall(PetscB200AllreduceHost(x->hdr.comm, val, 1, 0));
    *val = sqrt(*val);
  } else PetscCall(PetscB200AllreduceHost(x->hdr.comm, val, 1, type == NORM_1 ? 0 : 1));
  return PETSC_SUCCESS;
}

PetscErrorCode VecCreate_MPIB200(Vec v)
{
  VecSetOps_Common(v);
  v->ops.dot  = VecDot_MPIB200;
  v->ops.mdot = VecMDot_MPIB200;
  v->ops.norm = VecNorm_MPIB200;
  strcpy(v->hdr.type_name, VECMPIB200);
  return PETSC_SUCCESS;
}

/* VecCreate_B200: by communicator size, as VecCreate_CUDA does (vecreg.c:87-93 family rule) */
PetscErrorCode VecCreate_B200(Vec v)
{
  if (PetscB200CommSize(v->hdr.comm) == 1) return VecCreate_SeqB200(v);
  return VecCreate_MPIB200(v);
}
