/* sys.c -- the slice of PETSc's Sys layer the Krylov path needs: error traceback (src/sys/error/err.c), options database
   (src/sys/objects/options.c), function lists (src/sys/dll/reg.c), ownership split (src/sys/utils/psplit.c:88),
   and the NCCL-backed replacement of the few MPI collectives the path uses. */
#include "hostimpl.h"
#include <stdarg.h>

PetscB200Globals PetscB200 = {0, NULL, -1, 0, 1};

/* ------------------------------------------------------------------ errors */
static __thread char errbuf[8192];
static __thread int  errlen = 0;

PetscErrorCode PetscB200Error(const char *file, int line, const char *func, PetscErrorCode code, int initial, const char *fmt, ...)
{
  if (initial) {
    va_list ap;
    errlen = snprintf(errbuf, sizeof errbuf, "[%d]PETSC ERROR: ", PetscB200.rank);
    va_start(ap, fmt);
    errlen += vsnprintf(errbuf + errlen, sizeof errbuf - (size_t)errlen, fmt, ap);
    va_end(ap);
    if (errlen > (int)sizeof errbuf - 1) errlen = (int)sizeof errbuf - 1;
  }
  if (errlen < (int)sizeof errbuf - 200) errlen += snprintf(errbuf + errlen, sizeof errbuf - (size_t)errlen, "\n[%d]PETSC ERROR: #%s() at %s:%d (error code %d)", PetscB200.rank, func, file, line, code);
  return code ? code : PETSC_ERR_PLIB;
}
const char *PetscB200GetLastErrorMessage(void) { return errbuf; }

/* ------------------------------------------------------------------ function lists */
PetscErrorCode PetscFunctionListAdd(PetscFunctionList *fl, const char name[], void *fn)
{
  PetscFunctionList e;
  for (e = *fl; e; e = e->next)
    if (!strcmp(e->name, name)) {
      e->fn = fn;
      return PETSC_SUCCESS;
    }
  e = (PetscFunctionList)malloc(sizeof(*e));
  PetscCheck(e, 0, PETSC_ERR_MEM, "out of memory");
  e->name = strdup(name);
  e->fn   = fn;
  e->next = *fl;
  *fl     = e;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscFunctionListFind(PetscFunctionList fl, const char name[], void **fn)
{
  *fn = NULL;
  for (; fl; fl = fl->next)
    if (!strcmp(fl->name, name)) {
      *fn = fl->fn;
      break;
    }
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ options database */
#define MAXOPT 512
static struct {
  char *name, *value;
} opts[MAXOPT];
static int nopts = 0;

PetscErrorCode PetscOptionsSetValue(void *options, const char name[], const char value[])
{
  (void)options;
  PetscCheck(name && name[0] == '-', 0, PETSC_ERR_ARG_WRONG, "option name must start with '-': %s", name ? name : "(null)");
  for (int i = 0; i < nopts; i++)
    if (!strcmp(opts[i].name, name + 1)) {
      free(opts[i].value);
      opts[i].value = value ? strdup(value) : NULL;
      return PETSC_SUCCESS;
    }
  PetscCheck(nopts < MAXOPT, 0, PETSC_ERR_MEM, "options table full");
  opts[nopts].name  = strdup(name + 1);
  opts[nopts].value = value ? strdup(value) : NULL;
  nopts++;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsClearValue(void *options, const char name[])
{
  (void)options;
  for (int i = 0; i < nopts; i++)
    if (!strcmp(opts[i].name, name + 1)) {
      free(opts[i].name);
      free(opts[i].value);
      opts[i] = opts[--nopts];
      break;
    }
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsClear(void *options)
{
  (void)options;
  for (int i = 0; i < nopts; i++) {
    free(opts[i].name);
    free(opts[i].value);
  }
  nopts = 0;
  return PETSC_SUCCESS;
}
static int is_value_token(const char *t)
{
  /* a token is a value unless it looks like an option name: '-' followed by a non-digit, non-'.' */
  if (t[0] != '-') return 1;
  return (t[1] >= '0' && t[1] <= '9') || t[1] == '.';
}
PetscErrorCode PetscOptionsInsertString(void *options, const char in_str[])
{
  char *s, *tok, *save = NULL, *pend = NULL;
  if (!in_str) return PETSC_SUCCESS;
  s = strdup(in_str);
  for (tok = strtok_r(s, " \t\n", &save); tok; tok = strtok_r(NULL, " \t\n", &save)) {
    if (!is_value_token(tok)) {
      if (pend) PetscCall(PetscOptionsSetValue(options, pend, NULL));
      pend = tok;
    } else if (pend) {
      PetscCall(PetscOptionsSetValue(options, pend, tok));
      pend = NULL;
    }
  }
  if (pend) PetscCall(PetscOptionsSetValue(options, pend, NULL));
  free(s);
  return PETSC_SUCCESS;
}
const char *PetscB200OptionsFind(const char *pre, const char *name)
{
  char key[256];
  snprintf(key, sizeof key, "%s%s", pre ? pre : "", name[0] == '-' ? name + 1 : name);
  for (int i = 0; i < nopts; i++)
    if (!strcmp(opts[i].name, key)) return opts[i].value ? opts[i].value : "";
  return NULL;
}
PetscErrorCode PetscOptionsGetInt(void *o, const char pre[], const char name[], PetscInt *v, PetscBool *set)
{
  const char *s = PetscB200OptionsFind(pre, name);
  (void)o;
  if (set) *set = s && *s ? PETSC_TRUE : PETSC_FALSE;
  if (s && *s) *v = (PetscInt)strtol(s, NULL, 10);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsGetReal(void *o, const char pre[], const char name[], PetscReal *v, PetscBool *set)
{
  const char *s = PetscB200OptionsFind(pre, name);
  (void)o;
  if (set) *set = s && *s ? PETSC_TRUE : PETSC_FALSE;
  if (s && *s) *v = strtod(s, NULL);
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsGetBool(void *o, const char pre[], const char name[], PetscBool *v, PetscBool *set)
{
  const char *s = PetscB200OptionsFind(pre, name);
  (void)o;
  if (set) *set = s ? PETSC_TRUE : PETSC_FALSE;
  if (s) *v = (!*s || !strcmp(s, "1") || !strcmp(s, "true") || !strcmp(s, "yes") || !strcmp(s, "on")) ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscOptionsGetString(void *o, const char pre[], const char name[], char str[], size_t len, PetscBool *set)
{
  const char *s = PetscB200OptionsFind(pre, name);
  (void)o;
  if (set) *set = s && *s ? PETSC_TRUE : PETSC_FALSE;
  if (s && *s) {
    strncpy(str, s, len);
    str[len - 1] = 0;
  }
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ init / finalize / comm */
static int requested_device = -1;
PetscErrorCode PetscB200SetDevice(int device)
{
  PetscCheck(!PetscB200.initialized, 0, PETSC_ERR_ORDER, "PetscB200SetDevice() must precede PetscInitialize()");
  requested_device = device;
  return PETSC_SUCCESS;
}

PetscErrorCode PetscB200EnsureInit(void)
{
  if (PetscB200.initialized) return PETSC_SUCCESS;
  int dev = requested_device, ndev = 0;
  PetscCallB200(b200DeviceCount(&ndev));
  PetscCheck(ndev > 0, 0, PETSC_ERR_GPU, "no CUDA device visible: libpetscb200 has no CPU fallback");
  if (dev < 0) { /* rank -> GPU = rank % ndev (cupmdevice.cxx:293), using torchrun's LOCAL_RANK when present */
    const char *lr = getenv("LOCAL_RANK");
    dev            = lr ? atoi(lr) % ndev : 0;
  }
  PetscCallB200(b200Create(&PetscB200.h, dev));
  PetscB200.device      = dev;
  PetscB200.initialized = 1;
  return PETSC_SUCCESS;
}

PetscErrorCode PetscInitialize(int *argc, char ***args, const char file[], const char help[])
{
  (void)file;
  (void)help;
  if (argc && args) {
    for (int i = 1; i < *argc; i++) {
      const char *t = (*args)[i];
      if (!is_value_token(t)) {
        const char *val = (i + 1 < *argc && is_value_token((*args)[i + 1])) ? (*args)[++i] : NULL;
        PetscCall(PetscOptionsSetValue(NULL, t, val));
      }
    }
  }
  {
    const char *env = getenv("PETSC_OPTIONS"); /* options.c: PETSC_OPTIONS environment variable */
    if (env) PetscCall(PetscOptionsInsertString(NULL, env));
  }
  PetscCall(PetscB200EnsureInit());
  return PETSC_SUCCESS;
}
PetscErrorCode PetscInitializeNoArguments(void) { return PetscInitialize(NULL, NULL, NULL, NULL); }

PetscErrorCode PetscFinalize(void)
{
  if (PetscB200.initialized) {
    b200Destroy(PetscB200.h);
    PetscB200.h           = NULL;
    PetscB200.initialized = 0;
    PetscB200.rank        = 0;
    PetscB200.size        = 1;
  }
  PetscOptionsClear(NULL);
  return PETSC_SUCCESS;
}

PetscErrorCode PetscB200GetHandle(void **h)
{
  PetscCall(PetscB200EnsureInit());
  *h = PetscB200.h;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscB200CommGetUniqueId(void *id128)
{
  PetscCallB200(b200CommGetUniqueId(id128));
  return PETSC_SUCCESS;
}
PetscErrorCode PetscB200CommInit(PetscMPIInt rank, PetscMPIInt size, const void *id128)
{
  PetscCall(PetscB200EnsureInit());
  if (size > 1) PetscCallB200(b200CommInitRank(PetscB200.h, size, rank, id128));
  PetscB200.rank = rank;
  PetscB200.size = size;
  return PETSC_SUCCESS;
}
PetscErrorCode MPI_Comm_rank(MPI_Comm comm, PetscMPIInt *rank)
{
  *rank = PetscB200CommRank(comm);
  return PETSC_SUCCESS;
}
PetscErrorCode MPI_Comm_size(MPI_Comm comm, PetscMPIInt *size)
{
  *size = PetscB200CommSize(comm);
  return PETSC_SUCCESS;
}

PetscErrorCode PetscSplitOwnership(MPI_Comm comm, PetscInt *n, PetscInt *N)
{
  int size = PetscB200CommSize(comm), rank = PetscB200CommRank(comm);
  PetscCheck(*n != PETSC_DECIDE || *N != PETSC_DECIDE, 0, PETSC_ERR_ARG_INCOMP, "Both n and N cannot be PETSC_DECIDE");
  if (*N == PETSC_DECIDE) {
    double v = (double)*n;
    PetscCall(PetscB200AllreduceHost(comm, &v, 1, 0));
    *N = (PetscInt)v;
  } else if (*n == PETSC_DECIDE) {
    *n = *N / size + ((*N % size) > rank); /* psplit.c:88 */
  }
  return PETSC_SUCCESS;
}

/* host scalars through the device: used only at setup time and for the <=32-double reductions of MDot/Norm on the
   host-synchronising paths that did not already reduce on the device */
PetscErrorCode PetscB200AllreduceHost(MPI_Comm comm, double *vals, int n, int op)
{
  if (PetscB200CommSize(comm) == 1 || n == 0) return PETSC_SUCCESS;
  PetscCall(PetscB200EnsureInit());
  /* persistent device scratch (grown on demand, never freed per call): a cudaMalloc/cudaFree pair synchronises the whole
     device and would defeat the halo-stream overlap on the CG path (2 dots + 1 norm per iteration) */
  static double *d     = NULL;
  static size_t  d_cap = 0;
  if ((size_t)n > d_cap) {
    if (d) PetscCallB200(b200Free(PetscB200.h, d));
    d_cap = (size_t)n < 1024 ? 1024 : (size_t)n;
    PetscCallB200(b200Malloc(PetscB200.h, (void **)&d, sizeof(double) * d_cap));
  }
  PetscCallB200(b200MemcpyHtoDAsync(PetscB200.h, d, vals, sizeof(double) * (size_t)n));
  if (op == 0) PetscCallB200(b200CommAllreduceSum(PetscB200.h, d, n));
  else PetscCallB200(b200CommAllreduceMax(PetscB200.h, d, n));
  PetscCallB200(b200MemcpyDtoH(PetscB200.h, vals, d, sizeof(double) * (size_t)n));
  return PETSC_SUCCESS;
}

PetscErrorCode PetscB200AllgatherInt64(MPI_Comm comm, int64_t mine, int64_t *all)
{
  int     size = PetscB200CommSize(comm), rank = PetscB200CommRank(comm);
  double *v = (double *)calloc((size_t)size, sizeof(double));
  v[rank]   = (double)mine; /* exact below 2^53 */
  PetscCall(PetscB200AllreduceHost(comm, v, size, 0));
  for (int i = 0; i < size; i++) all[i] = (int64_t)v[i];
  free(v);
  return PETSC_SUCCESS;
}

/* sendbuf: concatenated per-destination segments in rank order. recvcounts is filled (all-reduced count matrix) and
   *recvbuf allocated with the concatenated per-source segments */
PetscErrorCode PetscB200AlltoallvInt(MPI_Comm comm, const int *sendcounts, const int *sendbuf, int *recvcounts, int **recvbuf)
{
  int size = PetscB200CommSize(comm), rank = PetscB200CommRank(comm);
  *recvbuf = NULL;
  if (size == 1) {
    recvcounts[0] = 0;
    return PETSC_SUCCESS;
  }
  double *mat = (double *)calloc((size_t)size * size, sizeof(double));
  for (int p = 0; p < size; p++) mat[(size_t)rank * size + p] = sendcounts[p];
  PetscCall(PetscB200AllreduceHost(comm, mat, size * size, 0));
  size_t ns = 0, nr = 0;
  for (int p = 0; p < size; p++) {
    recvcounts[p] = (int)mat[(size_t)p * size + rank];
    ns += (size_t)sendcounts[p];
    nr += (size_t)recvcounts[p];
  }
  free(mat);
  int *d_s = NULL, *d_r = NULL;
  PetscCallB200(b200Malloc(PetscB200.h, (void **)&d_s, sizeof(int) * (ns + 1)));
  PetscCallB200(b200Malloc(PetscB200.h, (void **)&d_r, sizeof(int) * (nr + 1)));
  if (ns) PetscCallB200(b200MemcpyHtoDAsync(PetscB200.h, d_s, sendbuf, sizeof(int) * ns));
  PetscCallB200(b200CommAlltoallvInt(PetscB200.h, sendcounts, d_s, recvcounts, d_r));
  *recvbuf = (int *)malloc(sizeof(int) * (nr + 1));
  if (nr) PetscCallB200(b200MemcpyDtoH(PetscB200.h, *recvbuf, d_r, sizeof(int) * nr));
  else PetscCallB200(b200Synchronize(PetscB200.h));
  PetscCallB200(b200Free(PetscB200.h, d_s));
  PetscCallB200(b200Free(PetscB200.h, d_r));
  return PETSC_SUCCESS;
}
