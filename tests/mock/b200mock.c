/*
 * b200mock.c -- a TEST DOUBLE of libpetscb200.so for the build container (no GPU): the same C ABI with "device memory" that is
 * plain malloc'ed host memory and kernels that are sequential loops in the reference's operation order.
 *
 * Purpose: the PETSc plugin (petsc_plugin/petscb200_plugin.c) is ~2500 lines of HOST logic -- offload masks, object states, lazy
 * host arrays, mirror invalidation, sub-classing -- that could otherwise only run on a GPU box.  With
 *     LD_PRELOAD=tests/mock/libb200mock.so  <a PETSc program>  -dll_append libpetscb200plugin.so -mat_type aijb200 -vec_type b200
 * the plugin's calls bind to this file instead of the CUDA library, so its logic runs (and can be valgrind'ed) on the CPU, e.g.
 * the reference's own test programs of tools/ref_conformance.py (tests/test_plugin_logic_mock_cpu.py).
 *
 * It is NOT a fallback: it is never installed next to the product, never linked by it, carries a different file name, and the
 * product library still fails loudly without a GPU (tests/test_abi_cpu.py::test_no_cpu_fallback_without_gpu).  Functions that are
 * pure host code in the real library (b200MpiaijSplitHost, b200MpiaijBuildGarray, b200IndexedGroupHost, b200HostFree) are not
 * defined here and resolve to the real library.  Several ranks = several processes whose collectives go through files in /dev/shm (an all-gather of byte strings).
 * ILU(0) / ICC(0) use the oracle's restatements (oracle/liboracle.so), whose layout is the reference's.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "petscb200.h"
#include "oracle.h"

struct b200Handle_s {
  int dummy;
};
static char g_err[512] = "";
static int  fail(int code, const char *fmt, ...)
{
  va_list ap;
  int     k = snprintf(g_err, sizeof g_err, "[petscb200 mock error %d] ", code);
  va_start(ap, fmt);
  vsnprintf(g_err + k, sizeof g_err - (size_t)k, fmt, ap);
  va_end(ap);
  return code;
}
#define CHECK(c, code, ...) \
  do { \
    if (!(c)) return fail(code, __VA_ARGS__); \
  } while (0)

/* ---- allocation registry: what b200PointerIsDevice answers */
static struct {
  char  *p;
  size_t n;
} *g_alloc;
static int g_nalloc, g_cap;
static long long g_launches, g_h2d, g_d2h;

int         b200Create(b200Handle *h, int device) { (void)device; *h = (b200Handle)calloc(1, sizeof(**h)); return 0; }
int         b200Destroy(b200Handle h) { free(h); return 0; }
int         b200SetStream(b200Handle h, void *s) { (void)h; (void)s; return 0; }
int         b200GetStream(b200Handle h, void **s) { (void)h; *s = NULL; return 0; }
int         b200Synchronize(b200Handle h) { (void)h; return 0; }
int         b200DeviceSynchronize(void) { return 0; }
int         b200GetDevice(b200Handle h, int *d) { (void)h; *d = 0; return 0; }
int         b200DeviceCount(int *n) { *n = 1; return 0; }
const char *b200GetLastErrorString(void) { return g_err; }
const char *b200Version(void) { return "petscb200 MOCK (host test double)"; }
long long   b200KernelLaunchCount(void) { return g_launches; }
int         b200TransferCounters(long long *a, long long *b) { if (a) *a = g_h2d; if (b) *b = g_d2h; return 0; }
int         b200PointerIsDevice(const void *ptr, int *is)
{
  *is = 0;
  for (int i = 0; i < g_nalloc; i++)
    if ((const char *)ptr >= g_alloc[i].p && (const char *)ptr < g_alloc[i].p + g_alloc[i].n) *is = 1;
  return 0;
}
int b200Malloc(b200Handle h, void **d, size_t bytes)
{
  (void)h;
  size_t n = bytes + 256;
  char  *p = (char *)malloc(n);
  CHECK(p, B200_ERR_MEM, "out of memory");
  memset(p, 0xA5, n); /* device memory is not zero-initialised: poison it */
  if (g_nalloc == g_cap) {
    g_cap   = g_cap ? 2 * g_cap : 256;
    g_alloc = realloc(g_alloc, sizeof(*g_alloc) * (size_t)g_cap);
  }
  g_alloc[g_nalloc].p = p;
  g_alloc[g_nalloc].n = n;
  g_nalloc++;
  *d = p;
  return 0;
}
int b200Free(b200Handle h, void *d)
{
  (void)h;
  if (!d) return 0;
  for (int i = 0; i < g_nalloc; i++)
    if (g_alloc[i].p == (char *)d) {
      memset(d, 0x5A, g_alloc[i].n); /* use-after-free shows up as garbage */
      free(d);
      g_alloc[i] = g_alloc[--g_nalloc];
      return 0;
    }
  return fail(B200_ERR_ARG_WRONG, "b200Free of a pointer that b200Malloc did not return");
}
int b200MallocHost(void **p, size_t b) { *p = malloc(b ? b : 1); return 0; }
int b200FreeHost(void *p) { free(p); return 0; }
int b200MallocMapped(void **hp, void **dp, size_t b) { *hp = *dp = calloc(1, b ? b : 1); return 0; }
static int dev(const void *p) { int is; b200PointerIsDevice(p, &is); return is; }
int b200MemcpyHtoD(b200Handle h, void *d, const void *s, size_t b)
{
  (void)h;
  CHECK(!b || dev(d), B200_ERR_ARG_WRONG, "b200MemcpyHtoD: destination is not device memory");
  memcpy(d, s, b); g_h2d += (long long)b;
  return 0;
}
int b200MemcpyDtoH(b200Handle h, void *d, const void *s, size_t b)
{
  (void)h;
  CHECK(!b || dev(s), B200_ERR_ARG_WRONG, "b200MemcpyDtoH: source is not device memory");
  memcpy(d, s, b); g_d2h += (long long)b;
  return 0;
}
int b200MemcpyDtoD(b200Handle h, void *d, const void *s, size_t b) { (void)h; memmove(d, s, b); return 0; }
int b200MemcpyHtoDAsync(b200Handle h, void *d, const void *s, size_t b) { return b200MemcpyHtoD(h, d, s, b); }
int b200MemcpyDtoHAsync(b200Handle h, void *d, const void *s, size_t b) { return b200MemcpyDtoH(h, d, s, b); }
int b200Memset(b200Handle h, void *d, int byte, size_t b) { (void)h; memset(d, byte, b); return 0; }
int b200MemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)1 << 36; return 0; }
#define DEVPTR(p) CHECK(!(p) || dev(p), B200_ERR_ARG_WRONG, "%s: %s is not device memory", __func__, #p)

/* ---- CSR SpMV */
struct b200CsrPlan_s {
  int        m, n;
  int64_t    nnz;
  const int *rp, *ci;
};
int b200CsrPlanCreate(b200Handle h, int m, int n, int64_t nnz, const int *rp, const int *ci, b200CsrPlan *plan)
{
  (void)h;
  DEVPTR(rp); DEVPTR(ci);
  b200CsrPlan p = calloc(1, sizeof(*p));
  p->m = m; p->n = n; p->nnz = nnz; p->rp = rp; p->ci = ci;
  *plan = p;
  return 0;
}
int b200CsrPlanDestroy(b200CsrPlan p) { free(p); return 0; }
int b200CsrPlanSetLayout(b200CsrPlan p, int a, int b, int c, int d) { (void)p; (void)a; (void)b; (void)c; (void)d; return 0; }
int b200CsrPlanSetSummation(b200CsrPlan p, int t) { (void)p; (void)t; return 0; }
int b200CsrPlanSetColumnBlocks(b200Handle h, b200CsrPlan p, int nb) { (void)h; (void)p; (void)nb; return 0; }
int b200CsrPlanPackValues(b200Handle h, b200CsrPlan p, const double *v) { (void)h; (void)p; (void)v; return 0; }
int b200CsrPlanAutoColumnBlocks(b200Handle h, b200CsrPlan p, int *nb) { (void)h; (void)p; if (nb) *nb = 0; return 0; }
int b200CsrPlanSetCacheHints(b200CsrPlan p, int hints) { (void)p; (void)hints; return 0; }
static int spmv(b200CsrPlan p, const double *a, const double *x, const double *y, const double *dinv, double *w, double *yout)
{
  for (int r = 0; r < p->m; r++) {
    double s = y ? y[r] : 0.0;
    for (int k = p->rp[r]; k < p->rp[r + 1]; k++) s += a[k] * x[p->ci[k]];
    if (yout) yout[r] = s;
    if (w) w[r] = dinv ? dinv[r] * s : s;
  }
  g_launches++;
  return 0;
}
int b200CsrSpMV(b200Handle h, b200CsrPlan p, const double *a, const double *x, double *y) { (void)h; DEVPTR(a); DEVPTR(x); DEVPTR(y); return spmv(p, a, x, NULL, NULL, y, NULL); }
int b200CsrSpMVAdd(b200Handle h, b200CsrPlan p, const double *a, const double *x, const double *y, double *z) { (void)h; DEVPTR(a); DEVPTR(x); DEVPTR(y); DEVPTR(z); return spmv(p, a, x, y, NULL, z, NULL); }
int b200CsrSpMVJacobi(b200Handle h, b200CsrPlan p, const double *a, const double *x, const double *dinv, double *w, double *y) { (void)h; DEVPTR(a); DEVPTR(x); DEVPTR(dinv); DEVPTR(w); DEVPTR(y); return spmv(p, a, x, NULL, dinv, w, y); }
int b200CsrSpMVAddCompressed(b200Handle h, int nc, const int *ci, const int *rindex, const int *cj, const double *a, const double *x, const double *y, double *z)
{
  (void)h;
  for (int c = 0; c < nc; c++) {
    const int r = rindex[c];
    double    s = y[r];
    for (int k = ci[c]; k < ci[c + 1]; k++) s += a[k] * x[cj[k]];
    z[r] = s;
  }
  g_launches++;
  return 0;
}
int b200CsrSpMVAddCompressedJacobi(b200Handle h, int nc, const int *ci, const int *rindex, const int *bj, const double *ba, const double *lvec, const int *ai, const int *aj, const double *aa, const double *x, const double *dinv, double *w)
{
  (void)h;
  for (int c = 0; c < nc; c++) {
    const int r = rindex[c];
    double    s = 0.0;
    for (int k = ai[r]; k < ai[r + 1]; k++) s += aa[k] * x[aj[k]];
    for (int k = ci[c]; k < ci[c + 1]; k++) s += ba[k] * lvec[bj[k]];
    w[r] = dinv[r] * s;
  }
  g_launches++;
  return 0;
}
int b200CsrGetDiagonal(b200Handle h, int m, const int *rp, const int *ci, const double *a, double *d, int *pos)
{
  (void)h;
  DEVPTR(rp); DEVPTR(a); DEVPTR(d);
  for (int r = 0; r < m; r++) {
    d[r] = 0.0;
    if (pos) pos[r] = -1;
    for (int k = rp[r]; k < rp[r + 1]; k++)
      if (ci[k] == r) {
        d[r] = a[k];
        if (pos) pos[r] = k;
        break;
      }
  }
  g_launches++;
  return 0;
}
int b200JacobiInvertDiagonal(b200Handle h, int64_t n, const double *d, double *dinv, int *nzero)
{
  (void)h;
  int z = 0;
  for (int64_t i = 0; i < n; i++) {
    if (d[i] == 0.0) { dinv[i] = 1.0; z++; }
    else dinv[i] = 1.0 / d[i];
  }
  if (nzero) *nzero = z;
  g_launches++;
  return 0;
}
int b200CsrSplitColumns(b200Handle h, int m, const int *di, const int *dj, const double *da, int cstart, int cend, int **Ai, int **Aj, double **Aa, int64_t *nzA, int **Bi, int **Bj, double **Ba, int64_t *nzB)
{
  int rc = b200MpiaijSplitHost(m, cstart, cend, di, dj, da, nzA, nzB, NULL, NULL, NULL, NULL, NULL, NULL);
  if (rc) return rc;
  b200Malloc(h, (void **)Ai, sizeof(int) * ((size_t)m + 1)); b200Malloc(h, (void **)Bi, sizeof(int) * ((size_t)m + 1));
  b200Malloc(h, (void **)Aj, sizeof(int) * ((size_t)*nzA + 1)); b200Malloc(h, (void **)Bj, sizeof(int) * ((size_t)*nzB + 1));
  b200Malloc(h, (void **)Aa, sizeof(double) * ((size_t)*nzA + 1)); b200Malloc(h, (void **)Ba, sizeof(double) * ((size_t)*nzB + 1));
  return b200MpiaijSplitHost(m, cstart, cend, di, dj, da, nzA, nzB, *Ai, *Aj, *Aa, *Bi, *Bj, *Ba);
}

/* ---- BLAS-1 */
#define LOOP for (int64_t i = 0; i < n; i++)
#define K1(name, ...) g_launches++; (void)h;
int b200VecSet(b200Handle h, int64_t n, double a, double *x) { K1() DEVPTR(x); LOOP x[i] = a; return 0; }
int b200VecCopy(b200Handle h, int64_t n, const double *x, double *y) { K1() DEVPTR(x); DEVPTR(y); LOOP y[i] = x[i]; return 0; }
int b200VecScale(b200Handle h, int64_t n, double a, double *x) { K1() DEVPTR(x); LOOP x[i] *= a; return 0; }
int b200VecAXPY(b200Handle h, int64_t n, double a, const double *x, double *y) { K1() DEVPTR(x); DEVPTR(y); LOOP y[i] += a * x[i]; return 0; }
int b200VecAYPX(b200Handle h, int64_t n, double a, const double *x, double *y) { K1() DEVPTR(x); DEVPTR(y); LOOP y[i] = x[i] + a * y[i]; return 0; }
int b200VecAXPBY(b200Handle h, int64_t n, double a, double b, const double *x, double *y) { K1() DEVPTR(x); DEVPTR(y); LOOP y[i] = a * x[i] + b * y[i]; return 0; }
int b200VecWAXPY(b200Handle h, int64_t n, double a, const double *x, const double *y, double *w) { K1() DEVPTR(x); DEVPTR(y); DEVPTR(w); LOOP w[i] = a * x[i] + y[i]; return 0; }
int b200VecPointwiseMult(b200Handle h, int64_t n, const double *x, const double *y, double *w) { K1() DEVPTR(x); DEVPTR(y); DEVPTR(w); LOOP w[i] = x[i] * y[i]; return 0; }
int b200VecPointwiseDivide(b200Handle h, int64_t n, const double *x, const double *y, double *w) { K1() DEVPTR(x); DEVPTR(y); DEVPTR(w); LOOP w[i] = (y[i] != 0.0) ? x[i] / y[i] : 0.0; return 0; } /* bvec2.c VecPointwiseDivide_Seq: 0 where y is 0 */
int b200VecReciprocal(b200Handle h, int64_t n, double *x) { K1() DEVPTR(x); LOOP if (x[i] != 0.0) x[i] = 1.0 / x[i]; return 0; }
int b200VecShift(b200Handle h, int64_t n, double s, double *x) { K1() DEVPTR(x); LOOP x[i] += s; return 0; }
int b200VecDot(b200Handle h, int64_t n, const double *x, const double *y, double *r) { K1() DEVPTR(x); DEVPTR(y); double s = 0; LOOP s += x[i] * y[i]; *r = s; return 0; }
int b200VecNorm2(b200Handle h, int64_t n, const double *x, double *r) { K1() DEVPTR(x); double s = 0; LOOP s += x[i] * x[i]; *r = sqrt(s); return 0; }
int b200VecNorm(b200Handle h, int64_t n, const double *x, int type, double *r)
{
  K1() DEVPTR(x);
  double s = 0;
  if (type == 0) LOOP s += fabs(x[i]);
  else if (type == 3) LOOP { if (fabs(x[i]) > s) s = fabs(x[i]); }
  else { LOOP s += x[i] * x[i]; s = sqrt(s); }
  *r = s;
  return 0;
}
int b200VecSum(b200Handle h, int64_t n, const double *x, double *r) { K1() DEVPTR(x); double s = 0; LOOP s += x[i]; *r = s; return 0; }
int b200VecMax(b200Handle h, int64_t n, const double *x, int64_t *idx, double *r)
{
  K1() DEVPTR(x);
  double  v = -INFINITY; int64_t k = -1;
  LOOP if (x[i] > v) { v = x[i]; k = i; }
  if (idx) *idx = k;
  *r = v;
  return 0;
}
int b200VecMin(b200Handle h, int64_t n, const double *x, int64_t *idx, double *r)
{
  K1() DEVPTR(x);
  double  v = INFINITY; int64_t k = -1;
  LOOP if (x[i] < v) { v = x[i]; k = i; }
  if (idx) *idx = k;
  *r = v;
  return 0;
}
int b200VecMDot(b200Handle h, int64_t n, int nv, const double *x, const double *const *y, double *r)
{
  K1() DEVPTR(x);
  for (int j = 0; j < nv; j++) { DEVPTR(y[j]); double s = 0; LOOP s += x[i] * y[j][i]; r[j] = s; }
  return 0;
}
int b200VecMDotAsync(b200Handle h, int64_t n, int nv, const double *x, const double *const *y, double *dr) { return b200VecMDot(h, n, nv, x, y, dr); } /* dr: device or mapped pinned memory */
int b200VecMAXPYAsync(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *x, double *sumsq)
{
  K1() DEVPTR(x);
  for (int j = 0; j < nv; j++) DEVPTR(y[j]);
  ora_vecmaxpy(n, nv, alpha, y, x); /* the reference's association (dvec2.c:658-693): remainder group first, then groups of 4 */
  if (sumsq) { double s = 0; LOOP s += x[i] * x[i]; *sumsq = s; }
  return 0;
}
int b200VecMAXPY(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *x, double *norm2)
{
  double ss = 0;
  int    rc = b200VecMAXPYAsync(h, n, nv, alpha, y, x, norm2 ? &ss : NULL);
  if (norm2) *norm2 = sqrt(ss);
  return rc;
}
int b200VecAXPYDot(b200Handle h, int64_t n, double a, const double *x, double *y, const double *z, double *r)
{
  K1() double s = 0;
  LOOP y[i] += a * x[i];
  LOOP s += y[i] * z[i];
  *r = s;
  return 0;
}
int b200VecPipeCGUpdate(b200Handle h, int64_t n, double alpha, double beta, int first, const double *vn, const double *vm, double *u, double *w, double *z, double *q, double *p, double *s, double *x, double *r)
{
  K1()
  LOOP {
    if (first) { z[i] = vn[i]; q[i] = vm[i]; p[i] = u[i]; s[i] = w[i]; }
    else { z[i] = vn[i] + beta * z[i]; q[i] = vm[i] + beta * q[i]; p[i] = u[i] + beta * p[i]; s[i] = w[i] + beta * s[i]; }
    x[i] += alpha * p[i]; u[i] -= alpha * q[i]; w[i] -= alpha * z[i]; r[i] -= alpha * s[i];
  }
  return 0;
}

/* ---- ILU(0) / ICC(0): the oracle's restatements */
struct b200IluPlan_s {
  int     n;
  int    *bi, *bj, *bdiag, *ai, *aj;
  double *ba;
  int     factored;
};
int b200Ilu0Symbolic(b200Handle h, int n, const int *ai, const int *aj, b200IluPlan *plan)
{
  (void)h;
  b200IluPlan p = calloc(1, sizeof(*p));
  const int   nz = ai[n];
  p->n = n;
  p->ai = malloc(sizeof(int) * ((size_t)n + 1)); p->aj = malloc(sizeof(int) * ((size_t)nz + 1));
  memcpy(p->ai, ai, sizeof(int) * ((size_t)n + 1)); memcpy(p->aj, aj, sizeof(int) * (size_t)nz);
  p->bi = malloc(sizeof(int) * ((size_t)n + 1)); p->bj = malloc(sizeof(int) * ((size_t)nz + 1)); p->bdiag = malloc(sizeof(int) * ((size_t)n + 1));
  p->ba = malloc(sizeof(double) * ((size_t)nz + 1));
  int rc = ora_ilu0_symbolic(n, ai, aj, p->bi, p->bj, p->bdiag);
  CHECK(!rc, B200_ERR_MAT_LU_ZRPVT, "ILU(0) symbolic: missing diagonal in row %d", rc - 1);
  *plan = p;
  return 0;
}
int b200Ilu0Destroy(b200IluPlan p) { if (p) { free(p->ai); free(p->aj); free(p->bi); free(p->bj); free(p->bdiag); free(p->ba); free(p); } return 0; }
int b200Ilu0Numeric(b200Handle h, b200IluPlan p, const double *aval, double zeropivot, double shiftamount, int *nshift)
{
  (void)h;
  DEVPTR(aval);
  int rc = ora_lu_numeric(p->n, p->ai, p->aj, aval, p->bi, p->bj, p->bdiag, p->ba, zeropivot, shiftamount);
  if (nshift) *nshift = rc > 0 ? rc : 0;
  if (rc < 0) return fail(B200_ERR_MAT_LU_ZRPVT, "ILU(0): zero pivot (oracle code %d)", rc);
  p->factored = 1;
  return 0;
}
int b200Ilu0Solve(b200Handle h, b200IluPlan p, const double *b, double *x)
{
  (void)h;
  DEVPTR(b); DEVPTR(x);
  CHECK(p->factored, B200_ERR_ORDER, "ILU(0) solve before the numeric factorisation");
  double *t = malloc(sizeof(double) * (size_t)(p->n ? p->n : 1));
  ora_matsolve_natural(p->n, p->bi, p->bj, p->bdiag, p->ba, b, t);
  memcpy(x, t, sizeof(double) * (size_t)p->n);
  free(t);
  g_launches += 2;
  return 0;
}
int b200Ilu0GetInfo(b200IluPlan p, int *a, int *b, int64_t *nnz) { if (a) *a = 1; if (b) *b = 1; if (nnz) *nnz = p->ai[p->n]; return 0; }
struct b200IccPlan_s {
  int     n, bad;
  int    *ui, *uj, *udiag, *ai, *aj;
  double *ua;
  int     factored;
};
int b200Icc0Symbolic(b200Handle h, int n, const int *ai, const int *aj, b200IccPlan *plan)
{
  (void)h;
  b200IccPlan p = calloc(1, sizeof(*p));
  const int   nz = ai[n];
  p->n = n;
  p->ai = malloc(sizeof(int) * ((size_t)n + 1)); p->aj = malloc(sizeof(int) * ((size_t)nz + 1));
  memcpy(p->ai, ai, sizeof(int) * ((size_t)n + 1)); memcpy(p->aj, aj, sizeof(int) * (size_t)nz);
  p->ui = malloc(sizeof(int) * ((size_t)n + 1)); p->uj = malloc(sizeof(int) * ((size_t)nz + 1)); p->udiag = malloc(sizeof(int) * ((size_t)n + 1));
  p->ua = malloc(sizeof(double) * ((size_t)nz + 1));
  int rc = ora_icc0_symbolic(n, ai, aj, p->ui, p->uj, p->udiag);
  CHECK(!rc, B200_ERR_MAT_LU_ZRPVT, "ICC(0) symbolic: missing diagonal in row %d", rc - 1);
  *plan = p;
  return 0;
}
int b200Icc0Destroy(b200IccPlan p) { if (p) { free(p->ai); free(p->aj); free(p->ui); free(p->uj); free(p->udiag); free(p->ua); free(p); } return 0; }
int b200Icc0Numeric(b200Handle h, b200IccPlan p, const double *aval, double zeropivot, int *bad)
{
  (void)h;
  DEVPTR(aval);
  int rc = ora_icc0_numeric(p->n, p->ai, p->aj, aval, p->ui, p->uj, p->udiag, p->ua, zeropivot);
  p->bad = rc < 0 ? -rc : 0;
  if (bad) *bad = p->bad;
  p->factored = 1;
  return 0;
}
int b200Icc0Solve(b200Handle h, b200IccPlan p, const double *b, double *x)
{
  (void)h;
  DEVPTR(b); DEVPTR(x);
  CHECK(p->factored && !p->bad, B200_ERR_ORDER, "ICC(0) solve without a valid factorisation");
  double *t = malloc(sizeof(double) * (size_t)(p->n ? p->n : 1));
  ora_matsolve_icc(p->n, p->ui, p->uj, p->udiag, p->ua, b, t);
  memcpy(x, t, sizeof(double) * (size_t)p->n);
  free(t);
  g_launches += 2;
  return 0;
}
int b200Icc0GetInfo(b200IccPlan p, int64_t *nz, int *a, int *b, int *c) { if (nz) *nz = p->ui[p->n]; if (a) *a = 1; if (b) *b = 1; if (c) *c = 1; return 0; }

/* ---- COO from the reference's maps */
struct b200CooPlan_s {
  int64_t  nnz, atot;
  int64_t *jmap, *perm;
};
int b200CooPlanCreateFromMaps(b200Handle h, int64_t nnz, int64_t atot, const int64_t *jmap, const int64_t *perm, b200CooPlan *plan)
{
  (void)h;
  b200CooPlan p = calloc(1, sizeof(*p));
  p->nnz = nnz; p->atot = atot;
  p->jmap = malloc(sizeof(int64_t) * ((size_t)nnz + 1)); p->perm = malloc(sizeof(int64_t) * ((size_t)atot + 1));
  memcpy(p->jmap, jmap, sizeof(int64_t) * ((size_t)nnz + 1)); memcpy(p->perm, perm, sizeof(int64_t) * (size_t)atot);
  *plan = p;
  return 0;
}
int b200CooPlanDestroy(b200CooPlan p) { if (p) { free(p->jmap); free(p->perm); free(p); } return 0; }
int b200CooSetValues(b200Handle h, b200CooPlan p, const double *v, int insert, double *a)
{
  (void)h;
  DEVPTR(v); DEVPTR(a);
  for (int64_t q = 0; q < p->nnz; q++) {
    double s = 0.0;
    for (int64_t k = p->jmap[q]; k < p->jmap[q + 1]; k++) s += v[p->perm[k]];
    a[q] = (insert ? 0.0 : a[q]) + s;
  }
  g_launches++;
  return 0;
}

/* ---- transposed product */
struct b200CsrTranspose_s {
  int           m, n;
  const int    *rp, *ci;
  const double *a;
  struct b200CsrPlan_s plan;
};
int b200CsrTransposeCreate(b200Handle h, int m, int n, int64_t nnz, const int *rp, const int *ci, b200CsrTranspose *T)
{
  (void)h; (void)nnz;
  b200CsrTranspose t = calloc(1, sizeof(*t));
  t->m = m; t->n = n; t->rp = rp; t->ci = ci;
  *T = t;
  return 0;
}
int b200CsrTransposeDestroy(b200CsrTranspose T) { free(T); return 0; }
int b200CsrTransposeSetValues(b200Handle h, b200CsrTranspose T, const double *a) { (void)h; DEVPTR(a); T->a = a; return 0; }
int b200CsrTransposeSpMV(b200Handle h, b200CsrTranspose T, const double *x, const double *z, double *y)
{
  (void)h;
  DEVPTR(x); DEVPTR(y);
  double *t = malloc(sizeof(double) * (size_t)(T->n ? T->n : 1));
  for (int c = 0; c < T->n; c++) t[c] = z ? z[c] : 0.0;
  for (int r = 0; r < T->m; r++)
    for (int k = T->rp[r]; k < T->rp[r + 1]; k++) t[T->ci[k]] += T->a[k] * x[r];
  memcpy(y, t, sizeof(double) * (size_t)T->n);
  free(t);
  g_launches++;
  return 0;
}
int b200CsrTransposeGetPlan(b200CsrTranspose T, b200CsrPlan *plan) { *plan = &T->plan; return 0; }

/* ---- indexed scatter with op (PetscSF local part), sequential = the reference's loop */
struct b200IndexedPlan_s {
  int64_t n;
  int    *s, *d;
};
int b200IndexedPlanCreate(b200Handle h, int64_t n, const int *sidx, int s0, const int *didx, int d0, b200IndexedPlan *plan)
{
  (void)h;
  b200IndexedPlan p = calloc(1, sizeof(*p));
  p->n = n;
  p->s = malloc(sizeof(int) * ((size_t)n + 1)); p->d = malloc(sizeof(int) * ((size_t)n + 1));
  for (int64_t i = 0; i < n; i++) {
    p->s[i] = sidx ? sidx[i] : s0 + (int)i;
    p->d[i] = didx ? didx[i] : d0 + (int)i;
    CHECK(p->s[i] >= 0 && p->d[i] >= 0, B200_ERR_ARG_OUTOFRANGE, "negative index");
  }
  *plan = p;
  return 0;
}
int b200IndexedPlanDestroy(b200Handle h, b200IndexedPlan p) { (void)h; if (p) { free(p->s); free(p->d); free(p); } return 0; }
int b200IndexedOp(b200Handle h, b200IndexedPlan p, int dtype, int bs, int op, const void *src, void *dst)
{
  (void)h;
  if (!p->n) return 0;
  DEVPTR(src); DEVPTR(dst);
  CHECK(src != dst, B200_ERR_SUP, "in-place indexed operation");
  CHECK(op >= 0 && op <= 4, B200_ERR_SUP, "unknown op");
  if (dtype == B200_SF_F64) ora_sf_scatter_f64(p->n, bs, op, p->s, p->d, (const double *)src, (double *)dst);
  else if (dtype == B200_SF_I32) ora_sf_scatter_i32(p->n, bs, op, p->s, p->d, (const int *)src, (int *)dst);
  else return fail(B200_ERR_SUP, "unknown dtype");
  g_launches++;
  return 0;
}

/* ---- events (timing is meaningless here: 1 ms per interval) and the benchmark generators (oracle generators: same operators) */
struct b200Event_s {
  int dummy;
};
int b200EventCreate(b200Event *ev) { *ev = calloc(1, sizeof(**ev)); return 0; }
int b200EventDestroy(b200Event ev) { free(ev); return 0; }
int b200EventRecord(b200Handle h, b200Event ev) { (void)h; (void)ev; return 0; }
int b200EventSynchronize(b200Event ev) { (void)ev; return 0; }
int b200EventElapsedMs(b200Event a, b200Event b, double *ms) { (void)a; (void)b; *ms = 1.0; return 0; }
int b200GenLaplace7Nnz(int nx, int ny, int nz, int64_t r0, int64_t r1, int64_t *nnz) { *nnz = ora_lap7_rows_nnz(nx, ny, nz, r0, r1); return 0; }
int b200GenLaplace7(b200Handle h, int nx, int ny, int nz, int64_t r0, int64_t r1, int *rowptr, int *colidx, double *val)
{
  (void)h;
  DEVPTR(rowptr); DEVPTR(colidx); DEVPTR(val);
  ora_lap7_rows(nx, ny, nz, r0, r1, rowptr, colidx, val); /* local row pointer, global columns */
  g_launches++;
  return 0;
}
int b200GenLaplace27Nnz(int n, int64_t *nnz) { *nnz = ora_lap27_nnz(n); return 0; }
int b200GenLaplace27(b200Handle h, int n, int *rowptr, int *colidx, double *val)
{
  (void)h;
  DEVPTR(rowptr); DEVPTR(colidx); DEVPTR(val);
  ora_lap27(n, rowptr, colidx, val);
  g_launches++;
  return 0;
}

/* ---- several ranks = several processes: collectives through files in /dev/shm (an all-gather of byte strings; every collective
   below is built on it).  Correctness only; the order of additions is by rank / plan order like the real library's. */
#include <sys/stat.h>
#include <unistd.h>
static int  g_rank = 0, g_nranks = 1;
static long g_seq  = 0;
static char g_dir[256];
static int allgather_bytes(const void *mine, size_t bytes, char ***all, size_t **sizes)
{
  char path[512], tmp[512];
  snprintf(tmp, sizeof tmp, "%s/%ld_%d.tmp", g_dir, g_seq, g_rank);
  snprintf(path, sizeof path, "%s/%ld_%d", g_dir, g_seq, g_rank);
  FILE *f = fopen(tmp, "wb");
  CHECK(f, B200_ERR_LIB, "mock collective: cannot write %s", tmp);
  if (bytes) fwrite(mine, 1, bytes, f);
  fclose(f);
  rename(tmp, path);
  *all   = calloc((size_t)g_nranks, sizeof(char *));
  *sizes = calloc((size_t)g_nranks, sizeof(size_t));
  for (int p = 0; p < g_nranks; p++) {
    struct stat st;
    int         spins = 0;
    snprintf(path, sizeof path, "%s/%ld_%d", g_dir, g_seq, p);
    while (stat(path, &st) != 0) {
      usleep(100);
      CHECK(++spins < 600000, B200_ERR_LIB, "mock collective %ld: rank %d never arrived", g_seq, p); /* 60 s */
    }
    (*sizes)[p] = (size_t)st.st_size;
    (*all)[p]   = malloc((size_t)st.st_size + 1);
    f           = fopen(path, "rb");
    if (st.st_size) { size_t got = fread((*all)[p], 1, (size_t)st.st_size, f); (void)got; }
    fclose(f);
  }
  if (g_seq >= 2) { /* everybody has finished collective seq-2 once all files of seq-1 exist, which this rank has seen */
    snprintf(path, sizeof path, "%s/%ld_%d", g_dir, g_seq - 2, g_rank);
    unlink(path);
  }
  g_seq++;
  return 0;
}
static void free_all(char **all, size_t *sizes)
{
  for (int p = 0; p < g_nranks; p++) free(all[p]);
  free(all);
  free(sizes);
}
int b200CommGetUniqueId(void *id128)
{
  unsigned char *b = id128;
  FILE          *f = fopen("/dev/urandom", "rb");
  if (f) { size_t got = fread(b, 1, 128, f); (void)got; fclose(f); }
  return 0;
}
int b200CommInitRank(b200Handle h, int n, int r, const void *id)
{
  (void)h;
  const unsigned char *b = id;
  g_nranks = n; g_rank = r; g_seq = 0;
  snprintf(g_dir, sizeof g_dir, "/dev/shm/b200mock_%02x%02x%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
  mkdir(g_dir, 0700);
  return 0;
}
int b200CommDestroy(b200Handle h) { (void)h; return 0; }
int b200CommRank(b200Handle h, int *r, int *n) { (void)h; *r = g_rank; *n = g_nranks; return 0; }
int b200CommBarrier(b200Handle h)
{
  (void)h;
  if (g_nranks == 1) return 0;
  char **all; size_t *sz;
  int    rc = allgather_bytes("", 0, &all, &sz);
  if (!rc) free_all(all, sz);
  return rc;
}
static int allreduce(double *buf, int count, int max)
{
  if (g_nranks == 1) return 0;
  char **all; size_t *sz;
  int    rc = allgather_bytes(buf, sizeof(double) * (size_t)count, &all, &sz);
  if (rc) return rc;
  for (int k = 0; k < count; k++) {
    double v = ((double *)all[0])[k];
    for (int p = 1; p < g_nranks; p++) {
      const double w = ((double *)all[p])[k];
      v = max ? (w > v ? w : v) : v + w;
    }
    buf[k] = v;
  }
  free_all(all, sz);
  return 0;
}
int b200CommAllreduceSum(b200Handle h, double *b, int c) { (void)h; return allreduce(b, c, 0); }
int b200CommAllreduceMax(b200Handle h, double *b, int c) { (void)h; return allreduce(b, c, 1); }
struct b200Halo_s {
  int      m, ec, *garray; /* mine */
  int64_t *ranges;         /* [nranks + 1] */
  int    **pg, *pec;       /* every rank's garray */
  char   **red; size_t *redsz; /* lvecs in flight of a reverse scatter */
};
int b200HaloCreateFromGarray(b200Handle h, int m, int ec, const int *g, int64_t *ranges, b200Halo *halo)
{
  (void)h;
  if (g_nranks == 1) {
    CHECK(ec == 0, B200_ERR_ARG_OUTOFRANGE, "garray holds columns that are locally owned or outside the global range");
    ranges[0] = 0; ranges[1] = m;
    *halo = NULL;
    return 0;
  }
  b200Halo H = calloc(1, sizeof(*H));
  char **all; size_t *sz;
  int    rc = allgather_bytes(&m, sizeof(int), &all, &sz);
  if (rc) return rc;
  ranges[0] = 0;
  for (int p = 0; p < g_nranks; p++) ranges[p + 1] = ranges[p] + *(int *)all[p];
  free_all(all, sz);
  H->m = m; H->ec = ec;
  H->garray = malloc(sizeof(int) * ((size_t)ec + 1)); memcpy(H->garray, g, sizeof(int) * (size_t)ec);
  H->ranges = malloc(sizeof(int64_t) * ((size_t)g_nranks + 1)); memcpy(H->ranges, ranges, sizeof(int64_t) * ((size_t)g_nranks + 1));
  rc = allgather_bytes(g, sizeof(int) * (size_t)ec, &all, &sz);
  if (rc) return rc;
  H->pg = calloc((size_t)g_nranks, sizeof(int *)); H->pec = calloc((size_t)g_nranks, sizeof(int));
  for (int p = 0; p < g_nranks; p++) {
    H->pec[p] = (int)(sz[p] / sizeof(int));
    H->pg[p]  = malloc(sz[p] + 4); memcpy(H->pg[p], all[p], sz[p]);
  }
  free_all(all, sz);
  for (int q = 0; q < ec; q++) CHECK(g[q] < ranges[g_rank] || g[q] >= ranges[g_rank + 1], B200_ERR_ARG_OUTOFRANGE, "garray holds a locally owned column");
  *halo = H;
  return 0;
}
int b200HaloDestroy(b200Halo H)
{
  if (!H) return 0;
  for (int p = 0; p < g_nranks; p++) free(H->pg[p]);
  free(H->pg); free(H->pec); free(H->garray); free(H->ranges); free(H);
  return 0;
}
int b200HaloBegin(b200Handle h, b200Halo H, const double *x, double *lvec)
{
  (void)h;
  if (!H) return 0;
  char **all; size_t *sz;
  int    rc = allgather_bytes(x, sizeof(double) * (size_t)H->m, &all, &sz);
  if (rc) return rc;
  for (int q = 0, p = 0; q < H->ec; q++) {
    while (H->garray[q] >= H->ranges[p + 1]) p++;
    lvec[q] = ((double *)all[p])[H->garray[q] - H->ranges[p]];
  }
  free_all(all, sz);
  g_launches++;
  return 0;
}
int b200HaloEnd(b200Handle h, b200Halo H) { (void)h; (void)H; return 0; }
int b200HaloReduceBegin(b200Handle h, b200Halo H, const double *lvec)
{
  (void)h;
  if (!H) return 0;
  return allgather_bytes(lvec, sizeof(double) * (size_t)H->ec, &H->red, &H->redsz);
}
int b200HaloReduceEnd(b200Handle h, b200Halo H, double *y)
{
  (void)h;
  if (!H) return 0;
  for (int p = 0; p < g_nranks; p++) { /* peer after peer in plan order, a peer's entries in its garray order */
    if (p == g_rank) continue;
    for (int q = 0; q < H->pec[p]; q++) {
      const int c = H->pg[p][q];
      if (c >= H->ranges[g_rank] && c < H->ranges[g_rank + 1]) y[c - H->ranges[g_rank]] += ((double *)H->red[p])[q];
    }
  }
  free_all(H->red, H->redsz);
  H->red = NULL;
  g_launches++;
  return 0;
}
