/*
 * oracle.c -- CPU restatement of the reference's Krylov hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product library.
 *
 * Compiled with -O2 -ffp-contract=off and WITHOUT -march=native: the reference's -O2 x86-64 build
 * has no FMA contraction and uses the generic (non-AVX512) PetscSparseDensePlusDot branch
 * (aij.h:609-614), so the arithmetic below is bit-for-bit what MatMult_SeqAIJ etc. perform.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* CPU-baseline timing only: when set, the elementwise vector loops below also run under OpenMP (results of elementwise
   loops are unchanged by threading; reductions use the *_omp variants explicitly) */
static int g_omp = 0;
#define OMP_FOR _Pragma("omp parallel for schedule(static) if (g_omp && n > 65536)")

void ora_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int ora_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ generators */
int64_t ora_lap5_nnz(int m, int n)
{
  return (int64_t)5 * m * n - 2 * (int64_t)m - 2 * (int64_t)n;
}

/* ex2.c:70-92: Ii = i*n + j, i in [0,m), j in [0,n); neighbours Ii-n (i>0), Ii+n (i<m-1), Ii-1 (j>0), Ii+1 (j<n-1).
   Columns emitted sorted (MatAssemblyEnd_SeqAIJ keeps rows sorted). */
void ora_lap5(int m, int n, int *ai, int *aj, double *aa)
{
  int64_t k = 0;
  ai[0] = 0;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      int Ii = i * n + j;
      if (i > 0) { aj[k] = Ii - n; aa[k++] = -1.0; }
      if (j > 0) { aj[k] = Ii - 1; aa[k++] = -1.0; }
      aj[k] = Ii; aa[k++] = 4.0;
      if (j < n - 1) { aj[k] = Ii + 1; aa[k++] = -1.0; }
      if (i < m - 1) { aj[k] = Ii + n; aa[k++] = -1.0; }
      ai[Ii + 1] = (int)k;
    }
}

int64_t ora_lap7_rows_nnz(int nx, int ny, int nz, int64_t r0, int64_t r1)
{
  int64_t k = 0, nxy = (int64_t)nx * ny;
  for (int64_t r = r0; r < r1; r++) {
    int x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / nxy);
    k += 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
  }
  return k;
}

int64_t ora_lap7_nnz(int nx, int ny, int nz)
{
  int64_t N = (int64_t)nx * ny * nz;
  return 7 * N - 2 * ((int64_t)nx * ny + (int64_t)ny * nz + (int64_t)nx * nz);
}

void ora_lap7_rows(int nx, int ny, int nz, int64_t r0, int64_t r1, int *ai, int *aj, double *aa)
{
  int64_t k = 0, nxy = (int64_t)nx * ny;
  ai[0] = 0;
  for (int64_t r = r0; r < r1; r++) {
    int x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / nxy);
    if (z > 0) { aj[k] = (int)(r - nxy); aa[k++] = -1.0; }
    if (y > 0) { aj[k] = (int)(r - nx); aa[k++] = -1.0; }
    if (x > 0) { aj[k] = (int)(r - 1); aa[k++] = -1.0; }
    aj[k] = (int)r; aa[k++] = 6.0;
    if (x < nx - 1) { aj[k] = (int)(r + 1); aa[k++] = -1.0; }
    if (y < ny - 1) { aj[k] = (int)(r + nx); aa[k++] = -1.0; }
    if (z < nz - 1) { aj[k] = (int)(r + nxy); aa[k++] = -1.0; }
    ai[r - r0 + 1] = (int)k;
  }
}

void ora_lap7(int nx, int ny, int nz, int *ai, int *aj, double *aa)
{
  ora_lap7_rows(nx, ny, nz, 0, (int64_t)nx * ny * nz, ai, aj, aa);
}

/* Same operator, filled by all threads (first-touch places each thread's rows on its own NUMA node: only the CPU
   baseline timing uses this; the content is identical to ora_lap7) */
void ora_lap7_omp(int nx, int ny, int nz, int *ai, int *aj, double *aa)
{
  const int64_t N = (int64_t)nx * ny * nz;
#ifdef _OPENMP
  const int nt = omp_get_max_threads();
#else
  const int nt = 1;
#endif
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nt + 1));
  start[0]       = 0;
  for (int t = 0; t < nt; t++) start[t + 1] = start[t] + ora_lap7_rows_nnz(nx, ny, nz, N * t / nt, N * (t + 1) / nt);
  ai[0] = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; t++) { /* thread t writes ai[r0+1 .. r1] and its own slice of aj/aa: no overlap */
    const int64_t r0 = N * t / nt, r1 = N * (t + 1) / nt, nxy = (int64_t)nx * ny;
    int64_t       k = start[t];
    for (int64_t r = r0; r < r1; r++) {
      int x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / nxy);
      if (z > 0) { aj[k] = (int)(r - nxy); aa[k++] = -1.0; }
      if (y > 0) { aj[k] = (int)(r - nx); aa[k++] = -1.0; }
      if (x > 0) { aj[k] = (int)(r - 1); aa[k++] = -1.0; }
      aj[k] = (int)r; aa[k++] = 6.0;
      if (x < nx - 1) { aj[k] = (int)(r + 1); aa[k++] = -1.0; }
      if (y < ny - 1) { aj[k] = (int)(r + nx); aa[k++] = -1.0; }
      if (z < nz - 1) { aj[k] = (int)(r + nxy); aa[k++] = -1.0; }
      ai[r + 1] = (int)k;
    }
  }
  free(start);
}

int64_t ora_lap27_nnz(int n)
{
  /* sum over rows of prod_d (1 + (c_d>0) + (c_d<n-1)) = (3n-2)^3 */
  int64_t t = 3 * (int64_t)n - 2;
  return t * t * t;
}

/* bench_kspsolve.c:115-303: weights by stencil class; emitted in sorted column order */
void ora_lap27(int n, int *ai, int *aj, double *aa)
{
  const double h = 1.0 / (n - 1);
  const double w[4] = {44.0 / 13 * h, -3.0 / 13 * h, -3.0 / 26 * h, -1.0 / 13 * h};
  int64_t      k = 0, n2 = (int64_t)n * n;
  ai[0] = 0;
  for (int z = 0; z < n; z++)
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++) {
        int64_t r = x + (int64_t)n * y + n2 * z;
        for (int dz = -1; dz <= 1; dz++) {
          if (z + dz < 0 || z + dz >= n) continue;
          for (int dy = -1; dy <= 1; dy++) {
            if (y + dy < 0 || y + dy >= n) continue;
            for (int dx = -1; dx <= 1; dx++) {
              if (x + dx < 0 || x + dx >= n) continue;
              aj[k]   = (int)(r + dx + (int64_t)n * dy + n2 * dz);
              aa[k++] = w[(dx != 0) + (dy != 0) + (dz != 0)];
            }
          }
        }
        ai[r + 1] = (int)k;
      }
}

/* ------------------------------------------------------------------ Mat */
void ora_matmult_seqaij(int m, const int *ai, const int *aj, const double *aa, const double *x, double *y)
{
  for (int i = 0; i < m; i++) {
    double sum = 0.0;
    for (int k = ai[i]; k < ai[i + 1]; k++) sum += aa[k] * x[aj[k]];
    y[i] = sum;
  }
}

void ora_matmult_seqaij_omp(int m, const int *ai, const int *aj, const double *aa, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; i++) {
    double sum = 0.0;
    for (int k = ai[i]; k < ai[i + 1]; k++) sum += aa[k] * x[aj[k]];
    y[i] = sum;
  }
}

void ora_matmultadd_seqaij(int m, const int *ai, const int *aj, const double *aa, const double *x, const double *y, double *z)
{
  for (int i = 0; i < m; i++) {
    double sum = y[i];
    for (int k = ai[i]; k < ai[i + 1]; k++) sum += aa[k] * x[aj[k]];
    z[i] = sum;
  }
}

void ora_getdiagonal_seqaij(int m, const int *ai, const int *aj, const double *aa, double *d, int *diagpos)
{
  for (int i = 0; i < m; i++) {
    int pos = -1;
    for (int k = ai[i]; k < ai[i + 1]; k++)
      if (aj[k] == i) { pos = k; break; }
    if (diagpos) diagpos[i] = pos;
    d[i] = pos >= 0 ? aa[pos] : 0.0;
  }
}

/* ------------------------------------------------------------------ Vec */
double ora_vecdot(int64_t n, const double *x, const double *y)
{
  double s = 0.0;
  for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
double ora_vecdot_omp(int64_t n, const double *x, const double *y)
{
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
double ora_vecnorm2(int64_t n, const double *x) { return sqrt(ora_vecdot(n, x, x)); }
void ora_vecaxpy(int64_t n, double a, const double *x, double *y)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
}
void ora_vecaypx(int64_t n, double a, const double *x, double *y)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}
void ora_vecaxpby(int64_t n, double a, double b, const double *x, double *y)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) y[i] = a * x[i] + b * y[i];
}
void ora_vecwaxpy(int64_t n, double a, const double *x, const double *y, double *w)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) w[i] = a * x[i] + y[i];
}
void ora_vecscale(int64_t n, double a, double *x)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) x[i] *= a;
}
void ora_vecpointwisemult(int64_t n, const double *x, const double *y, double *w)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) w[i] = x[i] * y[i];
}
static void vcopy(int64_t n, const double *x, double *y)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) y[i] = x[i];
}
static void vzero(int64_t n, double *x)
{
  OMP_FOR
  for (int64_t i = 0; i < n; i++) x[i] = 0.0;
}
void ora_vecreciprocal(int64_t n, double *x)
{
  for (int64_t i = 0; i < n; i++)
    if (x[i] != 0.0) x[i] = 1.0 / x[i];
}

/* one group of g (1..4) vectors of dvec2.c:83-303: head remainder n&3 handled element by element (highest index first,
   as the fall-through switch does), then quads with the 4-term inner sum added to the accumulator in one step */
static void mdot_group(int64_t n, int g, const double *x, const double *const *y, double *z)
{
  double  s[4] = {0, 0, 0, 0};
  int64_t rem  = n & 3;
  for (int64_t e = rem - 1; e >= 0; e--)
    for (int v = 0; v < g; v++) s[v] += x[e] * y[v][e];
  for (int64_t j = rem; j < n; j += 4) {
    const double x0 = x[j], x1 = x[j + 1], x2 = x[j + 2], x3 = x[j + 3];
    for (int v = 0; v < g; v++) {
      const double *p = y[v] + j;
      s[v] += x0 * p[0] + x1 * p[1] + x2 * p[2] + x3 * p[3];
    }
  }
  for (int v = 0; v < g; v++) z[v] = s[v];
}

void ora_vecmdot(int64_t n, int nv, const double *x, const double *const *y, double *z)
{
  int g = nv & 3;
  if (n == 0) {
    for (int v = 0; v < nv; v++) z[v] = 0.0;
    return;
  }
  if (g) mdot_group(n, g, x, y, z);
  for (int v = g; v < nv; v += 4) mdot_group(n, 4, x, y + v, z + v);
}

void ora_vecmdot_omp(int64_t n, int nv, const double *x, const double *const *y, double *z)
{
  int     nt = ora_max_threads();
  double *part = (double *)calloc((size_t)nt * nv, sizeof(double));
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    int64_t chunk = ((n + nt - 1) / nt + 3) & ~(int64_t)3, lo = t * chunk, hi = lo + chunk;
    if (hi > n) hi = n;
    if (lo < hi) {
      const double *yy[64];
      for (int v = 0; v < nv; v++) yy[v] = y[v] + lo;
      ora_vecmdot(hi - lo, nv, x + lo, yy, part + (size_t)t * nv);
    }
  }
  for (int v = 0; v < nv; v++) {
    double s = 0;
    for (int t = 0; t < nt; t++) s += part[(size_t)t * nv + v];
    z[v] = s;
  }
  free(part);
}

static void maxpy_range(int64_t lo, int64_t hi, int nv, const double *alpha, const double *const *y, double *x)
{
  int g = nv & 3;
  if (g == 3) {
    const double a1 = alpha[0], a2 = alpha[1], a3 = alpha[2];
    for (int64_t i = lo; i < hi; i++) x[i] += a1 * y[0][i] + a2 * y[1][i] + a3 * y[2][i];
  } else if (g == 2) {
    const double a1 = alpha[0], a2 = alpha[1];
    for (int64_t i = lo; i < hi; i++) x[i] += a1 * y[0][i] + a2 * y[1][i];
  } else if (g == 1) {
    const double a1 = alpha[0];
    for (int64_t i = lo; i < hi; i++) x[i] += a1 * y[0][i];
  }
  for (int v = g; v < nv; v += 4) {
    const double  a1 = alpha[v], a2 = alpha[v + 1], a3 = alpha[v + 2], a4 = alpha[v + 3];
    const double *p1 = y[v], *p2 = y[v + 1], *p3 = y[v + 2], *p4 = y[v + 3];
    for (int64_t i = lo; i < hi; i++) x[i] += a1 * p1[i] + a2 * p2[i] + a3 * p3[i] + a4 * p4[i];
  }
}

void ora_vecmaxpy(int64_t n, int nv, const double *alpha, const double *const *y, double *x)
{
  maxpy_range(0, n, nv, alpha, y, x);
}

void ora_vecmaxpy_omp(int64_t n, int nv, const double *alpha, const double *const *y, double *x)
{
  int nt = ora_max_threads();
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    int64_t chunk = (n + nt - 1) / nt, lo = t * chunk, hi = lo + chunk;
    if (hi > n) hi = n;
    if (lo < hi) maxpy_range(lo, hi, nv, alpha, y, x);
  }
}

/* ------------------------------------------------------------------ ILU(0) */
int ora_ilu0_symbolic(int n, const int *ai, const int *aj, int *bi, int *bj, int *bdiag)
{
  int64_t k = 0;
  int    *adiag = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    adiag[i] = -1;
    for (int p = ai[i]; p < ai[i + 1]; p++)
      if (aj[p] == i) { adiag[i] = p; break; }
    if (adiag[i] < 0) { free(adiag); return -(i + 1); } /* MatGetDiagonalMarkers: missing diagonal is an error for ILU */
  }
  bi[0] = 0;
  for (int i = 0; i < n; i++) { /* L part: entries left of the diagonal */
    int nz    = adiag[i] - ai[i];
    bi[i + 1] = bi[i] + nz;
    for (int j = 0; j < nz; j++) bj[k++] = aj[ai[i] + j];
  }
  bdiag[n] = bi[n] - 1;
  for (int i = n - 1; i >= 0; i--) { /* U part, rows stored n-1 .. 0, diagonal last */
    int nz = ai[i + 1] - adiag[i] - 1;
    for (int j = 0; j < nz; j++) bj[k++] = aj[adiag[i] + 1 + j];
    bj[k++]  = i;
    bdiag[i] = bdiag[i + 1] + nz + 1;
  }
  free(adiag);
  return 0;
}

int ora_lu_numeric(int n, const int *ai, const int *aj, const double *aa, const int *bi, const int *bj, const int *bdiag,
                   double *ba, double zeropivot, double shiftamount)
{
  double *rtmp   = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  double  shift  = 0.0;
  int     nshift = 0, newshift;
  do {
    newshift = 0;
    for (int i = 0; i < n; i++) {
      int nzL = bi[i + 1] - bi[i];
      int nzU = bdiag[i] - bdiag[i + 1]; /* includes diagonal */
      for (int j = 0; j < nzL; j++) rtmp[bj[bi[i] + j]] = 0.0;
      for (int j = 0; j < nzU; j++) rtmp[bj[bdiag[i + 1] + 1 + j]] = 0.0;
      for (int p = ai[i]; p < ai[i + 1]; p++) rtmp[aj[p]] = aa[p];
      rtmp[i] += shift;
      for (int k = 0; k < nzL; k++) {
        int     row = bj[bi[i] + k];
        double *pc  = rtmp + row;
        if (*pc != 0.0) {
          double        mult = *pc * ba[bdiag[row]];
          const int    *pj   = bj + bdiag[row + 1] + 1;
          const double *pv   = ba + bdiag[row + 1] + 1;
          int           nz   = bdiag[row] - bdiag[row + 1] - 1;
          *pc = mult;
          for (int j = 0; j < nz; j++) rtmp[pj[j]] -= mult * pv[j];
        }
      }
      double rs = 0.0;
      for (int j = 0; j < nzL; j++) {
        double v      = rtmp[bj[bi[i] + j]];
        ba[bi[i] + j] = v;
        rs += fabs(v);
      }
      for (int j = 0; j < nzU - 1; j++) {
        double v                 = rtmp[bj[bdiag[i + 1] + 1 + j]];
        ba[bdiag[i + 1] + 1 + j] = v;
        rs += fabs(v);
      }
      /* MatPivotCheck_nz, matimpl.h:795-811 */
      if (fabs(rtmp[i]) <= zeropivot * rs && !isnan(rtmp[i])) {
        shift    = nshift ? shift * 2.0 : shiftamount;
        newshift = 1;
        nshift++;
        if (nshift > 60) { free(rtmp); return -1; }
        break;
      }
      ba[bdiag[i]] = 1.0 / rtmp[i];
    }
  } while (newshift);
  free(rtmp);
  return nshift;
}

void ora_matsolve_natural(int n, const int *bi, const int *bj, const int *bdiag, const double *ba, const double *b, double *x)
{
  if (!n) return;
  x[0] = b[0];
  for (int i = 1; i < n; i++) {
    double sum = b[i];
    for (int p = bi[i]; p < bi[i + 1]; p++) sum -= ba[p] * x[bj[p]];
    x[i] = sum;
  }
  for (int i = n - 1; i >= 0; i--) {
    double sum = x[i];
    for (int p = bdiag[i + 1] + 1; p < bdiag[i]; p++) sum -= ba[p] * x[bj[p]];
    x[i] = sum * ba[bdiag[i]];
  }
}

/* ------------------------------------------------------------------ MPIAIJ setup */
void ora_split_ownership(int64_t N, int size, int64_t *rstart)
{
  rstart[0] = 0;
  for (int r = 0; r < size; r++) rstart[r + 1] = rstart[r] + N / size + ((N % size) > r);
}

static int cmp_i64(const void *a, const void *b)
{
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

int ora_mpiaij_split(int mloc, int64_t cstart, int64_t cend, const int *ai, const int64_t *ajg, const double *aa, int *Ai,
                     int *Aj, double *Aa, int *Bi, int *Bj, double *Ba, int64_t *garray)
{
  int64_t nzB = 0, ka = 0, kb = 0;
  int     ec  = 0;
  /* mmaij.c:25-51: collect distinct off-process columns, sort ascending */
  for (int i = 0; i < mloc; i++)
    for (int p = ai[i]; p < ai[i + 1]; p++)
      if (ajg[p] < cstart || ajg[p] >= cend) garray[nzB++] = ajg[p];
  if (nzB) {
    qsort(garray, (size_t)nzB, sizeof(int64_t), cmp_i64);
    ec = 1;
    for (int64_t k = 1; k < nzB; k++)
      if (garray[k] != garray[ec - 1]) garray[ec++] = garray[k];
  }
  Ai[0] = Bi[0] = 0;
  for (int i = 0; i < mloc; i++) {
    for (int p = ai[i]; p < ai[i + 1]; p++) {
      int64_t c = ajg[p];
      if (c >= cstart && c < cend) {
        Aj[ka]   = (int)(c - cstart);
        Aa[ka++] = aa[p];
      } else { /* mmaij.c:55-61: renumber into position within garray */
        int lo = 0, hi = ec - 1;
        while (lo < hi) {
          int mid = (lo + hi) / 2;
          if (garray[mid] < c) lo = mid + 1;
          else hi = mid;
        }
        Bj[kb]   = lo;
        Ba[kb++] = aa[p];
      }
    }
    Ai[i + 1] = (int)ka;
    Bi[i + 1] = (int)kb;
  }
  return ec;
}

/* ------------------------------------------------------------------ KSP */
void ora_ksp_default_opts(ora_ksp_opts *o)
{
  o->pc_type    = ORA_PC_ILU0;
  o->restart    = 30;
  o->cgs_refine = ORA_CGS_REFINE_NEVER;
  o->max_it     = 10000;
  o->rtol       = 1e-5;
  o->abstol     = 1e-50;
  o->dtol       = 1e4;
  o->nblocks    = 1;
  o->use_omp    = 0;
}

typedef struct {
  int           n, type, use_omp;
  const int    *ai, *aj;
  const double *aa;
  double       *dinv;                 /* jacobi: 1/diag */
  int           nblk;                 /* ilu / bjacobi */
  int64_t      *rstart;
  int         **bi, **bj, **bdiag;
  double      **ba;
} ora_pc;

static int pc_setup(ora_pc *pc, int n, const int *ai, const int *aj, const double *aa, const ora_ksp_opts *o)
{
  memset(pc, 0, sizeof(*pc));
  pc->n = n; pc->ai = ai; pc->aj = aj; pc->aa = aa; pc->type = o->pc_type; pc->use_omp = o->use_omp;
  if (o->pc_type == ORA_PC_JACOBI) {
    /* jacobi.c:172-270: diag -> reciprocal; zero diagonal entries become 1.0 */
    pc->dinv = (double *)malloc(sizeof(double) * (size_t)n);
    ora_getdiagonal_seqaij(n, ai, aj, aa, pc->dinv, NULL);
    for (int i = 0; i < n; i++) pc->dinv[i] = pc->dinv[i] != 0.0 ? 1.0 / pc->dinv[i] : 1.0;
  } else if (o->pc_type == ORA_PC_ILU0 || o->pc_type == ORA_PC_BJACOBI_ILU0) {
    int nb     = o->pc_type == ORA_PC_ILU0 ? 1 : o->nblocks;
    pc->nblk   = nb;
    pc->rstart = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb + 1));
    ora_split_ownership(n, nb, pc->rstart);
    pc->bi = (int **)calloc((size_t)nb, sizeof(int *)); pc->bj = (int **)calloc((size_t)nb, sizeof(int *));
    pc->bdiag = (int **)calloc((size_t)nb, sizeof(int *)); pc->ba = (double **)calloc((size_t)nb, sizeof(double *));
    for (int b = 0; b < nb; b++) {
      /* bjacobi.c:117-123: the block is the diagonal block a->A of the row-partitioned matrix */
      int     r0 = (int)pc->rstart[b], r1 = (int)pc->rstart[b + 1], ml = r1 - r0;
      int64_t cap = ai[r1] - ai[r0];
      int    *li = (int *)malloc(sizeof(int) * (size_t)(ml + 1)), *lj = (int *)malloc(sizeof(int) * (size_t)(cap + 1));
      double *la = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
      int64_t k = 0;
      li[0] = 0;
      for (int i = r0; i < r1; i++) {
        for (int p = ai[i]; p < ai[i + 1]; p++)
          if (aj[p] >= r0 && aj[p] < r1) { lj[k] = aj[p] - r0; la[k++] = aa[p]; }
        li[i - r0 + 1] = (int)k;
      }
      pc->bi[b] = (int *)malloc(sizeof(int) * (size_t)(ml + 1)); pc->bdiag[b] = (int *)malloc(sizeof(int) * (size_t)(ml + 1));
      pc->bj[b] = (int *)malloc(sizeof(int) * (size_t)(k + 1)); pc->ba[b] = (double *)calloc((size_t)(k + 1), sizeof(double));
      if (ora_ilu0_symbolic(ml, li, lj, pc->bi[b], pc->bj[b], pc->bdiag[b])) return -1;
      /* ilu.c PCCreate_ILU defaults: shifttype NONZERO, shiftamount 100 eps, zeropivot 100 eps */
      if (ora_lu_numeric(ml, li, lj, la, pc->bi[b], pc->bj[b], pc->bdiag[b], pc->ba[b], 100.0 * 2.220446049250313e-16, 100.0 * 2.220446049250313e-16) < 0) return -2;
      free(li); free(lj); free(la);
    }
  } else if (o->pc_type == ORA_PC_ICC0) {
    /* icc.c PCCreate_ICC defaults: levels 0, natural ordering, shifttype POSITIVE_DEFINITE, zeropivot 100 eps */
    pc->nblk   = 1;
    pc->rstart = (int64_t *)malloc(sizeof(int64_t) * 2);
    pc->rstart[0] = 0; pc->rstart[1] = n;
    pc->bi = (int **)calloc(1, sizeof(int *)); pc->bj = (int **)calloc(1, sizeof(int *));
    pc->bdiag = (int **)calloc(1, sizeof(int *)); pc->ba = (double **)calloc(1, sizeof(double *));
    pc->bi[0] = (int *)malloc(sizeof(int) * ((size_t)n + 1)); pc->bdiag[0] = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    pc->bj[0] = (int *)malloc(sizeof(int) * ((size_t)ai[n] + 1)); pc->ba[0] = (double *)calloc((size_t)ai[n] + 1, sizeof(double));
    if (ora_icc0_symbolic(n, ai, aj, pc->bi[0], pc->bj[0], pc->bdiag[0])) return -1;
    if (ora_icc0_numeric(n, ai, aj, aa, pc->bi[0], pc->bj[0], pc->bdiag[0], pc->ba[0], 100.0 * 2.220446049250313e-16)) return -2;
  }
  return 0;
}

static void pc_apply(const ora_pc *pc, const double *x, double *y)
{
  if (pc->type == ORA_PC_NONE) memcpy(y, x, sizeof(double) * (size_t)pc->n);
  else if (pc->type == ORA_PC_JACOBI) ora_vecpointwisemult(pc->n, x, pc->dinv, y); /* jacobi.c:354-362 */
  else if (pc->type == ORA_PC_ICC0) ora_matsolve_icc(pc->n, pc->bi[0], pc->bj[0], pc->bdiag[0], pc->ba[0], x, y);
  else
    for (int b = 0; b < pc->nblk; b++) {
      int r0 = (int)pc->rstart[b], ml = (int)(pc->rstart[b + 1] - pc->rstart[b]);
      ora_matsolve_natural(ml, pc->bi[b], pc->bj[b], pc->bdiag[b], pc->ba[b], x + r0, y + r0);
    }
}

static void pc_destroy(ora_pc *pc)
{
  free(pc->dinv);
  for (int b = 0; b < pc->nblk; b++) { free(pc->bi[b]); free(pc->bj[b]); free(pc->bdiag[b]); free(pc->ba[b]); }
  free(pc->bi); free(pc->bj); free(pc->bdiag); free(pc->ba); free(pc->rstart);
}

static void k_matmult(const ora_pc *pc, const double *x, double *y)
{
  if (pc->use_omp) ora_matmult_seqaij_omp(pc->n, pc->ai, pc->aj, pc->aa, x, y);
  else ora_matmult_seqaij(pc->n, pc->ai, pc->aj, pc->aa, x, y);
}

/* KSPConvergedDefault, iterativ.c:1490-1581 (zero initial guess: rnorm0 = first residual) */
typedef struct { double rnorm0, ttol, rtol, abstol, dtol; } conv_ctx;
static int converged_default(conv_ctx *c, int it, double rnorm)
{
  if (!it) { c->rnorm0 = rnorm; c->ttol = fmax(c->rtol * rnorm, c->abstol); }
  if (isnan(rnorm) || isinf(rnorm)) return -9;
  if (rnorm <= c->ttol) return rnorm < c->abstol ? 3 : 2;
  if (rnorm >= c->dtol * c->rnorm0) return -4;
  return 0;
}

#define HH(a, b)  (hh + (b) * (max_k + 2) + (a))
#define HES(a, b) (hes + (b) * (max_k + 1) + (a))

int ora_ksp_gmres(int n, const int *ai, const int *aj, const double *aa, const double *b, double *x, const ora_ksp_opts *o,
                  ora_ksp_result *res, double *hist, int histcap)
{
  const int max_k = o->restart;
  ora_pc    pc;
  g_omp = o->use_omp;
  int       rc = pc_setup(&pc, n, ai, aj, aa, o);
  if (rc) return rc;
  double **vv = (double **)malloc(sizeof(double *) * (size_t)(max_k + 1));
  for (int i = 0; i <= max_k; i++) vv[i] = (double *)malloc(sizeof(double) * (size_t)n);
  double *temp = (double *)malloc(sizeof(double) * (size_t)n), *tmat = (double *)malloc(sizeof(double) * (size_t)n);
  double *hh = (double *)calloc((size_t)(max_k + 2) * (max_k + 1), sizeof(double));
  double *hes = (double *)calloc((size_t)(max_k + 1) * (max_k + 1), sizeof(double));
  double *grs = (double *)calloc((size_t)max_k + 2, sizeof(double)), *cc = (double *)calloc((size_t)max_k + 2, sizeof(double));
  double *ss = (double *)calloc((size_t)max_k + 2, sizeof(double)), *nrs = (double *)calloc((size_t)max_k + 2, sizeof(double));
  double *lhh = (double *)calloc((size_t)max_k + 2, sizeof(double));
  conv_ctx cv = {0, 0, o->rtol, o->abstol, o->dtol};
  int      its = 0, reason = 0, nh = 0, guess_zero = 1, itcount = 0;
  double   rnorm = -1.0;
  vzero(n, x);

#define LOGRES(r) do { if (hist && nh < histcap) hist[nh] = (r); nh++; } while (0)
  while (!reason) {
    /* KSPInitialResidual, itres.c:35-73, left PC */
    if (!guess_zero) {
      k_matmult(&pc, x, temp);
      vcopy(n, b, tmat);
      ora_vecaxpy(n, -1.0, temp, tmat);
      pc_apply(&pc, tmat, vv[0]);
    } else {
      vcopy(n, b, tmat);
      pc_apply(&pc, b, vv[0]);
    }
    /* KSPGMRESCycle, gmres.c:88-193 */
    int    it = 0, hapend = 0;
    double resn = ora_vecnorm2(n, vv[0]), tt;
    if (resn != 0.0) ora_vecscale(n, 1.0 / resn, vv[0]);
    grs[0] = resn;
    rnorm  = resn;
    LOGRES(resn);
    if (!resn) { reason = 3; break; }
    reason = converged_default(&cv, its, resn);
    while (!reason && it < max_k && its < o->max_it) {
      if (it) LOGRES(resn);
      /* KSP_PCApplyBAorAB, left: MatMult then PCApply (precon.c:853-854) */
      k_matmult(&pc, vv[it], tmat);
      pc_apply(&pc, tmat, vv[it + 1]);
      /* borthog2.c:33-114 classical Gram-Schmidt */
      {
        double *h = HH(0, it), *he = HES(0, it);
        int     nref = (o->cgs_refine == ORA_CGS_REFINE_ALWAYS) ? 2 : 1;
        for (int j = 0; j <= it; j++) h[j] = he[j] = 0.0;
        for (int pass = 0; pass < nref; pass++) {
          if (pc.use_omp) ora_vecmdot_omp(n, it + 1, vv[it + 1], (const double *const *)vv, lhh);
          else ora_vecmdot(n, it + 1, vv[it + 1], (const double *const *)vv, lhh);
          for (int j = 0; j <= it; j++) lhh[j] = -lhh[j];
          if (pc.use_omp) ora_vecmaxpy_omp(n, it + 1, lhh, (const double *const *)vv, vv[it + 1]);
          else ora_vecmaxpy(n, it + 1, lhh, (const double *const *)vv, vv[it + 1]);
          for (int j = 0; j <= it; j++) { h[j] -= lhh[j]; he[j] -= lhh[j]; }
          if (pass == 0 && o->cgs_refine == ORA_CGS_REFINE_IFNEEDED) {
            double hnrm = 0.0, wnrm;
            for (int j = 0; j <= it; j++) hnrm += lhh[j] * lhh[j];
            hnrm = sqrt(hnrm);
            wnrm = ora_vecnorm2(n, vv[it + 1]);
            if (wnrm < hnrm) nref = 2;
          }
        }
      }
      tt = pc.use_omp ? sqrt(ora_vecdot_omp(n, vv[it + 1], vv[it + 1])) : ora_vecnorm2(n, vv[it + 1]);
      if (tt != 0.0) ora_vecscale(n, 1.0 / tt, vv[it + 1]);
      *HH(it + 1, it)  = tt;
      *HES(it + 1, it) = tt;
      {
        double hapbnd = fabs(tt / grs[it]);
        if (hapbnd > 1.0e-30) hapbnd = 1.0e-30; /* gmres.c:905 haptol */
        if (tt < hapbnd) hapend = 1;
      }
      /* KSPGMRESUpdateHessenberg, gmres.c:346-397 */
      {
        double *h = HH(0, it), t2;
        for (int j = 1; j <= it; j++) {
          t2   = *h;
          *h   = cc[j - 1] * t2 + ss[j - 1] * *(h + 1);
          h++;
          *h = cc[j - 1] * *h - (ss[j - 1] * t2);
        }
        if (!hapend) {
          t2 = sqrt(*h * *h + *(h + 1) * *(h + 1));
          if (t2 == 0.0) { reason = -2; break; }
          cc[it]      = *h / t2;
          ss[it]      = *(h + 1) / t2;
          grs[it + 1] = -(ss[it] * grs[it]);
          grs[it]     = cc[it] * grs[it];
          *h          = cc[it] * *h + ss[it] * *(h + 1);
          resn        = fabs(grs[it + 1]);
        } else resn = 0.0;
      }
      it++;
      its++;
      rnorm  = resn;
      reason = converged_default(&cv, its, resn);
      if (hapend && !reason) { reason = -5; break; }
    }
    /* KSPGMRESBuildSoln, gmres.c:298-341 */
    if (it > 0) {
      int k = it - 1;
      if (*HH(k, k) != 0.0) {
        nrs[k] = grs[k] / *HH(k, k);
        for (int ii = 1; ii <= k; ii++) {
          int    kk = k - ii;
          double t3 = grs[kk];
          for (int j = kk + 1; j <= k; j++) t3 = t3 - *HH(kk, j) * nrs[j];
          if (*HH(kk, kk) == 0.0) { reason = -5; break; }
          nrs[kk] = t3 / *HH(kk, kk);
        }
        vzero(n, temp);
        ora_vecmaxpy(n, it, nrs, (const double *const *)vv, temp);
        ora_vecaxpy(n, 1.0, temp, x);
      } else reason = -5;
    }
    if (!reason && its >= o->max_it) reason = -3;
    if (it && reason) LOGRES(resn);
    itcount += it;
    if (itcount >= o->max_it) { if (!reason) reason = -3; break; }
    guess_zero = 0;
  }
  res->its = its; res->reason = reason; res->rnorm = rnorm; res->nhist = nh;
  for (int i = 0; i <= max_k; i++) free(vv[i]);
  free(vv); free(temp); free(tmat); free(hh); free(hes); free(grs); free(cc); free(ss); free(nrs); free(lhh);
  pc_destroy(&pc);
  return 0;
}

int ora_ksp_cg(int n, const int *ai, const int *aj, const double *aa, const double *b, double *x, const ora_ksp_opts *o,
               ora_ksp_result *res, double *hist, int histcap)
{
  ora_pc pc;
  g_omp = o->use_omp;
  int    rc = pc_setup(&pc, n, ai, aj, aa, o);
  if (rc) return rc;
  double  *R = (double *)malloc(sizeof(double) * (size_t)n), *Z = (double *)malloc(sizeof(double) * (size_t)n);
  double  *P = (double *)malloc(sizeof(double) * (size_t)n), *W = Z;
  conv_ctx cv = {0, 0, o->rtol, o->abstol, o->dtol};
  double   dp, beta, betaold = 1.0, bb = 0.0, a, dpi = 0.0, dpiold;
  int      i = 0, reason = 0, nh = 0, its = 0;
  memset(x, 0, sizeof(double) * (size_t)n);
  memcpy(R, b, sizeof(double) * (size_t)n);      /* r <- b (x = 0), cg.c:163 */
  pc_apply(&pc, R, Z);                             /* z <- Br */
  dp = ora_vecnorm2(n, Z);                         /* KSP_NORM_PRECONDITIONED */
  LOGRES(dp);
  reason = converged_default(&cv, 0, dp);
  if (!reason) {
    beta = ora_vecdot(n, Z, R);
    do {
      its = i + 1;
      if (beta == 0.0) { reason = 3; break; }
      else if (i > 0 && beta * betaold < 0.0) { reason = -8; break; }
      if (!i) { memcpy(P, Z, sizeof(double) * (size_t)n); bb = 0.0; }
      else { bb = beta / betaold; ora_vecaypx(n, bb, Z, P); } /* p <- z + b p */
      dpiold = dpi;
      k_matmult(&pc, P, W);                         /* w <- Ap */
      dpi     = ora_vecdot(n, P, W);
      betaold = beta;
      if (dpi == 0.0 || (i > 0 && ((dpi > 0) - (dpi < 0)) * ((dpiold > 0) - (dpiold < 0)) < 0)) { reason = -10; break; }
      a = beta / dpi;
      ora_vecaxpy(n, a, P, x);
      ora_vecaxpy(n, -a, W, R);
      pc_apply(&pc, R, Z);
      dp = ora_vecnorm2(n, Z);
      LOGRES(dp);
      reason = converged_default(&cv, i + 1, dp);
      if (reason) break;
      beta = ora_vecdot(n, Z, R);
      i++;
    } while (i < o->max_it);
    if (i >= o->max_it && !reason) reason = -3;
  }
  (void)bb;
  res->its = its; res->reason = reason; res->rnorm = dp; res->nhist = nh;
  free(R); free(Z); free(P);
  pc_destroy(&pc);
  return 0;
}

/* KSPSolve_PGMRES / KSPPGMRESCycle (gmres/pgmres/pgmres.c:17-206, update :226-301, build :176-206): pipelined GMRES with one
   reduction per iteration, posted one iteration before it is used.  The reference's split reductions (VecNormBegin /
   VecMDotBegin) compute their local part when they are BEGUN, so the values below are taken at those points.  Left
   preconditioning, preconditioned norm; restated for the next round's pipelined device KSP. */
int ora_ksp_pgmres(int n, const int *ai, const int *aj, const double *aa, const double *b, double *x, const ora_ksp_opts *o,
                   ora_ksp_result *res, double *hist, int histcap)
{
  const int max_k = o->restart;
  ora_pc    pc;
  g_omp = o->use_omp;
  int rc = pc_setup(&pc, n, ai, aj, aa, o);
  if (rc) return rc;
  const int nvv = max_k + 3;
  double  **vv  = (double **)malloc(sizeof(double *) * (size_t)nvv);
  for (int i = 0; i < nvv; i++) vv[i] = (double *)malloc(sizeof(double) * (size_t)n);
  double *temp = (double *)malloc(sizeof(double) * (size_t)n), *tmat = (double *)malloc(sizeof(double) * (size_t)n);
  double *hh  = (double *)calloc((size_t)(max_k + 2) * (max_k + 2), sizeof(double));
  double *hes = (double *)calloc((size_t)(max_k + 2) * (max_k + 2), sizeof(double));
  double *grs = (double *)calloc((size_t)max_k + 3, sizeof(double)), *cc = (double *)calloc((size_t)max_k + 3, sizeof(double));
  double *ss = (double *)calloc((size_t)max_k + 3, sizeof(double)), *nrs = (double *)calloc((size_t)max_k + 3, sizeof(double));
  double *work = (double *)calloc((size_t)max_k + 3, sizeof(double));
#define PHH(a, b)  (hh + (size_t)(b) * (max_k + 2) + (a))
#define PHES(a, b) (hes + (size_t)(b) * (max_k + 2) + (a))
  conv_ctx cv = {0, 0, o->rtol, o->abstol, o->dtol};
  int      its = 0, reason = 0, nh = 0, guess_zero = 1, itcount = 0;
  double   rnorm = -1.0;
  vzero(n, x);
  while (!reason) {
    /* KSPInitialResidual, left PC */
    if (!guess_zero) {
      k_matmult(&pc, x, temp);
      vcopy(n, b, tmat);
      ora_vecaxpy(n, -1.0, temp, tmat);
      pc_apply(&pc, tmat, vv[0]);
    } else pc_apply(&pc, b, vv[0]);
    /* ---- KSPPGMRESCycle ---- */
    int    it = 0, hapend = 0;
    double resn = ora_vecnorm2(n, vv[0]), pending_norm = 0.0;
    if (resn != 0.0) ora_vecscale(n, 1.0 / resn, vv[0]);
    grs[0] = resn;
    rnorm  = resn;
    LOGRES(rnorm);
    if (!resn) { reason = 3; break; }
    reason = converged_default(&cv, its, rnorm);
    for (; !reason; it++) {
      double *Zcur = vv[it], *Znext = vv[it + 1];
      if (it < max_k + 1 && its + 1 < (o->max_it > 2 ? o->max_it : 2)) { /* Znext <- B A Zcur */
        k_matmult(&pc, Zcur, tmat);
        pc_apply(&pc, tmat, Znext);
      }
      if (it > 1) *PHH(it - 1, it - 2) = pending_norm;  /* VecNormEnd of the norm begun one iteration ago */
      /* (the VecMDotEnd of this point delivers column it-1 of H: written when it was begun, below) */
      if (it > 1) {
        ora_vecscale(n, 1.0 / *PHH(it - 1, it - 2), vv[it - 1]);
        { /* KSPPGMRESUpdateHessenberg(ksp, it - 2, ...) */
          const int c  = it - 2;
          double   *h  = PHH(0, c);
          for (int j = 0; j <= c + 1; j++) *PHES(j, c) = h[j];
          double hapbnd = fabs(h[c + 1] / grs[c]);
          if (hapbnd > 1.0e-30) hapbnd = 1.0e-30;
          if (fabs(h[c + 1]) < hapbnd) hapend = 1;
          for (int j = 0; j < c; j++) {
            const double hhj = h[j];
            h[j]     = cc[j] * hhj + ss[j] * h[j + 1];
            h[j + 1] = -ss[j] * hhj + cc[j] * h[j + 1];
          }
          if (!hapend) {
            const double delta = sqrt(h[c] * h[c] + h[c + 1] * h[c + 1]);
            if (delta == 0.0) { reason = -2; break; }
            cc[c]      = h[c] / delta;
            ss[c]      = h[c + 1] / delta;
            h[c]       = cc[c] * h[c] + ss[c] * h[c + 1];
            grs[c + 1] = -ss[c] * grs[c];
            grs[c]     = cc[c] * grs[c];
            resn       = fabs(grs[c + 1]);
          } else resn = 0.0;
        }
        its++;
        rnorm  = resn;
        reason = converged_default(&cv, its, rnorm);
        if (reason) break;
        if (it < max_k + 1) LOGRES(rnorm);
        if (hapend) { reason = -5; break; }
        if (!(it < max_k + 1 && its < o->max_it)) break;
        {
          const double sc = *PHH(it - 1, it - 2);
          ora_vecscale(n, 1.0 / sc, Zcur);
          ora_vecscale(n, 1.0 / sc, Znext);
          for (int k = 0; k < it; k++) *PHH(k, it - 1) /= sc;
          *PHH(it - 1, it - 1) /= sc;
        }
      }
      if (it > 0) {
        for (int k = 0; k < it + 1; k++) {
          work[k] = 0;
          for (int j = (k - 1 > 0 ? k - 1 : 0); j < it - 1; j++) work[k] -= *PHES(k, j) * *PHH(j, it - 1);
        }
        ora_vecmaxpy(n, it + 1, work, (const double *const *)vv, Znext);
        ora_vecaxpy(n, -*PHH(it - 1, it - 1), Zcur, Znext);
        for (int k = 0; k < it; k++) work[k] = -*PHH(k, it - 1);
        ora_vecmaxpy(n, it, work, (const double *const *)vv, Zcur);
        pending_norm = ora_vecnorm2(n, vv[it]); /* VecNormBegin(VEC_VV(it)) */
      }
      ora_vecmdot(n, it + 1, Znext, (const double *const *)vv, PHH(0, it)); /* VecMDotBegin(Znext, it+1, VV, HH(0,it)) */
    }
    itcount += it - 1 > 0 ? it - 1 : 0;
    /* KSPPGMRESBuildSoln(RS, x, x, ksp, it - 2) */
    {
      const int k = it - 2;
      if (k >= 0) {
        nrs[k] = *PHH(k, k) != 0.0 ? grs[k] / *PHH(k, k) : 0.0;
        for (int kk = k - 1; kk >= 0; kk--) {
          double tt = grs[kk];
          for (int j = kk + 1; j <= k; j++) tt -= *PHH(kk, j) * nrs[j];
          nrs[kk] = tt / *PHH(kk, kk);
        }
        vzero(n, temp);
        ora_vecmaxpy(n, k + 1, nrs, (const double *const *)vv, temp);
        ora_vecaxpy(n, 1.0, temp, x);
      }
    }
    if (!reason && its == o->max_it) reason = -3;
    if (reason) LOGRES(rnorm);
    if (itcount >= o->max_it) { if (!reason) reason = -3; break; }
    guess_zero = 0;
  }
#undef PHH
#undef PHES
  res->its = its; res->reason = reason; res->rnorm = rnorm; res->nhist = nh;
  for (int i = 0; i < nvv; i++) free(vv[i]);
  free(vv); free(temp); free(tmat); free(hh); free(hes); free(grs); free(cc); free(ss); free(nrs); free(work);
  pc_destroy(&pc);
  return 0;
}

/* KSPSolve_PIPECG (cg/pipecg/pipecg.c:19-160), zero initial guess, KSP_NORM_PRECONDITIONED: the three reductions of an
   iteration (|u|, r.u, w.u) are posted together and collected after the PCApply + MatMult they overlap with -- the
   single-reduction CG of SURVEY 8f.3.  Restated for the next round's fused-reduction device KSP. */
int ora_ksp_pipecg(int n, const int *ai, const int *aj, const double *aa, const double *b, double *x, const ora_ksp_opts *o,
                   ora_ksp_result *res, double *hist, int histcap)
{
  ora_pc pc;
  g_omp = o->use_omp;
  int rc = pc_setup(&pc, n, ai, aj, aa, o);
  if (rc) return rc;
  const size_t sz = sizeof(double) * (size_t)n;
  double *R = (double *)malloc(sz), *Z = (double *)malloc(sz), *P = (double *)malloc(sz), *N = (double *)malloc(sz), *W = (double *)malloc(sz);
  double *Q = (double *)malloc(sz), *U = (double *)malloc(sz), *M = (double *)malloc(sz), *S = (double *)malloc(sz);
  conv_ctx cv = {0, 0, o->rtol, o->abstol, o->dtol};
  double   dp, gamma = 0.0, gammaold = 0.0, delta, alpha = 0.0, beta;
  int      i = 0, reason = 0, nh = 0, its = 0;
  memset(x, 0, sz);
  memcpy(R, b, sz);                 /* r <- b (x is 0) */
  pc_apply(&pc, R, U);              /* u <- Br */
  dp = ora_vecnorm2(n, U);
  k_matmult(&pc, U, W);             /* w <- Au */
  LOGRES(dp);
  reason = converged_default(&cv, 0, dp);
  if (!reason) {
    do {
      if (i > 0) dp = ora_vecnorm2(n, U);
      gamma = ora_vecdot(n, R, U);
      delta = ora_vecdot(n, W, U);
      pc_apply(&pc, W, M);          /* m <- Bw */
      k_matmult(&pc, M, N);         /* n <- Am */
      if (i > 0) {
        LOGRES(dp);
        reason = converged_default(&cv, i, dp);
        if (reason) break;
      }
      if (i == 0) {
        alpha = gamma / delta;
        memcpy(Z, N, sz); memcpy(Q, M, sz); memcpy(P, U, sz); memcpy(S, W, sz);
      } else {
        beta  = gamma / gammaold;
        alpha = gamma / (delta - beta / alpha * gamma);
        ora_vecaypx(n, beta, N, Z); /* z <- n + beta z */
        ora_vecaypx(n, beta, M, Q);
        ora_vecaypx(n, beta, U, P);
        ora_vecaypx(n, beta, W, S);
      }
      ora_vecaxpy(n, alpha, P, x);
      ora_vecaxpy(n, -alpha, Q, U);
      ora_vecaxpy(n, -alpha, Z, W);
      ora_vecaxpy(n, -alpha, S, R);
      gammaold = gamma;
      i++;
      its = i;
    } while (i <= o->max_it);
    if (!reason) reason = -3;
  }
  res->its = its; res->reason = reason; res->rnorm = dp; res->nhist = nh;
  free(R); free(Z); free(P); free(N); free(W); free(Q); free(U); free(M); free(S);
  pc_destroy(&pc);
  return 0;
}

/* ================================================================================================================== */
/* widening rows: transposed product and COO assembly                                                                 */
/* ================================================================================================================== */
void ora_matmulttranspose_seqaij(int m, int n, const int *ai, const int *aj, const double *aa, const double *x, const double *z, double *y)
{
  for (int c = 0; c < n; c++) y[c] = z ? z[c] : 0.0;
  for (int i = 0; i < m; i++) {
    const double alpha = x[i];
    for (int k = ai[i]; k < ai[i + 1]; k++) {
      const double p = alpha * aa[k];
      y[aj[k]] = y[aj[k]] + p;
    }
  }
}

/* the reference's quicksort on a key array with one or two companion arrays (sorti.c:198-240); the permutation it produces
   for equal keys is part of MatSetValuesCOO's summation order, so the recursion, the pivot choice and the partition are
   followed exactly: Y may be NULL (two-array variant) */
static int64_t coo_median_pos(const int *X, int64_t hi)
{
  const int64_t a = hi / 4, b = hi / 2, c = hi / 4 * 3;
  if (X[a] < X[b]) {
    if (X[b] < X[c]) return b;
    return X[a] < X[c] ? c : a;
  }
  if (X[c] < X[b]) return b;
  return X[a] < X[c] ? a : c;
}
static void coo_swap(int *X, int *Y, int64_t *Z, int64_t p, int64_t q)
{
  int     t = X[p];
  int64_t u = Z[p];
  X[p] = X[q]; X[q] = t;
  Z[p] = Z[q]; Z[q] = u;
  if (Y) { t = Y[p]; Y[p] = Y[q]; Y[q] = t; }
}
static void coo_quicksort(int64_t n, int *X, int *Y, int64_t *Z)
{
  const int64_t hi = n - 1;
  if (n < 8) {
    for (int64_t i = 0; i < n; i++) {
      int pivot = X[i];
      for (int64_t j = i + 1; j < n; j++) {
        if (pivot > X[j]) {
          coo_swap(X, Y, Z, i, j);
          pivot = X[i];
        }
      }
    }
    return;
  }
  const int pivot = X[coo_median_pos(X, hi)];
  int64_t   l = 0, r = hi;
  for (;;) {
    while (X[l] < pivot) l++;
    while (X[r] > pivot) r--;
    if (l >= r) { r++; break; }
    coo_swap(X, Y, Z, l, r);
    l++; r--;
  }
  coo_quicksort(l, X, Y, Z);
  coo_quicksort(hi - r + 1, X + r, Y ? Y + r : NULL, Z + r);
}

int ora_coo_prealloc(int M, int N, int64_t coo_n, const int *coo_i, const int *coo_j, int *Ai, int *Aj, int64_t *jmap, int64_t *perm,
                     int64_t *nnz_out, int64_t *atot_out)
{
  int     *i = (int *)malloc(sizeof(int) * (size_t)(coo_n + 1)), *j = (int *)malloc(sizeof(int) * (size_t)(coo_n + 1));
  int64_t *pm = (int64_t *)malloc(sizeof(int64_t) * (size_t)(coo_n + 1));
  int64_t  k, q = 0, nnz = 0, nneg;
  int      sorted = 1, prev = INT32_MIN, rc = 0;
  for (k = 0; k < coo_n; k++) {
    i[k] = coo_j[k] < 0 ? -1 : coo_i[k];
    j[k] = coo_j[k];
    if (sorted) {
      if (i[k] < prev) sorted = 0;
      else prev = i[k];
    }
    pm[k] = k;
  }
  if (!sorted) coo_quicksort(coo_n, i, j, pm);
  if (coo_n && i[coo_n - 1] >= M) { rc = 1; goto done; }
  for (k = 0; k < coo_n; k++)
    if (i[k] >= 0) break;
  nneg = k;
  for (int r = 0; r <= M; r++) Ai[r] = 0;
  while (k < coo_n) {
    const int     row = i[k];
    const int64_t start = k;
    int           jprev = INT32_MIN, strict = 1;
    while (k < coo_n && i[k] == row) {
      if (strict) {
        if (j[k] <= jprev) strict = 0;
        else jprev = j[k];
      }
      k++;
    }
    if (!strict) coo_quicksort(k - start, j + start, NULL, pm + start);
    if (k > start && j[k - 1] >= N) { rc = 2; goto done; }
    /* unique columns of this row; jmap[q+1] first holds the repeat count of the q-th nonzero */
    for (int64_t p = start; p < k; p++) {
      if (p == start || j[p] != j[p - 1]) {
        Aj[q]       = j[p];
        jmap[q + 1] = 1;
        Ai[row + 1]++;
        q++;
        nnz++;
      } else jmap[q]++;
    }
  }
  for (int r = 0; r < M; r++) Ai[r + 1] += Ai[r];
  jmap[0] = 0;
  for (k = 0; k < nnz; k++) jmap[k + 1] += jmap[k];
  for (k = 0; k < coo_n - nneg; k++) perm[k] = pm[k + nneg];
  *nnz_out  = nnz;
  *atot_out = coo_n - nneg;
done:
  free(i); free(j); free(pm);
  return rc;
}

void ora_coo_setvalues(int64_t nnz, const int64_t *jmap, const int64_t *perm, const double *v, int insert, double *Aa)
{
  for (int64_t q = 0; q < nnz; q++) {
    double sum = 0.0;
    for (int64_t k = jmap[q]; k < jmap[q + 1]; k++) sum += v[perm[k]];
    Aa[q] = (insert ? 0.0 : Aa[q]) + sum;
  }
}

/* ================================================================================================================== */
/* ICC(0), natural ordering                                                                                           */
/* ================================================================================================================== */
int ora_icc0_symbolic(int n, const int *ai, const int *aj, int *ui, int *uj, int *udiag)
{
  ui[0] = 0;
  for (int i = 0; i < n; i++) {
    int d = -1;
    for (int k = ai[i]; k < ai[i + 1]; k++)
      if (aj[k] == i) { d = k; break; }
    if (d < 0) return -(i + 1);
    int q = ui[i];
    for (int k = d + 1; k < ai[i + 1]; k++) uj[q++] = aj[k]; /* strictly upper part, column order */
    uj[q]     = i;                                           /* the diagonal is the last entry of the row */
    udiag[i]  = q;
    ui[i + 1] = q + 1;
  }
  return 0;
}

int ora_icc0_numeric(int n, const int *ai, const int *aj, const double *aa, const int *ui, const int *uj, const int *udiag, double *ua, double zeropivot)
{
  double *rtmp = (double *)calloc((size_t)n + 1, sizeof(double));
  int    *il = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *c2r = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  int     rc = 0;
  /* c2r[col]: head of the list of finished rows whose first not-yet-used entry lies in column col (and, for a row that is in
     such a list, the next row of that list); il[i]: position of that entry in row i */
  for (int i = 0; i < n; i++) c2r[i] = n;
  if (n) il[0] = 0;
  for (int k = 0; k < n; k++) {
    for (int j = ui[k]; j < ui[k + 1]; j++) rtmp[uj[j]] = 0.0;
    {
      double *bval = ua + ui[k];
      for (int j = ai[k]; j < ai[k + 1]; j++)
        if (aj[j] >= k) { /* upper triangle of A only */
          rtmp[aj[j]] = aa[j];
          *bval++     = 0.0;
        }
    }
    double dk = rtmp[k];
    int    i  = c2r[k];
    while (i < k) {
      const int    nexti = c2r[i];
      const int    ili   = il[i];
      const double uikdi = -ua[ili] * ua[udiag[i]];
      dk += uikdi * ua[ili];
      ua[ili] = uikdi;
      const int jmin = ili + 1, jmax = ui[i + 1]; /* through the diagonal slot, exactly as the reference's loop runs */
      if (jmin < jmax) {
        for (int j = jmin; j < jmax; j++) rtmp[uj[j]] += uikdi * ua[j];
        il[i] = jmin;
        const int j = uj[jmin];
        c2r[i]      = c2r[j];
        c2r[j]      = i;
      }
      i = nexti;
    }
    double    rs   = 0.0;
    const int jmin = ui[k], jmax = ui[k + 1] - 1;
    if (jmin < jmax) {
      for (int j = jmin; j < jmax; j++) {
        ua[j] = rtmp[uj[j]];
        rs += fabs(ua[j]);
      }
      il[k]        = jmin;
      const int c  = uj[jmin];
      c2r[k]       = c2r[c];
      c2r[c]       = k;
    }
    if (dk <= zeropivot * rs) { /* MatPivotCheck_pd would start shifting here */
      rc = -(k + 1);
      break;
    }
    ua[udiag[k]] = 1.0 / dk;
  }
  free(rtmp); free(il); free(c2r);
  return rc;
}

void ora_matsolve_icc(int n, const int *ui, const int *uj, const int *udiag, const double *ua, const double *b, double *x)
{
  (void)udiag;
  memcpy(x, b, sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) { /* U^T D y = b */
    const double xi = x[i];
    const int    nz = ui[i + 1] - ui[i] - 1;
    const int   *vj = uj + ui[i];
    const double *v = ua + ui[i];
    for (int j = 0; j < nz; j++) x[vj[j]] += v[j] * xi;
    x[i] = xi * v[nz];
  }
  for (int i = n - 2; i >= 0; i--) { /* U x = y, entries from the end of the row backwards */
    double     xi = x[i];
    const int  nz = ui[i + 1] - ui[i] - 1;
    const int  e  = ui[i + 1] - 2; /* last off-diagonal entry */
    for (int j = 0; j < nz; j++) xi += ua[e - j] * x[uj[e - j]];
    x[i] = xi;
  }
}

/* ---- PetscSF local scatter: PetscSFLinkScatterLocal (sfpack.c:1082) with the ScatterAnd<Op> loop of sfpack.c:211-218 ----
   for i in order: dst[didx[i]*bs + c] = dst[...] op src[sidx[i]*bs + c]; NULL index = contiguous from 0.
   op: 0 REPLACE (OP_ASSIGN), 1 SUM, 2 PROD (OP_BINARY), 3 MAX, 4 MIN (OP_FUNCTION with PetscMax / PetscMin, petscmath.h) */
void ora_sf_scatter_f64(int64_t n, int bs, int op, const int *sidx, const int *didx, const double *src, double *dst)
{
  for (int64_t i = 0; i < n; i++) {
    const int64_t s = (int64_t)(sidx ? sidx[i] : i) * bs, t = (int64_t)(didx ? didx[i] : i) * bs;
    for (int c = 0; c < bs; c++) {
      const double u = src[s + c];
      double      *v = &dst[t + c];
      switch (op) {
      case 0: *v = u; break;
      case 1: *v = *v + u; break;
      case 2: *v = *v * u; break;
      case 3: *v = (*v < u) ? u : *v; break;
      case 4: *v = (*v < u) ? *v : u; break;
      }
    }
  }
}
void ora_sf_scatter_i32(int64_t n, int bs, int op, const int *sidx, const int *didx, const int *src, int *dst)
{
  for (int64_t i = 0; i < n; i++) {
    const int64_t s = (int64_t)(sidx ? sidx[i] : i) * bs, t = (int64_t)(didx ? didx[i] : i) * bs;
    for (int c = 0; c < bs; c++) {
      const int u = src[s + c];
      int      *v = &dst[t + c];
      switch (op) {
      case 0: *v = u; break;
      case 1: *v = (int)((unsigned)*v + (unsigned)u); break;
      case 2: *v = (int)((unsigned)*v * (unsigned)u); break;
      case 3: *v = (*v < u) ? u : *v; break;
      case 4: *v = (*v < u) ? *v : u; break;
      }
    }
  }
}
