/*
 * icc.cu -- ICC(0) for MATSEQAIJ on sm_100a: symbolic layout + merge schedule, numeric factorisation, and the two
 * triangular sweeps of the solve.
 *
 * Reference: MatICCFactorSymbolic_SeqAIJ (levels 0, natural ordering; aijfact.c:2049-2094), MatCholeskyFactorNumeric_SeqAIJ
 * (aijfact.c:1701-1866), MatSolve_SeqSBAIJ_1_NaturalOrdering (sbaij/seq/sbaijfact2.c:2030-2065).  PCICC is PETSc's default
 * preconditioner for a sequential matrix flagged symmetric (ex2.c does that).
 *
 * Factor layout = the reference's: row i of U holds the strictly upper entries of A's row i in column order and the diagonal
 * LAST; after the numeric phase the off-diagonal slot (i,c) holds -U(i,c)/D(i) and the diagonal slot 1/D(i).
 *
 * Numeric phase.  The reference merges, into row k, the finished rows i < k with U(i,k) != 0 in the order of its linked lists
 * (c2r / il); that LIFO order fixes the floating-point association of every entry of row k and depends on the pattern alone.
 * The host walks the same lists on indices once (merge schedule: per row k its contributors (i, position of U(i,k))), and the
 * device kernel recomputes each row from A's row and the ORIGINAL (unscaled) entries of its contributors in exactly that
 * order -- rows of one dependency level are independent, a row waits for its contributors through per-row flags (release /
 * acquire), no grid barrier.  `orig` keeps the unscaled entries (what later rows read), `final` receives what the solves
 * read; each off-diagonal slot has exactly one consumer row (its column's), so nothing races.  FMA-free (__dmul_rn /
 * __dadd_rn): the factor is bit-identical to the reference's.
 *
 * Solve.  The reference's forward sweep scatters (x[c] += v(i,c) * x_i for rows i ascending); a parallel sweep performs it
 * as a GATHER along column c of U (explicit column view tptr/trow + values re-gathered after every factorisation) in the same
 * ascending-i order; the backward sweep gathers along row i from its last off-diagonal entry to its first.  Both run on
 * the segment-marching sweep kernel of ilu.cu (modes 2 and 3), strict order, FMA-free: bit-identical to the reference.
 */
#include "b200_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct b200IccPlan_s {
  int     n;
  int64_t nnzA, nzu;
  int    *d_ai, *d_adiag;          /* A: row start, diagonal position */
  int    *d_ui, *d_uj, *d_udiag;   /* factor layout */
  double *d_orig, *d_final, *d_dinv;
  int    *d_mptr, *d_mrow, *d_mpos; /* merge schedule */
  int    *d_order, nslot, nlev;    /* rows in dependency-level order, padded */
  int    *d_flag, *d_status, epoch;
  int    *d_tptr, *d_trow, *d_tpos; /* column view of the strictly upper part */
  double *d_tval, *d_y, *d_xf;
  int2   *d_segF, *d_segB;
  int     nslotF, nslotB, nlevF, nlevB, GS;
  int    *h_ui, *h_uj, *h_udiag;
  int     factored;
};

#define ICC_TPB 256
#define ICC_G 8 /* lanes per row in the numeric kernel */

__device__ __forceinline__ int icc_ld_acquire(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void icc_st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

/* position of column c inside the sorted segment uj[lo,hi), or -1 */
__device__ __forceinline__ int icc_find(const int *__restrict__ uj, int lo, int hi, int c)
{
  while (lo < hi) {
    const int mid = (lo + hi) >> 1, v = uj[mid];
    if (v == c) return mid;
    if (v < c) lo = mid + 1;
    else hi = mid;
  }
  return -1;
}

/* one group of ICC_G lanes per row k, rows in level order (round-robin over a co-resident grid: a row only waits for rows that
   precede it in that order) */
__global__ void __launch_bounds__(ICC_TPB) icc_numeric_kernel(int nslot, const int *__restrict__ order, const int *__restrict__ ai, const int *__restrict__ adiag, const double *__restrict__ aval,
                                                              const int *__restrict__ ui, const int *__restrict__ uj, const int *__restrict__ udiag, const int *__restrict__ mptr, const int *__restrict__ mrow,
                                                              const int *__restrict__ mpos, double *orig, double *fin, double *dinv, double zeropivot, int *flag, int epoch, int *status)
{
  constexpr int  G = ICC_G, RPW = 32 / G;
  const int      gl = threadIdx.x % G, grp = (threadIdx.x & 31) / G;
  const unsigned gmask = ((1u << G) - 1u) << (grp * G);
  const int      wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), W = (int)((gridDim.x * blockDim.x) >> 5);
  for (int64_t chunk = wid; chunk * RPW < nslot; chunk += W) {
    const int64_t slot = chunk * RPW + grp;
    const int     k    = slot < nslot ? order[slot] : -1;
    if (k < 0) continue; /* padding: whole groups only, the other groups of the warp keep going (no warp-wide barrier below) */
    const int u0 = ui[k], ud = udiag[k]; /* off-diagonal slots [u0, ud), diagonal slot ud */
    /* rtmp = upper triangle of A's row (aijfact.c:1745-1752): the factor row has exactly A's upper pattern (levels 0) */
    const int ad = adiag[k];
    for (int t = gl; t < ud - u0; t += G) orig[u0 + t] = aval[ad + 1 + t];
    double dk = aval[ad];
    __syncwarp(gmask);
    for (int q = mptr[k]; q < mptr[k + 1]; q++) { /* contributors in the reference's list order */
      const int i = mrow[q], e = mpos[q];
      {
        unsigned ns = 0;
        while (icc_ld_acquire(flag + i) != epoch) {
          if (ns) __nanosleep(ns);
          ns = ns ? (ns < 256 ? ns * 2 : 256) : 32;
        }
      }
      const double u     = __ldcg(orig + e);
      const double uikdi = __dmul_rn(-u, __ldcg(dinv + i)); /* -ba[ili] * ba[bdiag[i]] */
      dk                 = __dadd_rn(dk, __dmul_rn(uikdi, u));
      if (gl == 0) fin[e] = uikdi;
      const int iend = udiag[i];
      for (int t = e + 1 + gl; t < iend; t += G) { /* later off-diagonal entries of row i (the reference's loop also touches the
                                                      diagonal slot and columns outside row k's pattern: dead stores) */
        const int pos = icc_find(uj, u0, ud, uj[t]);
        if (pos >= 0) orig[pos] = __dadd_rn(__ldcg(orig + pos), __dmul_rn(uikdi, __ldcg(orig + t)));
      }
      __syncwarp(gmask); /* the next contributor may update the same slots from other lanes */
    }
    double rs = 0.0;
    for (int t = u0 + gl; t < ud; t += G) rs += fabs(__ldcg(orig + t));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) rs += __shfl_xor_sync(gmask, rs, o, G);
    if (gl == 0) {
      if (!(dk > zeropivot * rs)) atomicMax(status, k + 1); /* MatPivotCheck_pd would start shifting here (matimpl.h:813-833) */
      const double di = 1.0 / dk;
      dinv[k] = di;
      fin[ud] = di;
      __threadfence();
      icc_st_release(flag + k, epoch);
    }
  }
}

__global__ void icc_gather_kernel(int64_t n, const int *__restrict__ tpos, const double *__restrict__ fin, double *__restrict__ tval)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) tval[q] = fin[tpos[q]];
}

extern "C" int b200Icc0Destroy(b200IccPlan p)
{
  if (!p) return 0;
  cudaFree(p->d_ai); cudaFree(p->d_adiag); cudaFree(p->d_ui); cudaFree(p->d_uj); cudaFree(p->d_udiag); cudaFree(p->d_orig); cudaFree(p->d_final); cudaFree(p->d_dinv);
  cudaFree(p->d_mptr); cudaFree(p->d_mrow); cudaFree(p->d_mpos); cudaFree(p->d_order); cudaFree(p->d_flag); cudaFree(p->d_status);
  cudaFree(p->d_tptr); cudaFree(p->d_trow); cudaFree(p->d_tpos); cudaFree(p->d_tval); cudaFree(p->d_y); cudaFree(p->d_xf); cudaFree(p->d_segF); cudaFree(p->d_segB);
  free(p->h_ui); free(p->h_uj); free(p->h_udiag);
  free(p);
  return 0;
}

extern "C" int b200Icc0Symbolic(b200Handle h, int n, const int *ai, const int *aj, b200IccPlan *plan)
{
  B200_CHECK(h && plan, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 0 && (n == 0 || (ai && aj)), B200_ERR_ARG_NULL, "null pattern");
  b200IccPlan p = (b200IccPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->n    = n;
  p->nnzA = n ? ai[n] : 0;
  int *adiag = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *ui = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *udiag = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  B200_CHECK(adiag && ui && udiag, B200_ERR_MEM, "out of host memory");
  /* MatICCFactorSymbolic_SeqAIJ, levels 0 (aijfact.c:2078-2094) */
  ui[0] = 0;
  for (int i = 0; i < n; i++) {
    int lo = ai[i], hi = ai[i + 1], pos = -1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (aj[mid] == i) { pos = mid; break; }
      if (aj[mid] < i) lo = mid + 1;
      else hi = mid;
    }
    if (pos < 0) {
      free(adiag); free(ui); free(udiag); free(p);
      B200_CHECK(0, B200_ERR_ARG_WRONGSTATE, "Matrix is missing diagonal entry %d", i); /* aijfact.c:2071 */
    }
    adiag[i]  = pos;
    ui[i + 1] = ui[i] + (ai[i + 1] - pos);
  }
  const int64_t nzu = n ? ui[n] : 0;
  p->nzu            = nzu;
  int *uj = (int *)malloc(sizeof(int) * ((size_t)nzu + 1));
  B200_CHECK(uj, B200_ERR_MEM, "out of host memory");
  for (int i = 0; i < n; i++) {
    const int len = ai[i + 1] - adiag[i] - 1;
    memcpy(uj + ui[i], aj + adiag[i] + 1, sizeof(int) * (size_t)len);
    uj[ui[i] + len] = i;
    udiag[i]        = ui[i] + len;
  }
  /* merge schedule (the lists of aijfact.c:1750-1800 walked on indices) + dependency levels */
  const int64_t noff = nzu - n;
  int *mptr = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *mrow = (int *)malloc(sizeof(int) * ((size_t)noff + 1)), *mpos = (int *)malloc(sizeof(int) * ((size_t)noff + 1));
  int *level = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *c2r = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *il = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  B200_CHECK(mptr && mrow && mpos && level && c2r && il, B200_ERR_MEM, "out of host memory");
  int q = 0, nlev = 0;
  for (int i = 0; i <= n; i++) c2r[i] = n;
  mptr[0] = 0;
  for (int k = 0; k < n; k++) {
    int i = c2r[k], lev = 0;
    while (i < k) {
      const int nexti = c2r[i], ili = il[i], jmin = ili + 1, jmax = ui[i + 1];
      mrow[q] = i;
      mpos[q] = ili;
      q++;
      if (level[i] + 1 > lev) lev = level[i] + 1;
      if (jmin < jmax) { /* advance row i to its next entry and push it on that column's list */
        const int j = uj[jmin];
        il[i]  = jmin;
        c2r[i] = c2r[j];
        c2r[j] = i;
      }
      i = nexti;
    }
    if (ui[k] < ui[k + 1] - 1) { /* row k has off-diagonal entries: it joins the list of its first one */
      const int c = uj[ui[k]];
      il[k]  = ui[k];
      c2r[k] = c2r[c];
      c2r[c] = k;
    }
    level[k]    = lev;
    mptr[k + 1] = q;
    if (lev + 1 > nlev) nlev = lev + 1;
  }
  free(c2r); free(il);
  p->nlev = nlev;
  /* rows in level order, each level padded to a multiple of the rows per warp */
  const int rpw = 32 / ICC_G;
  int64_t  *cnt = (int64_t *)calloc((size_t)nlev + 2, sizeof(int64_t)), tot = 0;
  for (int k = 0; k < n; k++) cnt[level[k] + 1]++;
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nlev + 1));
  for (int l = 0; l < nlev; l++) {
    start[l] = tot;
    tot += (cnt[l + 1] + rpw - 1) / rpw * rpw;
  }
  int *order = (int *)malloc(sizeof(int) * ((size_t)tot + 1));
  for (int64_t k = 0; k < tot; k++) order[k] = -1;
  for (int k = 0; k < n; k++) order[start[level[k]]++] = k;
  p->nslot = (int)tot;
  free(cnt); free(start); free(level);
  /* column view of the strictly upper part: for column c the entries (i,c) in ascending i (sbaijfact2.c:2045-2052 as a gather) */
  int *tptr = (int *)calloc((size_t)n + 2, sizeof(int)), *trow = (int *)malloc(sizeof(int) * ((size_t)noff + 1)), *tpos = (int *)malloc(sizeof(int) * ((size_t)noff + 1));
  B200_CHECK(tptr && trow && tpos, B200_ERR_MEM, "out of host memory");
  for (int i = 0; i < n; i++)
    for (int t = ui[i]; t < udiag[i]; t++) tptr[uj[t] + 1]++;
  for (int c = 0; c < n; c++) tptr[c + 1] += tptr[c];
  {
    int *fill = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    memcpy(fill, tptr, sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++)
      for (int t = ui[i]; t < udiag[i]; t++) {
        const int c   = uj[t];
        trow[fill[c]] = i;
        tpos[fill[c]] = t;
        fill[c]++;
      }
    free(fill);
  }
  /* segment schedules of the two marching sweeps */
  {
    double avg = n ? (double)noff / n : 0.0;
    int    GS  = 2;
    while (GS < 32 && GS < avg + 0.5) GS <<= 1;
    p->GS = GS;
    const char *e1 = getenv("PETSCB200_ILU_SEG_MIN"), *e2 = getenv("PETSCB200_ILU_SEG_MAX");
    const int   minlen = e1 && atoi(e1) > 0 ? atoi(e1) : 8, maxlen = e2 && atoi(e2) > 0 ? atoi(e2) : 1024;
    int2 *sF = b200_build_segments(n, 2, tptr, trow, 32 / GS, minlen, maxlen, &p->nslotF, &p->nlevF);
    int2 *sB = b200_build_segments(n, 3, ui, uj, 32 / GS, minlen, maxlen, &p->nslotB, &p->nlevB);
    B200_CUDA(cudaMalloc(&p->d_segF, sizeof(int2) * ((size_t)p->nslotF + 64)));
    B200_CUDA(cudaMalloc(&p->d_segB, sizeof(int2) * ((size_t)p->nslotB + 64)));
    B200_CUDA(cudaMemcpyAsync(p->d_segF, sF, sizeof(int2) * (size_t)p->nslotF, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaMemcpyAsync(p->d_segB, sB, sizeof(int2) * (size_t)p->nslotB, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    free(sF); free(sB);
  }
#define UP(dst, src, cnt_, T) \
  do { \
    B200_CUDA(cudaMalloc(&p->dst, sizeof(T) * ((size_t)(cnt_) + 64))); \
    B200_CUDA(cudaMemcpyAsync(p->dst, src, sizeof(T) * (size_t)(cnt_), cudaMemcpyHostToDevice, h->stream)); \
  } while (0)
  UP(d_ai, ai, n + 1, int);
  UP(d_adiag, adiag, n, int);
  UP(d_ui, ui, n + 1, int);
  UP(d_uj, uj, nzu, int);
  UP(d_udiag, udiag, n, int);
  UP(d_mptr, mptr, n + 1, int);
  UP(d_mrow, mrow, noff, int);
  UP(d_mpos, mpos, noff, int);
  UP(d_order, order, p->nslot, int);
  UP(d_tptr, tptr, n + 1, int);
  UP(d_trow, trow, noff, int);
  UP(d_tpos, tpos, noff, int);
#undef UP
  B200_CUDA(cudaMalloc(&p->d_orig, sizeof(double) * ((size_t)nzu + 64)));
  B200_CUDA(cudaMalloc(&p->d_final, sizeof(double) * ((size_t)nzu + 64)));
  B200_CUDA(cudaMalloc(&p->d_tval, sizeof(double) * ((size_t)noff + 64)));
  B200_CUDA(cudaMalloc(&p->d_dinv, sizeof(double) * ((size_t)n + 64)));
  B200_CUDA(cudaMalloc(&p->d_y, sizeof(double) * ((size_t)n + 64)));
  B200_CUDA(cudaMalloc(&p->d_xf, sizeof(double) * ((size_t)n + 64)));
  B200_CUDA(cudaMalloc(&p->d_flag, sizeof(int) * ((size_t)n + 64)));
  B200_CUDA(cudaMemsetAsync(p->d_flag, 0, sizeof(int) * ((size_t)n + 64), h->stream));
  B200_CUDA(cudaMalloc(&p->d_status, 64));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  free(adiag); free(mptr); free(mrow); free(mpos); free(order); free(tptr); free(trow); free(tpos);
  p->h_ui = ui; p->h_uj = uj; p->h_udiag = udiag;
  *plan = p;
  return 0;
}

/* numeric factorisation from A's device values.  *zero_pivot_row = 0 on success, else 1 + the largest row whose pivot failed
   dk > zeropivot * rowsum (the reference's MatPivotCheck_pd would shift and refactor; this path reports it instead) */
extern "C" int b200Icc0Numeric(b200Handle h, b200IccPlan p, const double *d_aval, double zeropivot, int *zero_pivot_row)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (zero_pivot_row) *zero_pivot_row = 0;
  if (p->n == 0) {
    p->factored = 1;
    return 0;
  }
  B200_CHECK(d_aval, B200_ERR_ARG_NULL, "null values");
  int coop = 0;
  B200_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
  B200_CHECK(coop == 1, B200_ERR_SUP, "the ICC(0) factorisation needs cooperative kernel launch (co-resident grid)");
  static int occ = 0;
  if (!occ) {
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icc_numeric_kernel, ICC_TPB, 0));
    if (occ < 1) occ = 1;
  }
  p->epoch++;
  B200_CUDA(cudaMemsetAsync(p->d_status, 0, 64, h->stream));
  const int rpc = (ICC_TPB / 32) * (32 / ICC_G);
  int       g   = (p->nslot + rpc - 1) / rpc;
  if (g > occ * h->num_sms) g = occ * h->num_sms;
  if (g < 1) g = 1;
  void *args[] = {&p->nslot, &p->d_order, &p->d_ai, &p->d_adiag, &d_aval, &p->d_ui, &p->d_uj, &p->d_udiag, &p->d_mptr, &p->d_mrow, &p->d_mpos, &p->d_orig, &p->d_final, &p->d_dinv, &zeropivot, &p->d_flag, &p->epoch, &p->d_status};
  B200_CUDA(cudaLaunchCooperativeKernel((void *)icc_numeric_kernel, dim3(g), dim3(ICC_TPB), args, 0, h->stream));
  B200_LAUNCHED(1);
  const int64_t noff = p->nzu - p->n;
  if (noff) {
    int gg = (int)((noff + 255) / 256);
    if (gg > h->num_sms * 8) gg = h->num_sms * 8;
    icc_gather_kernel<<<gg, 256, 0, h->stream>>>(noff, p->d_tpos, p->d_final, p->d_tval);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
  }
  int status = 0;
  B200_CUDA(cudaMemcpyAsync(&status, p->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (zero_pivot_row) *zero_pivot_row = status;
  p->factored = status ? 0 : 1;
  return 0;
}

/* x = U^-1 D^-1 U^-T b exactly as MatSolve_SeqSBAIJ_1_NaturalOrdering orders the operations */
extern "C" int b200Icc0Solve(b200Handle h, b200IccPlan p, const double *d_b, double *d_x)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(p->factored, B200_ERR_ORDER, "b200Icc0Numeric must succeed first");
  if (p->n == 0) return 0;
  B200_CHECK(d_b && d_x && d_b != d_x, B200_ERR_ARG_WRONG, "b and x must be distinct non-null vectors");
  B200_CUDA(cudaMemsetAsync(p->d_y, 0xFF, sizeof(double) * (size_t)p->n, h->stream)); /* sentinel fill (value-as-flag) */
  B200_CUDA(cudaMemsetAsync(d_x, 0xFF, sizeof(double) * (size_t)p->n, h->stream));
  const int64_t noff = p->nzu - p->n;
  int rc = b200_sweep_march(h, p->GS, 2, p->nslotF, p->d_segF, p->d_tptr, p->d_trow, p->d_tval, d_b, p->d_y, noff > 0 ? noff : 1, p->d_dinv, p->d_xf);
  if (rc) return rc;
  return b200_sweep_march(h, p->GS, 3, p->nslotB, p->d_segB, p->d_ui, p->d_uj, p->d_final, p->d_xf, d_x, p->nzu, NULL, NULL);
}

extern "C" int b200Icc0GetFactor(b200Handle h, b200IccPlan p, int *ui, int *uj, int *udiag, double *ua)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (ui) memcpy(ui, p->h_ui, sizeof(int) * ((size_t)p->n + 1));
  if (udiag) memcpy(udiag, p->h_udiag, sizeof(int) * (size_t)p->n);
  if (uj) memcpy(uj, p->h_uj, sizeof(int) * (size_t)p->nzu);
  if (ua && p->nzu) {
    B200_CUDA(cudaMemcpyAsync(ua, p->d_final, sizeof(double) * (size_t)p->nzu, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

extern "C" int b200Icc0GetInfo(b200IccPlan p, int64_t *nz_factor, int *nlevels_numeric, int *nlev_forward, int *nlev_backward)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  if (nz_factor) *nz_factor = p->nzu;
  if (nlevels_numeric) *nlevels_numeric = p->nlev;
  if (nlev_forward) *nlev_forward = p->nlevF;
  if (nlev_backward) *nlev_backward = p->nlevB;
  return 0;
}
