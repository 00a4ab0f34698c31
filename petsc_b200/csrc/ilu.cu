/*
 * ilu.cu -- ILU(0) for MATSEQAIJ on sm_100a: symbolic layout, numeric factorisation and the two triangular sweeps.
 *
 * Reference: MatILUFactorSymbolic_SeqAIJ_ilu0 (aijfact.c:1471-1534), MatLUFactorNumeric_SeqAIJ (aijfact.c:216-389),
 * MatSolve_SeqAIJ_NaturalOrdering (aijfact.c:2413-2457); the reference's GPU path is cusparseXcsrilu02 + cusparseSpSV
 * (aijcusparse.cu:766-827, 643-693).
 *
 * Factor layout = the reference's (aijfact.c:1454-1469): bj/ba hold L(0,:)..L(n-1,:) then U(n-1,:)..U(0,:); bi[i] = start
 * of L(i,:); bdiag[i] = position of U's diagonal (stored INVERTED), U(i,:) = (strict upper ..., diagonal).
 *
 * Parallel schedule: dependency levels are computed once (level(i) = 1 + max level over the row's L entries; mirrored
 * for U), rows are laid out in level order with every level padded to a whole warp, and ONE kernel per sweep walks that
 * order.  CTAs take their slice through an atomic ticket (so they start in dependency order) and a row whose inputs
 * are not final yet spins on per-row ready flags written with release semantics by the producer (no grid-wide barrier,
 * no kernel launch per level: the critical path costs one L2 round trip per level instead of one launch).
 *
 * Arithmetic: G lanes cooperate on a row (they fetch the factor entries and the gathered x values in parallel), but the
 * row sum is accumulated strictly left to right with __dmul_rn/__dsub_rn, exactly PetscSparseDenseMinusDot
 * (aij.h:531-536) in the reference's FMA-free -O2 build.  The numeric factorisation applies the eliminations of a row
 * in the reference's order (L entries ascending; each target entry updated once per pivot row).  Both the factor and
 * the solve are therefore bit-identical to the CPU reference.
 */
#include "b200_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct b200IluPlan_s {
  int     n;
  int64_t nnz;
  int    *d_ai, *d_adiag;            /* A pattern: row start and diagonal position */
  int    *d_bi, *d_bj, *d_bdiag;     /* factor layout */
  double *d_ba;
  int    *d_orderL, *d_orderU;       /* rows in level order, -1 padded to warp multiples */
  int4   *d_metaL, *d_metaU;         /* per slot: (row, first entry, end entry, 0) of the sweep's row segment */
  int     nslotL, nslotU;
  /* slot-space ("level-set reordered") copy for the packed sweeps: everything a sweep touches is laid out in the order it is
     processed, so every access of a warp is one coalesced run instead of 32/G scattered sectors */
  int     packed;                    /* 1: every triangular row fits G entries -> packed sweeps are used */
  int    *d_slotL, *d_slotU;         /* row -> slot */
  int    *d_pkcolL, *d_pkcolU;       /* [nslot*G] dependency SLOT of entry e of the slot's row, -1 = padding */
  double *d_pkvalL, *d_pkvalU, *d_pkdinvU, *d_tL, *d_xU;
  int    *d_mapLU;                   /* U slot -> L slot of the same row (right-hand side of the upper sweep = result of the lower one) */
  int2   *d_segL, *d_segU;           /* segment schedule of the marching sweeps: (first row, number of rows) per slot, level order */
  int     nsegslotL, nsegslotU, nseglevL, nseglevU, GS;
  int    *d_flag;                    /* per-row ready epoch (numeric factorisation) */
  double *d_tmp;                     /* result of the lower sweep */
  int    *d_ticket;                  /* [4] tickets + status */
  int     epoch;
  int     nlevL, nlevU;
  int     maxwL, maxwU;              /* widest level (rows) */
  int     G;                         /* lanes per row in the sweeps */
  int    *h_bi, *h_bj, *h_bdiag;     /* host copy of the layout (kept for GetFactor) */
  int     factored;
};

#define ILU_TPB 256
__global__ void ilu_pack_cols_kernel(int nslot, int G, bool upper, const int *__restrict__ order, const int *__restrict__ slotof, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, int *__restrict__ pkcol);
__global__ void ilu_pack_vals_kernel(int nslot, int G, bool upper, const int *__restrict__ order, const int *__restrict__ bi, const int *__restrict__ bdiag, const double *__restrict__ ba, double *__restrict__ pkval, double *__restrict__ pkdinv);
/* PETSCB200_ILU_MARCH=1 selects the segment-marching sweeps for ILU(0).  Measured on the B200 (profiles/round2_notes.md): bit-exact,
   but 8.05 ms vs 7.09 ms per PCApply on the 27-point 256^3 operator and 5x slower on the 7-point one -- a marching line catches up
   with its producer line and then pays an L2 round trip on every row, and the 32/G lines of a warp stall each other in lockstep.
   The level-scheduled pipe kernels stay the default; the marching kernel serves the ICC(0) sweeps (modes 2 and 3). */
static int   g_ilu_march = 0;

__device__ __forceinline__ int ld_acquire(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void wait_ready(const int *flag, int row, int epoch)
{
  unsigned ns = 0; /* numeric factorisation only (setup path); same back-off as wait_value */
  while (ld_acquire(flag + row) != epoch) {
    if (ns) __nanosleep(ns);
    ns = ns ? (ns < 512 ? ns * 2 : 512) : 32;
  }
}

#define ILU_BATCH 4
__device__ __forceinline__ int warp_ticket(int *ticket)
{
  int t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(ticket, 1);
  return __shfl_sync(0xffffffffu, t, 0);
}


/* ------------------------------------------------------------------ triangular sweeps */
/* lower: t[i] = b[i] - sum_{k in L(i,:)} ba[k] t[bj[k]]                          (aijfact.c:2431-2440)
   upper: x[i] = (t[i] - sum_{k in U(i,:) strict} ba[k] x[bj[k]]) * ba[bdiag[i]]   (aijfact.c:2443-2451)

   Readiness travels WITH the value: the output vector of a sweep is pre-filled with a sentinel bit pattern
   (0xFF..FF, one cudaMemsetAsync) and a consumer spins on the 8-byte entry itself until it is no longer the sentinel.
   An aligned 8-byte store is single-copy atomic, so there is no separate flag, no release fence on the producer and no
   acquire/L1-invalidate on the consumer: the per-level critical path is one L2 store + one L2 load.  (The first
   version used a flag + __threadfence + ld.acquire per row and ran ~17 us per level; see profiles/round1_notes.md.) */
#define ILU_SENTINEL 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const double *p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double *p, double v)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  if (u == ILU_SENTINEL) u = 0x7FFFFFFFFFFFFFFFull; /* a NaN that happens to equal the sentinel: store the canonical NaN */
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(u) : "memory");
}
__device__ int g_ilu_backoff_ns = 0; /* tunable: PETSCB200_ILU_BACKOFF_NS (0 = pure spinning; measured best) */
static int     g_ilu_lookahead = 2, g_ilu_batch = 1, g_ilu_pipe = 1; /* tunables: PETSCB200_ILU_LOOKAHEAD, PETSCB200_ILU_BATCH */

/* Warp-convergent wait: every lane polls its own dependency (or nothing), the warp loops until ALL its lanes have seen a
   value.  The rows of one warp-chunk belong to one level, so they become ready together; a convergent loop issues one
   poll wavefront and at most one nanosleep per iteration for the whole warp (divergent per-lane spin loops serialise). */
__device__ __forceinline__ double wait_value_warp(const double *p, bool active)
{
  unsigned long long v     = 0;
  bool               ready = !active;
  unsigned           ns = 0, spins = 0;
  const unsigned     cap = (unsigned)g_ilu_backoff_ns;
  for (;;) {
    if (!ready) {
      v     = ld_relaxed_u64(p);
      ready = (v != ILU_SENTINEL);
    }
    if (__all_sync(0xffffffffu, ready)) break;
    if (++spins > 4 && cap) {
      ns = ns ? (ns * 2 < cap ? ns * 2 : cap) : 32;
      __nanosleep(ns);
    }
  }
  return __longlong_as_double((long long)v);
}

template <int G, bool UPPER>
__device__ __forceinline__ void ilu_sweep_row(int slot, int nslot, const int *__restrict__ order, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, const double *__restrict__ ba, const double *__restrict__ rhs, double *out)
{
  /* warp-convergent: groups without a row (padding, tail) run the same control flow with an empty range */
  const int      gl    = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int      i     = slot < nslot ? order[slot] : -1;
  const bool     valid = i >= 0;
  int            ks = 0, ke = 0;
  double         sum = 0.0, dinv = 0.0;
  if (valid) {
    sum = rhs[i]; /* lower: b[i]; upper: the lower sweep's t[i] (complete: previous kernel) */
    if (!UPPER) {
      ks = bi[i];
      ke = bi[i + 1];
    } else {
      ks   = bdiag[i + 1] + 1;
      ke   = bdiag[i];
      dinv = ba[ke];
    }
  }
  int nchunk = (ke - ks + G - 1) / G;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nchunk = max(nchunk, __shfl_xor_sync(0xffffffffu, nchunk, o));
  for (int c = 0; c < nchunk; c++) {
    const int  k0  = ks + c * G;
    const int  k   = k0 + gl;
    const bool act = k < ke;
    const double a = act ? ba[k] : 0.0; /* factor entry fetched while the dependency is still in flight */
    double       p = wait_value_warp(out + (act ? bj[k] : 0), act);
    if (act) p = __dmul_rn(a, p);
    const int cnt = min(G, ke - k0);
#pragma unroll
    for (int l = 0; l < G; l++) {
      const double pl = __shfl_sync(gmask, p, l, G);
      if (l < cnt) sum = __dsub_rn(sum, pl); /* strict left-to-right, FMA-free */
    }
  }
  if (valid && gl == 0) {
    if (UPPER) sum = __dmul_rn(sum, dinv);
    st_relaxed_f64(out + i, sum);
  }
}

/* Persistent warps: every warp repeatedly takes the next batch of ILU_BATCH warp-chunks (a chunk = 32/G consecutive slots
   of the level order, never straddling a level) through one atomic ticket.  Tickets are handed out in dependency order,
   so whatever a warp waits for is owned by a warp that already holds a ticket; no CTA is launched per chunk (the first
   version launched ~1 M CTAs per sweep and was bound by CTA dispatch, 19 us per level; profiles/round1_notes.md). */
template <int G, bool UPPER>
__global__ void __launch_bounds__(ILU_TPB) ilu_sweep_kernel(int nslot, const int *__restrict__ order, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, const double *__restrict__ ba, const double *__restrict__ rhs, double *out, int *ticket, int batch)
{
  constexpr int RPW = 32 / G; /* rows per warp-chunk */
  for (;;) {
    const int64_t base = (int64_t)warp_ticket(ticket) * (batch * RPW);
    if (base >= nslot) break;
#pragma unroll 1
    for (int q = 0; q < batch; q++) {
      const int64_t slot = base + q * RPW + (threadIdx.x & 31) / G;
      ilu_sweep_row<G, UPPER>((int)(slot < nslot ? slot : nslot), nslot, order, bi, bdiag, bj, ba, rhs, out);
      __syncwarp();
    }
  }
}

/* ------------------------------------------------------------------ numeric factorisation */
/* position of column c inside a sorted segment bj[lo,hi), or -1 */
__device__ __forceinline__ int find_col(const int *__restrict__ bj, int lo, int hi, int c)
{
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    int v   = bj[mid];
    if (v == c) return mid;
    if (v < c) lo = mid + 1;
    else hi = mid;
  }
  return -1;
}

template <int G>
__device__ __forceinline__ void ilu_numeric_row(int slot, int nslot, const int *__restrict__ order, const int *__restrict__ ai, const int *__restrict__ adiag, const double *__restrict__ aval, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, double *ba, double shift, double zeropivot, int *flag, int epoch, int *status)
{
  const int      gl    = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  if (slot >= nslot) return;
  const int i = order[slot];
  if (i < 0) return;
  const int l0 = bi[i], l1 = bi[i + 1];             /* L(i,:) */
  const int u0 = bdiag[i + 1] + 1, u1 = bdiag[i];    /* U(i,:) strict part, diagonal at u1 */
  const int a0 = ai[i], ad = adiag[i], a1 = ai[i + 1];
  /* load the unfactored row into the factor storage (aijfact.c:282-286) */
  for (int k = gl; k < l1 - l0; k += G) ba[l0 + k] = aval[a0 + k];
  for (int k = gl; k < u1 - u0; k += G) ba[u0 + k] = aval[ad + 1 + k];
  if (gl == 0) ba[u1] = aval[ad] + shift;
  (void)a1;
  __syncwarp(gmask);
  /* elimination (aijfact.c:289-306): pivot rows in ascending column order */
  for (int kl = l0; kl < l1; kl++) {
    const int row = bj[kl];
    wait_ready(flag, row, epoch);
    const double pc = __ldcg(ba + kl);
    if (pc != 0.0) {
      const double mult = __dmul_rn(pc, __ldcg(ba + bdiag[row]));
      const int    r0 = bdiag[row + 1] + 1, r1 = bdiag[row]; /* U(row,:) without its diagonal */
      for (int t = r0 + gl; t < r1; t += G) {
        const int c = bj[t];
        int       pos;
        if (c < i) pos = find_col(bj, kl + 1, l1, c);
        else if (c == i) pos = u1;
        else pos = find_col(bj, u0, u1, c);
        if (pos >= 0) ba[pos] = __dsub_rn(__ldcg(ba + pos), __dmul_rn(mult, __ldcg(ba + t)));
      }
      __syncwarp(gmask);
      if (gl == 0) ba[kl] = mult;
    }
    __syncwarp(gmask);
  }
  /* pivot check (MatPivotCheck_nz, matimpl.h:795-811) and inversion of the diagonal (aijfact.c:332-334) */
  double rs = 0.0;
  for (int k = l0 + gl; k < l1; k += G) rs += fabs(__ldcg(ba + k));
  for (int k = u0 + gl; k < u1; k += G) rs += fabs(__ldcg(ba + k));
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) rs += __shfl_xor_sync(gmask, rs, o, G);
  if (gl == 0) {
    const double pv = __ldcg(ba + u1);
    if (fabs(pv) <= zeropivot * rs && !isnan(pv)) atomicMax(status, 1);
    ba[u1] = 1.0 / pv;
    __threadfence();
    st_release(flag + i, epoch);
  }
}

template <int G>
__global__ void __launch_bounds__(ILU_TPB) ilu_numeric_kernel(int nslot, const int *__restrict__ order, const int *__restrict__ ai, const int *__restrict__ adiag, const double *__restrict__ aval, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, double *ba, double shift, double zeropivot, int *flag, int epoch, int *ticket, int *status)
{
  constexpr int RPW = 32 / G;
  for (;;) {
    const int64_t base = (int64_t)warp_ticket(ticket) * (ILU_BATCH * RPW);
    if (base >= nslot) break;
#pragma unroll 1
    for (int q = 0; q < ILU_BATCH; q++) {
      const int64_t slot = base + q * RPW + (threadIdx.x & 31) / G;
      ilu_numeric_row<G>((int)(slot < nslot ? slot : nslot), nslot, order, ai, adiag, aval, bi, bdiag, bj, ba, shift, zeropivot, flag, epoch, status);
      __syncwarp();
    }
  }
}

/* ------------------------------------------------------------------ host: symbolic + level schedule */
static int *levels_to_order(int n, const int *lev, int nlev, int rpw, int *nslot_out, int *maxwidth_out)
{
  /* counting sort by level, each level padded to a multiple of rpw (rows per warp) with -1 */
  int64_t *cnt = (int64_t *)calloc((size_t)nlev + 1, sizeof(int64_t));
  for (int i = 0; i < n; i++) cnt[lev[i] + 1]++;
  int64_t tot = 0, mw = 0;
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nlev + 1));
  for (int l = 0; l < nlev; l++) {
    if (cnt[l + 1] > mw) mw = cnt[l + 1];
    start[l] = tot;
    tot += (cnt[l + 1] + rpw - 1) / rpw * rpw;
  }
  if (tot > 2147483000LL) {
    free(cnt);
    free(start);
    return NULL;
  }
  int *order = (int *)malloc(sizeof(int) * (size_t)(tot + 1));
  for (int64_t k = 0; k < tot; k++) order[k] = -1;
  for (int i = 0; i < n; i++) order[start[lev[i]]++] = i;
  *nslot_out    = (int)tot;
  *maxwidth_out = (int)mw;
  free(cnt);
  free(start);
  return order;
}

extern "C" int b200Ilu0Destroy(b200IluPlan p)
{
  if (!p) return 0;
  cudaFree(p->d_ai); cudaFree(p->d_adiag); cudaFree(p->d_bi); cudaFree(p->d_bj); cudaFree(p->d_bdiag); cudaFree(p->d_ba);
  cudaFree(p->d_segL); cudaFree(p->d_segU);
  cudaFree(p->d_slotL); cudaFree(p->d_slotU); cudaFree(p->d_pkcolL); cudaFree(p->d_pkcolU); cudaFree(p->d_pkvalL); cudaFree(p->d_pkvalU); cudaFree(p->d_pkdinvU);
  cudaFree(p->d_tL); cudaFree(p->d_xU); cudaFree(p->d_mapLU);
  cudaFree(p->d_orderL); cudaFree(p->d_orderU); cudaFree(p->d_metaL); cudaFree(p->d_metaU); cudaFree(p->d_flag); cudaFree(p->d_ticket); cudaFree(p->d_tmp);
  free(p->h_bi); free(p->h_bj); free(p->h_bdiag);
  free(p);
  return 0;
}

extern "C" int b200Ilu0Symbolic(b200Handle h, int n, const int *ai, const int *aj, b200IluPlan *plan)
{
  B200_CHECK(h && plan, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  B200_CHECK(n == 0 || (ai && aj), B200_ERR_ARG_NULL, "null pattern");
  b200IluPlan p = (b200IluPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->n   = n;
  p->nnz = n ? ai[n] : 0;
  const int64_t nnz = p->nnz;
  int *adiag = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  int *bi = (int *)malloc(sizeof(int) * (size_t)(n + 1)), *bdiag = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  int *bj = (int *)malloc(sizeof(int) * (size_t)(nnz + 1));
  int *levL = (int *)malloc(sizeof(int) * (size_t)(n + 1)), *levU = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  B200_CHECK(adiag && bi && bdiag && bj && levL && levU, B200_ERR_MEM, "out of host memory");
  /* diagonal markers (MatGetDiagonalMarkers_SeqAIJ): ILU needs every diagonal entry present */
  for (int i = 0; i < n; i++) {
    int lo = ai[i], hi = ai[i + 1], pos = -1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (aj[mid] == i) { pos = mid; break; }
      if (aj[mid] < i) lo = mid + 1;
      else hi = mid;
    }
    if (pos < 0) {
      free(adiag); free(bi); free(bdiag); free(bj); free(levL); free(levU); free(p);
      B200_CHECK(0, B200_ERR_ARG_WRONGSTATE, "Matrix is missing diagonal entry %d", i); /* aij.c MatMissingDiagonal */
    }
    adiag[i] = pos;
  }
  /* aijfact.c:1503-1521 */
  int64_t k = 0;
  bi[0]     = 0;
  for (int i = 0; i < n; i++) {
    int nz    = adiag[i] - ai[i];
    bi[i + 1] = bi[i] + nz;
    memcpy(bj + k, aj + ai[i], sizeof(int) * (size_t)nz);
    k += nz;
  }
  bdiag[n] = bi[n] - 1;
  for (int i = n - 1; i >= 0; i--) {
    int nz = ai[i + 1] - adiag[i] - 1;
    memcpy(bj + k, aj + adiag[i] + 1, sizeof(int) * (size_t)nz);
    k += nz;
    bj[k++]  = i;
    bdiag[i] = bdiag[i + 1] + nz + 1;
  }
  /* dependency levels */
  int nlevL = 0, nlevU = 0;
  for (int i = 0; i < n; i++) {
    int l = 0;
    for (int q = ai[i]; q < adiag[i]; q++) l = levL[aj[q]] + 1 > l ? levL[aj[q]] + 1 : l;
    levL[i] = l;
    if (l + 1 > nlevL) nlevL = l + 1;
  }
  for (int i = n - 1; i >= 0; i--) {
    int l = 0;
    for (int q = adiag[i] + 1; q < ai[i + 1]; q++) l = levU[aj[q]] + 1 > l ? levU[aj[q]] + 1 : l;
    levU[i] = l;
    if (l + 1 > nlevU) nlevU = l + 1;
  }
  p->nlevL = nlevL;
  p->nlevU = nlevU;
  /* lanes per row for the sweeps: about the mean number of off-diagonal entries per triangular row */
  {
    double avg = n ? (double)(nnz - n) / (2.0 * n) : 0.0;
    int    G   = 2;
    while (G < 32 && G < avg) G <<= 1;
    p->G = G;
  }
  /* the numeric kernel uses 8 lanes per row; pad levels for the larger of the two groupings */
  const int rpw = 32 / (p->G < 8 ? p->G : 8); /* a multiple of the sweeps' rows-per-warp (32/G) as well */
  int *orderL = levels_to_order(n, levL, nlevL, rpw, &p->nslotL, &p->maxwL);
  int *orderU = levels_to_order(n, levU, nlevU, rpw, &p->nslotU, &p->maxwU);
  free(levL); free(levU);
  if (!orderL || !orderU) {
    free(orderL); free(orderU); free(adiag); free(bi); free(bdiag); free(bj); free(p);
    B200_CHECK(0, B200_ERR_SUP, "level schedule exceeds 32-bit indexing");
  }
#define UP(dst, src, cnt, T) \
  do { \
    B200_CUDA(cudaMalloc(&p->dst, sizeof(T) * ((size_t)(cnt) + 64))); \
    B200_CUDA(cudaMemcpyAsync(p->dst, src, sizeof(T) * (size_t)(cnt), cudaMemcpyHostToDevice, h->stream)); \
  } while (0)
  UP(d_ai, ai, n + 1, int);
  UP(d_adiag, adiag, n, int);
  UP(d_bi, bi, n + 1, int);
  UP(d_bdiag, bdiag, n + 1, int);
  UP(d_bj, bj, nnz, int);
  UP(d_orderL, orderL, p->nslotL, int);
  UP(d_orderU, orderU, p->nslotU, int);
  { /* packed per-slot records for the sweeps: one coalesced 16-byte load replaces the order[] -> bi[]/bdiag[] chain */
    int4 *mL = (int4 *)malloc(sizeof(int4) * ((size_t)p->nslotL + 1)), *mU = (int4 *)malloc(sizeof(int4) * ((size_t)p->nslotU + 1));
    B200_CHECK(mL && mU, B200_ERR_MEM, "out of host memory");
    for (int s2 = 0; s2 < p->nslotL; s2++) {
      int i2 = orderL[s2];
      mL[s2] = i2 < 0 ? make_int4(-1, 0, 0, 0) : make_int4(i2, bi[i2], bi[i2 + 1], 0);
    }
    for (int s2 = 0; s2 < p->nslotU; s2++) {
      int i2 = orderU[s2];
      mU[s2] = i2 < 0 ? make_int4(-1, 0, 0, 0) : make_int4(i2, bdiag[i2 + 1] + 1, bdiag[i2], 0);
    }
    UP(d_metaL, mL, p->nslotL, int4);
    UP(d_metaU, mU, p->nslotU, int4);
    B200_CUDA(cudaStreamSynchronize(h->stream));
    free(mL); free(mU);
  }
#undef UP
  { /* segment schedule of the marching sweeps; GS lanes per row = next power of two >= the longest common triangular row,
       so that a whole row is one pipelined chunk (27-point: 13 -> 16, 7-point: 3 -> 4) */
    double avg = n ? (double)(nnz - n) / (2.0 * n) : 0.0;
    int    GS  = 2;
    while (GS < 32 && GS < avg + 0.5) GS <<= 1;
    p->GS = GS;
    const char *e1 = getenv("PETSCB200_ILU_SEG_MIN"), *e2 = getenv("PETSCB200_ILU_SEG_MAX");
    const int   minlen = e1 && atoi(e1) > 0 ? atoi(e1) : 8, maxlen = e2 && atoi(e2) > 0 ? atoi(e2) : 1024;
    int2 *sL = b200_build_segments(n, 0, bi, bj, 32 / GS, minlen, maxlen, &p->nsegslotL, &p->nseglevL);
    int2 *sU = b200_build_segments(n, 1, bdiag, bj, 32 / GS, minlen, maxlen, &p->nsegslotU, &p->nseglevU);
    B200_CUDA(cudaMalloc(&p->d_segL, sizeof(int2) * ((size_t)p->nsegslotL + 64)));
    B200_CUDA(cudaMalloc(&p->d_segU, sizeof(int2) * ((size_t)p->nsegslotU + 64)));
    B200_CUDA(cudaMemcpyAsync(p->d_segL, sL, sizeof(int2) * (size_t)p->nsegslotL, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaMemcpyAsync(p->d_segU, sU, sizeof(int2) * (size_t)p->nsegslotU, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    free(sL); free(sU);
  }
  { /* slot-space copy for the packed sweeps: possible when every triangular row fits the G lanes of a slot */
    int maxL = 0, maxU = 0;
    for (int i = 0; i < n; i++) {
      if (bi[i + 1] - bi[i] > maxL) maxL = bi[i + 1] - bi[i];
      const int lu = bdiag[i] - (bdiag[i + 1] + 1);
      if (lu > maxU) maxU = lu;
    }
    const size_t need = ((size_t)p->nslotL + (size_t)p->nslotU) * (size_t)p->G * 12 + ((size_t)p->nslotL + (size_t)p->nslotU) * 16 + (size_t)n * 8;
    size_t       fr = 0, totm = 0;
    cudaMemGetInfo(&fr, &totm);
    /* measured (profiles/round2_notes.md): the packed sweeps win where the levels are WIDE (throughput-bound: 7-point 512^3, 87 k
       rows per level: PCApply 16.5 -> 8.6 ms) and lose where they are narrow (latency-bound: 27-point 256^3, 9.4 k rows per level:
       7.1 -> 8.4 ms, padding 13 -> 16 entries and the same ~2 us per level) */
    const double rows_per_level = (double)n / (double)(nlevL > 0 ? nlevL : 1);
    const char  *ep = getenv("PETSCB200_ILU_PACKED_MIN_WIDTH");
    const double minw = ep ? atof(ep) : 24000.0;
    p->packed = (n > 0 && maxL <= p->G && maxU <= p->G && rows_per_level >= minw && need + ((size_t)2 << 30) < fr) ? 1 : 0;
    if (p->packed) {
      int *slotL = (int *)malloc(sizeof(int) * (size_t)(n + 1)), *slotU = (int *)malloc(sizeof(int) * (size_t)(n + 1));
      for (int s2 = 0; s2 < p->nslotL; s2++)
        if (orderL[s2] >= 0) slotL[orderL[s2]] = s2;
      for (int s2 = 0; s2 < p->nslotU; s2++)
        if (orderU[s2] >= 0) slotU[orderU[s2]] = s2;
      B200_CUDA(cudaMalloc(&p->d_slotL, sizeof(int) * ((size_t)n + 64)));
      B200_CUDA(cudaMalloc(&p->d_slotU, sizeof(int) * ((size_t)n + 64)));
      B200_CUDA(cudaMemcpyAsync(p->d_slotL, slotL, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(p->d_slotU, slotU, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
      B200_CUDA(cudaMalloc(&p->d_pkcolL, sizeof(int) * ((size_t)p->nslotL * p->G + 64)));
      B200_CUDA(cudaMalloc(&p->d_pkcolU, sizeof(int) * ((size_t)p->nslotU * p->G + 64)));
      B200_CUDA(cudaMalloc(&p->d_pkvalL, sizeof(double) * ((size_t)p->nslotL * p->G + 64)));
      B200_CUDA(cudaMalloc(&p->d_pkvalU, sizeof(double) * ((size_t)p->nslotU * p->G + 64)));
      B200_CUDA(cudaMalloc(&p->d_pkdinvU, sizeof(double) * ((size_t)p->nslotU + 64)));
      B200_CUDA(cudaMalloc(&p->d_tL, sizeof(double) * ((size_t)p->nslotL + 64)));
      B200_CUDA(cudaMalloc(&p->d_xU, sizeof(double) * ((size_t)p->nslotU + 64)));
      {
        int *mapLU = (int *)malloc(sizeof(int) * ((size_t)p->nslotU + 1));
        for (int s2 = 0; s2 < p->nslotU; s2++) mapLU[s2] = orderU[s2] >= 0 ? slotL[orderU[s2]] : -1;
        B200_CUDA(cudaMalloc(&p->d_mapLU, sizeof(int) * ((size_t)p->nslotU + 64)));
        B200_CUDA(cudaMemcpyAsync(p->d_mapLU, mapLU, sizeof(int) * (size_t)p->nslotU, cudaMemcpyHostToDevice, h->stream));
        B200_CUDA(cudaStreamSynchronize(h->stream));
        free(mapLU);
      }
      const int g = h->num_sms * 8;
      ilu_pack_cols_kernel<<<g, 256, 0, h->stream>>>(p->nslotL, p->G, false, p->d_orderL, p->d_slotL, p->d_bi, p->d_bdiag, p->d_bj, p->d_pkcolL);
      ilu_pack_cols_kernel<<<g, 256, 0, h->stream>>>(p->nslotU, p->G, true, p->d_orderU, p->d_slotU, p->d_bi, p->d_bdiag, p->d_bj, p->d_pkcolU);
      B200_LAUNCHED(2);
      B200_CUDA(cudaStreamSynchronize(h->stream));
      free(slotL); free(slotU);
    }
  }
  B200_CUDA(cudaMalloc(&p->d_ba, sizeof(double) * ((size_t)nnz + 64)));
  B200_CUDA(cudaMalloc(&p->d_tmp, sizeof(double) * ((size_t)n + 64)));
  B200_CUDA(cudaMalloc(&p->d_flag, sizeof(int) * ((size_t)n + 64)));
  B200_CUDA(cudaMemsetAsync(p->d_flag, 0, sizeof(int) * ((size_t)n + 64), h->stream));
  B200_CUDA(cudaMalloc(&p->d_ticket, 64));
  B200_CUDA(cudaMemsetAsync(p->d_ticket, 0, 64, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  free(orderL); free(orderU); free(adiag);
  p->h_bi = bi; p->h_bj = bj; p->h_bdiag = bdiag;
  p->epoch = 0;
  *plan    = p;
  return 0;
}

/* Persistent grid sized to the parallelism that exists: about two levels' worth of rows in flight (warps that run
   further ahead only poll), between one CTA per SM and the residency limit of 8 x 256 threads per SM */
static int ilu_grid(b200Handle h, int maxwidth, int G)
{
  const int64_t rows_per_cta = (ILU_TPB / 32) * (32 / G);
  int64_t       g            = (g_ilu_lookahead * (int64_t)maxwidth + rows_per_cta - 1) / rows_per_cta;
  if (g < h->num_sms) g = h->num_sms;
  if (g > (int64_t)h->num_sms * 8) g = (int64_t)h->num_sms * 8;
  return (int)g;
}

template <int G>
static int numeric_launch(b200Handle h, b200IluPlan p, const double *aval, double shift, double zeropivot)
{
  const int grid = ilu_grid(h, p->maxwL, G);
  ilu_numeric_kernel<G><<<grid, ILU_TPB, 0, h->stream>>>(p->nslotL, p->d_orderL, p->d_ai, p->d_adiag, aval, p->d_bi, p->d_bdiag, p->d_bj, p->d_ba, shift, zeropivot, p->d_flag, p->epoch, p->d_ticket, p->d_ticket + 8);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

extern "C" int b200Ilu0Numeric(b200Handle h, b200IluPlan p, const double *d_aval, double zeropivot, double shiftamount, int *nshift_out)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (nshift_out) *nshift_out = 0;
  if (p->n == 0) {
    p->factored = 1;
    return 0;
  }
  B200_CHECK(d_aval, B200_ERR_ARG_NULL, "null values");
  double shift  = 0.0;
  int    nshift = 0;
  for (;;) {
    p->epoch++;
    B200_CUDA(cudaMemsetAsync(p->d_ticket, 0, 64, h->stream));
    int rc = (p->G >= 8) ? numeric_launch<8>(h, p, d_aval, shift, zeropivot) : (p->G == 4 ? numeric_launch<4>(h, p, d_aval, shift, zeropivot) : numeric_launch<2>(h, p, d_aval, shift, zeropivot));
    if (rc) return rc;
    int status = 0;
    B200_CUDA(cudaMemcpyAsync(&status, p->d_ticket + 8, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (!status) break;
    /* MatPivotCheck_nz: first shift = shiftamount, then doubled; refactor from scratch (aijfact.c:264 do-while) */
    shift = nshift ? shift * 2.0 : shiftamount;
    nshift++;
    B200_CHECK(nshift <= 60 && shift > 0.0, B200_ERR_MAT_LU_ZRPVT, "Zero pivot in ILU(0) factorisation; shift could not repair it");
  }
  if (nshift_out) *nshift_out = nshift;
  if (p->packed) { /* slot-major copy of the factor values for the packed sweeps */
    const int g = h->num_sms * 8;
    ilu_pack_vals_kernel<<<g, 256, 0, h->stream>>>(p->nslotL, p->G, false, p->d_orderL, p->d_bi, p->d_bdiag, p->d_ba, p->d_pkvalL, NULL);
    ilu_pack_vals_kernel<<<g, 256, 0, h->stream>>>(p->nslotU, p->G, true, p->d_orderU, p->d_bi, p->d_bdiag, p->d_ba, p->d_pkvalU, p->d_pkdinvU);
    B200_LAUNCHED(2);
    B200_KERNEL_CHECK();
  }
  p->factored = 1;
  return 0;
}

/* Software-pipelined sweep: every warp keeps THREE tickets in flight -- stage A (two ahead): ticket + packed row record;
   stage B (one ahead): the first G factor entries, their column indices, the right-hand side and the inverted diagonal;
   stage C: wait for the dependencies, accumulate strictly left to right, publish.  The three dependent global-memory
   round trips of a chunk (ticket, record, entries) are thereby overlapped with the previous chunks' waits; what stays on
   the per-level critical path is the poll + the ordered subtraction + the publishing store.  Deadlock-free: a warp works
   through its tickets in increasing order and only ever waits on rows with a smaller slot. */
template <int G, bool UPPER>
__global__ void __launch_bounds__(ILU_TPB) ilu_sweep_pipe_kernel(int nslot, const int4 *__restrict__ meta, const int *__restrict__ bj, const double *__restrict__ ba, const double *__restrict__ rhs, double *out, int *ticket)
{
  constexpr int  RPW   = 32 / G;
  const int      gl    = threadIdx.x % G;
  const int      grp   = (threadIdx.x & 31) / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  auto record = [&](int tk) -> int4 {
    const int64_t slot = (int64_t)tk * RPW + grp;
    return slot < nslot ? __ldg(meta + slot) : make_int4(-1, 0, 0, 0);
  };
  /* chunk sequence of this warp: with a co-resident grid (cooperative launch) the static round-robin w, w+W, w+2W, ...
     needs no atomics at all -- a single ticket counter serialises at ~0.6 G atomics/s, which was the whole cost of the
     ticketed version at 256^3 (8.4 M tickets = 14 ms); ticket != NULL keeps the dynamic fallback */
  const int wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), W = (int)((gridDim.x * blockDim.x) >> 5);
  int       seq = 0;
  auto next_chunk = [&]() -> int {
    if (ticket) return warp_ticket(ticket);
    const int64_t t = (int64_t)wid + (int64_t)(seq++) * W;
    return t > 2147483000LL ? 2147483000 : (int)t;
  };
  int  tkC = next_chunk();
  int4 mC  = record(tkC);
  int  tkB = next_chunk();
  int4 mB  = record(tkB);
  /* stage-B loads of the first chunk */
  int    cC = 0;
  double aC = 0.0, rC = 0.0, dC = 0.0;
  if (mC.x >= 0) {
    rC = rhs[mC.x];
    if (UPPER) dC = ba[mC.z];
    if (mC.y + gl < mC.z) {
      cC = bj[mC.y + gl];
      aC = ba[mC.y + gl];
    }
  }
  for (;;) {
    if ((int64_t)tkC * RPW >= nslot) break; /* warp-uniform */
    /* stage A for the chunk two ahead, stage B for the chunk one ahead */
    const int  tkA = next_chunk();
    const int4 mA  = record(tkA);
    int        cB = 0;
    double     aB = 0.0, rB = 0.0, dB = 0.0;
    if (mB.x >= 0) {
      rB = rhs[mB.x];
      if (UPPER) dB = ba[mB.z];
      if (mB.y + gl < mB.z) {
        cB = bj[mB.y + gl];
        aB = ba[mB.y + gl];
      }
    }
    /* stage C */
    {
      const bool valid = mC.x >= 0;
      const int  ks = mC.y, ke = valid ? mC.z : mC.y;
      double     sum = rC;
      int        nchunk = (ke - ks + G - 1) / G;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nchunk = max(nchunk, __shfl_xor_sync(0xffffffffu, nchunk, o));
      for (int c = 0; c < nchunk; c++) {
        const int  k0  = ks + c * G, k = k0 + gl;
        const bool act = k < ke;
        int        col = cC;
        double     a   = aC;
        if (c > 0 && act) { /* rows longer than G entries: later chunks are fetched on the fly */
          col = bj[k];
          a   = ba[k];
        }
        double p = wait_value_warp(out + (act ? col : 0), act);
        if (act) p = __dmul_rn(a, p);
        const int cnt = min(G, ke - k0);
#pragma unroll
        for (int l = 0; l < G; l++) {
          const double pl = __shfl_sync(gmask, p, l, G);
          if (l < cnt) sum = __dsub_rn(sum, pl); /* strict left-to-right, FMA-free */
        }
      }
      if (valid && gl == 0) {
        if (UPPER) sum = __dmul_rn(sum, dC);
        st_relaxed_f64(out + mC.x, sum);
      }
    }
    __syncwarp();
    tkC = tkB; mC = mB; cC = cB; aC = aB; rC = rB; dC = dB;
    tkB = tkA; mB = mA;
  }
}


/* ------------------------------------------------------------------ segment-marching sweeps (round 2)
   The level-scheduled kernels above pay one L2 round trip per dependency LEVEL (~2 us x 3572 levels on the 27-point 256^3
   operator: latency, not bandwidth).  Here the unit of scheduling is a SEGMENT: a maximal run of consecutive rows in which
   every row depends on its predecessor (an x-line of a lexicographically ordered grid; capped and floored in length).  A group
   of G lanes MARCHES through its segment row after row:
     * the dependency on the rows just computed (the chain that makes a line sequential) never leaves the SM: the last RING
       results of the segment sit in a shared-memory ring, so the per-row critical path is ld.shared -> dmul -> dsub -> st.shared;
     * dependencies on OTHER segments are read from global memory with the value-as-flag protocol (sentinel-filled output,
       ld.relaxed.gpu poll), issued one row AHEAD of their use -- they were produced a whole segment-level earlier, so the
       poll normally succeeds at once and its latency is off the critical path;
     * the factor entries of a segment are one contiguous range (L rows ascending; U rows are stored descending, and the upper
       sweep marches descending): they are streamed two rows ahead with an L2 prefetch further out.
   Segments are scheduled in dependency-level order of the SEGMENT graph (766 levels instead of 1786 row levels for the 27-point
   operator) round-robin over a co-resident grid; a group only ever waits for rows of segments that precede its own in that
   order, so the sweep cannot deadlock.  Per row the sum is still accumulated strictly left to right with __dmul_rn/__dsub_rn:
   bit-identical to MatSolve_SeqAIJ_NaturalOrdering (aijfact.c:2413-2457). */
#define ILU_RING 32
/* MODE 0: ILU lower sweep   rows ascending,  entries [bi[i], bi[i+1]) ascending,            sum = b[i] - ...           (aijfact.c:2431-2440)
   MODE 1: ILU upper sweep   rows descending, entries [bdiag[i+1]+1, bdiag[i]) ascending,    sum = (t[i] - ...) * ba[bdiag[i]]   (aijfact.c:2443-2451)
   MODE 2: ICC forward       rows ascending over the COLUMN view (tptr/trow/tval),           y[c] = b[c] + sum tval*y[i];  out2[c] = y[c]*dinv[c]
   MODE 3: ICC backward      rows descending, entries [ui[i], ui[i+1]-1) taken DESCENDING,   x[i] = xf[i] + sum ua*x[col]
           (MatSolve_SeqSBAIJ_1_NaturalOrdering, sbaijfact2.c:2030-2065, as gathers; see icc.cu) */
template <int G, int MODE>
__global__ void __launch_bounds__(ILU_TPB) sweep_march_kernel(int nslot, const int2 *__restrict__ segs, const int *__restrict__ ext, const int *__restrict__ bj, const double *__restrict__ ba,
                                                              const double *__restrict__ rhs, double *out, int nnz_total, const double *__restrict__ dinv, double *out2)
{
  constexpr bool    DOWN = (MODE == 1 || MODE == 3), DESC = (MODE == 3), ADD = (MODE >= 2);
  constexpr int     RPW  = 32 / G;
  __shared__ double ring_all[ILU_TPB / G][ILU_RING];
  const int         gl = threadIdx.x % G, grp = (threadIdx.x & 31) / G;
  const unsigned    gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  double           *ring = ring_all[threadIdx.x / G];
  const int         wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), W = (int)((gridDim.x * blockDim.x) >> 5);
  for (int64_t chunk = wid; chunk * RPW < nslot; chunk += W) {
    const int64_t slot = chunk * RPW + grp;
    const int2    sg   = slot < nslot ? __ldg(segs + slot) : make_int2(0, 0);
    const int     f = sg.x, cnt = sg.y;
    int           maxc = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
    auto rowidx = [&](int r) { return DOWN ? f - r : f + r; };
    /* pipeline registers: E = entries of row r+2, P = polled values of row r+1, C = row r */
    int    ksC = 0, keC = 0, colC = -1, ksP = 0, keP = 0, colP = -1, ksE = 0, keE = 0, colE = -1;
    double aC = 0, aP = 0, aE = 0, rhsC = 0, rhsP = 0, rhsE = 0, vC = 0, vP = 0;
    bool   pollC = false, pollP = false;
    auto load_extent = [&](int r, int &ks, int &ke, double &rh) {
      if (r < cnt) {
        const int i = rowidx(r);
        if (MODE == 1) { ks = __ldg(ext + i + 1) + 1; ke = __ldg(ext + i); }
        else { ks = __ldg(ext + i); ke = __ldg(ext + i + 1) - (MODE == 3 ? 1 : 0); }
        rh = rhs[i];
      } else { ks = ke = 0; rh = 0.0; }
    };
    auto entry_pos = [&](int ks, int ke, int c) { return DESC ? ke - 1 - (c * G + gl) : ks + c * G + gl; };
    auto load_entries = [&](int ks, int ke, int &col, double &a) {
      const int k = entry_pos(ks, ke, 0);
      if (k >= ks && k < ke) { col = __ldg(bj + k); a = __ldg(ba + k); }
      else { col = -1; a = 0.0; }
    };
    /* is column c of row i served by the ring (computed by this group within the last ILU_RING rows)? */
    auto in_ring = [&](int c, int i) { return DOWN ? (c <= f && c - i <= ILU_RING) : (c >= f && i - c <= ILU_RING); };
    auto issue_poll = [&](int r, int col, bool &poll, double &v) {
      poll = false;
      if (col >= 0 && !in_ring(col, rowidx(r))) {
        poll = true;
        v    = __longlong_as_double((long long)ld_relaxed_u64(out + col));
      }
    };
    /* prologue */
    load_extent(0, ksC, keC, rhsC);
    load_entries(ksC, keC, colC, aC);
    issue_poll(0, colC, pollC, vC);
    load_extent(1, ksP, keP, rhsP);
    load_entries(ksP, keP, colP, aP);
    for (int r = 0; r < maxc; r++) {
      /* stage E: extents + first G entries of row r+2; L2 prefetch ~8 rows further along the march */
      load_extent(r + 2, ksE, keE, rhsE);
      load_entries(ksE, keE, colE, aE);
      if (r + 2 < cnt) {
        const int len = keE - ksE + (MODE == 1 || MODE == 3 ? 1 : 0);
        const int pf  = DESC ? max(ksE - 8 * len - gl, 0) : min(ksE + 8 * len + gl, nnz_total - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ba + pf));
        if ((gl & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(bj + pf));
      }
      /* stage P: polls of row r+1 (its columns arrived during the previous iteration) */
      issue_poll(r + 1, colP, pollP, vP);
      /* stage C: row r */
      {
        const bool valid = r < cnt;
        const int  i     = rowidx(r);
        double     sum   = rhsC;
        int        nch   = valid ? (keC - ksC + G - 1) / G : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nch = max(nch, __shfl_xor_sync(0xffffffffu, nch, o));
        for (int c = 0; c < nch; c++) {
          const int  k   = entry_pos(ksC, keC, c);
          const bool act = valid && k >= ksC && k < keC;
          int        col = colC;
          double     a = aC, v = vC;
          bool       poll = pollC;
          if (c > 0) { /* rows longer than G entries: later chunks are fetched on the fly */
            col  = act ? bj[k] : -1;
            a    = act ? ba[k] : 0.0;
            poll = false;
            if (act && !in_ring(col, i)) { poll = true; v = __longlong_as_double((long long)ld_relaxed_u64(out + col)); }
          }
          double p = 0.0;
          if (act && !poll) p = ring[col & (ILU_RING - 1)];
          { /* warp-convergent re-poll of the values that were not there yet */
            bool ready = !(act && poll) || ((unsigned long long)__double_as_longlong(v) != ILU_SENTINEL);
            while (!__all_sync(0xffffffffu, ready)) {
              if (!ready) {
                v     = __longlong_as_double((long long)ld_relaxed_u64(out + col));
                ready = ((unsigned long long)__double_as_longlong(v) != ILU_SENTINEL);
              }
            }
          }
          if (act && poll) p = v;
          p = act ? __dmul_rn(a, p) : 0.0;
          const int n_in = min(G, keC - ksC - c * G);
#pragma unroll
          for (int l = 0; l < G; l++) {
            const double pl = __shfl_sync(gmask, p, l, G);
            if (valid && l < n_in) sum = ADD ? __dadd_rn(sum, pl) : __dsub_rn(sum, pl); /* strict order, FMA-free */
          }
        }
        if (valid && gl == 0) {
          if (MODE == 1) sum = __dmul_rn(sum, __ldg(ba + keC));
          ring[i & (ILU_RING - 1)] = sum;
          st_relaxed_f64(out + i, sum);
          if (MODE == 2) out2[i] = __dmul_rn(sum, __ldg(dinv + i));
        }
      }
      __syncwarp();
      ksC = ksP; keC = keP; colC = colP; aC = aP; rhsC = rhsP; vC = vP; pollC = pollP;
      ksP = ksE; keP = keE; colP = colE; aP = aE; rhsP = rhsE;
    }
  }
}

/* launch of one marching sweep (shared with icc.cu): co-resident persistent grid; cudaLaunchCooperativeKernel fails rather than
   deadlocks if the grid does not fit */
template <int G, int MODE>
static int sweep_march_launch(b200Handle h, int nslot, const int2 *segs, const int *ext, const int *bj, const double *ba, const double *rhs, double *out, int64_t nnz, const double *dinv, double *out2)
{
  static int occ = 0;
  if (!occ) {
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_march_kernel<G, MODE>, ILU_TPB, 0));
    if (occ < 1) occ = 1;
  }
  const int rpc = (ILU_TPB / 32) * (32 / G); /* segments per CTA */
  int       g   = (nslot + rpc - 1) / rpc;
  if (g > occ * h->num_sms) g = occ * h->num_sms;
  if (g < 1) g = 1;
  int   nnz_total = (int)nnz;
  void *args[]    = {&nslot, &segs, &ext, &bj, &ba, &rhs, &out, &nnz_total, &dinv, &out2};
  B200_CUDA(cudaLaunchCooperativeKernel((void *)sweep_march_kernel<G, MODE>, dim3(g), dim3(ILU_TPB), args, 0, h->stream));
  B200_LAUNCHED(1);
  return 0;
}
int b200_sweep_march(b200Handle h, int G, int mode, int nslot, const int2 *segs, const int *ext, const int *bj, const double *ba, const double *rhs, double *out, int64_t nnz, const double *dinv, double *out2)
{
#define SM_CASE(g) \
  case g: \
    switch (mode) { \
    case 0: return sweep_march_launch<g, 0>(h, nslot, segs, ext, bj, ba, rhs, out, nnz, dinv, out2); \
    case 1: return sweep_march_launch<g, 1>(h, nslot, segs, ext, bj, ba, rhs, out, nnz, dinv, out2); \
    case 2: return sweep_march_launch<g, 2>(h, nslot, segs, ext, bj, ba, rhs, out, nnz, dinv, out2); \
    default: return sweep_march_launch<g, 3>(h, nslot, segs, ext, bj, ba, rhs, out, nnz, dinv, out2); \
    }
  switch (G) {
    SM_CASE(2)
    SM_CASE(4)
    SM_CASE(8)
    SM_CASE(16)
  default:
    SM_CASE(32)
  }
#undef SM_CASE
}

/* host: segments and their dependency levels.  A segment is a run of consecutive rows (in sweep direction) in which every row
   has its predecessor among its columns, cut at maxlen and merged up to minlen.  level(seg) = 1 + max level of the segments
   that hold its out-of-segment columns.  Output: int2 (first row, count) per slot in level order, each level padded to a
   multiple of rpw slots with empty segments. */
int2 *b200_build_segments(int n, int mode, const int *ext, const int *bj, int rpw, int minlen, int maxlen, int *nslot_out, int *nlev_out)
{
  /* row extents per sweep mode (see sweep_march_kernel) */
  const bool down = (mode == 1 || mode == 3);
#define EXT_KS(i) (mode == 1 ? ext[(i) + 1] + 1 : ext[i])
#define EXT_KE(i) (mode == 1 ? ext[i] : ext[(i) + 1] - (mode == 3 ? 1 : 0))
  int *segof = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  int *first = (int *)malloc(sizeof(int) * (size_t)(n + 1)), *count = (int *)malloc(sizeof(int) * (size_t)(n + 1)), *lev = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  int  ns = 0;
  /* pass 1: cut -- a row continues the segment when its predecessor in sweep direction is among its columns (columns are sorted:
     for a lower-type row that is the LAST entry, for an upper-type row the FIRST) */
  for (int q = 0; q < n; q++) {
    const int i = down ? n - 1 - q : q;
    bool      chained = false;
    if (q > 0) {
      const int ks = EXT_KS(i), ke = EXT_KE(i);
      if (ke > ks) chained = down ? (bj[ks] == i + 1) : (bj[ke - 1] == i - 1);
    }
    const bool newseg = q == 0 || (count[ns - 1] >= maxlen) || (!chained && count[ns - 1] >= minlen);
    if (newseg) { first[ns] = i; count[ns] = 0; ns++; }
    count[ns - 1]++;
    segof[i] = ns - 1;
  }
  /* pass 2: levels (segments are numbered in sweep order, so every dependency has a smaller number) */
  int nlev = 0;
  for (int s2 = 0; s2 < ns; s2++) {
    int l = 0;
    for (int r = 0; r < count[s2]; r++) {
      const int i = down ? first[s2] - r : first[s2] + r;
      const int ks = EXT_KS(i), ke = EXT_KE(i);
      for (int k = ks; k < ke; k++) {
        const int t = segof[bj[k]];
        if (t != s2 && lev[t] + 1 > l) l = lev[t] + 1;
      }
    }
    lev[s2] = l;
    if (l + 1 > nlev) nlev = l + 1;
  }
#undef EXT_KS
#undef EXT_KE
  /* pass 3: counting sort by level, padded */
  int64_t *cnt = (int64_t *)calloc((size_t)nlev + 2, sizeof(int64_t)), tot = 0;
  for (int s2 = 0; s2 < ns; s2++) cnt[lev[s2] + 1]++;
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nlev + 1));
  for (int l = 0; l < nlev; l++) {
    start[l] = tot;
    tot += (cnt[l + 1] + rpw - 1) / rpw * rpw;
  }
  int2 *out = (int2 *)malloc(sizeof(int2) * (size_t)(tot + 1));
  for (int64_t k = 0; k < tot; k++) out[k] = make_int2(0, 0);
  for (int s2 = 0; s2 < ns; s2++) out[start[lev[s2]]++] = make_int2(first[s2], count[s2]);
  *nslot_out = (int)tot;
  *nlev_out  = nlev;
  free(segof); free(first); free(count); free(lev); free(cnt); free(start);
  return out;
}


/* ------------------------------------------------------------------ packed ("level-set reordered") sweeps (round 2)
   ncu of the level-scheduled pipe kernel on the 7-point operator: 1.7-2.0 TB/s of DRAM traffic, only 1.5-1.7x the algorithmic
   bytes -- the L2 absorbs the scatter -- yet 5.5 us per level at 512^3: the limiter is L2 SECTOR THROUGHPUT.  Consecutive slots
   of a level are rows nx-1 apart, so every 8-byte access of a lane (factor entries, right-hand side, the three polled
   dependencies, the result) is its own 32-byte sector transaction: ~76 sectors per 8 rows.  Here the sweep runs entirely in
   SLOT space: factor entries are packed slot-major ([slot][G], padding col = -1 / value 0), dependencies are stored as SLOT
   numbers (the neighbours of consecutive slots are consecutive slots of the previous level: the polls coalesce too), and the
   right-hand sides / results are reached through per-slot maps inside the sweeps (b[row], t_L[L-slot of the row], x[row]).  Per row the same strict left-to-right FMA-free sum (a padded entry contributes an exact -(+0.0)): bit-identical. */
__global__ void ilu_pack_cols_kernel(int nslot, int G, bool upper, const int *__restrict__ order, const int *__restrict__ slotof, const int *__restrict__ bi, const int *__restrict__ bdiag, const int *__restrict__ bj, int *__restrict__ pkcol)
{
  const int64_t tot = (int64_t)nslot * G, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += stride) {
    const int s = (int)(q / G), e = (int)(q % G), i = order[s];
    int       c = -1;
    if (i >= 0) {
      const int ks = upper ? bdiag[i + 1] + 1 : bi[i], ke = upper ? bdiag[i] : bi[i + 1];
      if (ks + e < ke) c = slotof[bj[ks + e]];
    }
    pkcol[q] = c;
  }
}
__global__ void ilu_pack_vals_kernel(int nslot, int G, bool upper, const int *__restrict__ order, const int *__restrict__ bi, const int *__restrict__ bdiag, const double *__restrict__ ba, double *__restrict__ pkval, double *__restrict__ pkdinv)
{
  const int64_t tot = (int64_t)nslot * G, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += stride) {
    const int s = (int)(q / G), e = (int)(q % G), i = order[s];
    double    v = 0.0;
    if (i >= 0) {
      const int ks = upper ? bdiag[i + 1] + 1 : bi[i], ke = upper ? bdiag[i] : bi[i + 1];
      if (ks + e < ke) v = ba[ks + e];
      if (upper && e == 0) pkdinv[s] = ba[ke];
    } else if (upper && e == 0) pkdinv[s] = 0.0;
    pkval[q] = v;
  }
}
template <int G, bool UPPER>
__global__ void __launch_bounds__(ILU_TPB) ilu_sweep_packed_kernel(int nslot, const int *__restrict__ pkcol, const double *__restrict__ pkval, const double *__restrict__ pkdinv, const int *__restrict__ rhsmap,
                                                                   const double *__restrict__ rhs, double *out, const int *__restrict__ out2map, double *out2)
{
  /* rhsmap[s]: where slot s finds its right-hand side (lower sweep: the row, rhs = b; upper sweep: the row's slot of the lower
     sweep, rhs = t_L) -- folding the permutation passes b -> b_L and t_L -> t_U into the sweeps (as separate gather kernels they
     cost 4 ms of the 14 ms PCApply at 512^3).  out2/out2map (upper sweep): the result is also stored in natural order. */
  constexpr int  RPW  = 32 / G;
  const int      lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int64_t  wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, W = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t  nchunk = ((int64_t)nslot + RPW - 1) / RPW;
  /* three chunks in flight per warp (static round robin c, c+W, c+2W over a co-resident grid; a chunk only waits for earlier
     chunks): stage E = coalesced loads of the packed entries two chunks ahead, stage P = the dependency polls one chunk ahead
     (issued as soon as its slot numbers are there: in a wide level the values were published long ago, so the poll's L2 round
     trip overlaps the current chunk instead of sitting on the warp's critical path), stage C = re-poll what was not ready,
     ordered subtraction, publish.  Measured before this stage existed: 8 rows per warp per ~2 us = 24 G rows/s at 512^3. */
  int    colC = -1, colP = -1, colE = -1;
  double aC = 0, aP = 0, aE = 0, rC = 0, rP = 0, rE = 0, dC = 0, dP = 0, dE = 0, vC = 0, vP = 0;
  auto   load = [&](int64_t ch, int &col, double &a, double &r, double &d) {
    if (ch < nchunk) {
      const int64_t q = ch * 32 + lane, s = ch * RPW + grp;
      col = __ldg(pkcol + q);
      a   = __ldg(pkval + q);
      const int m = s < nslot ? __ldg(rhsmap + s) : -1;
      r   = m >= 0 ? rhs[m] : 0.0;
      if (UPPER) d = s < nslot ? __ldg(pkdinv + s) : 0.0;
    } else {
      col = -1;
      a = r = d = 0.0;
    }
  };
  auto poll = [&](int col) -> double { return col >= 0 ? __longlong_as_double((long long)ld_relaxed_u64(out + col)) : 0.0; };
  load(wid, colC, aC, rC, dC);
  load(wid + W, colP, aP, rP, dP);
  vC = poll(colC);
  for (int64_t c = wid; c < nchunk; c += W) {
    load(c + 2 * W, colE, aE, rE, dE);
    vP = poll(colP);
    {
      const bool act   = colC >= 0;
      double     v     = vC;
      bool       ready = !act || ((unsigned long long)__double_as_longlong(v) != ILU_SENTINEL);
      while (!__all_sync(0xffffffffu, ready)) { /* warp-convergent re-poll */
        if (!ready) {
          v     = __longlong_as_double((long long)ld_relaxed_u64(out + colC));
          ready = ((unsigned long long)__double_as_longlong(v) != ILU_SENTINEL);
        }
      }
      const double p   = act ? __dmul_rn(aC, v) : 0.0;
      double       sum = rC;
#pragma unroll
      for (int l = 0; l < G; l++) sum = __dsub_rn(sum, __shfl_sync(gmask, p, l, G)); /* strict left-to-right, FMA-free; padding: - (+0.0) */
      const int64_t s = c * RPW + grp;
      if (gl == 0 && s < nslot) {
        if (UPPER) sum = __dmul_rn(sum, dC);
        st_relaxed_f64(out + s, sum);
        if (out2) {
          const int i2 = __ldg(out2map + s);
          if (i2 >= 0) out2[i2] = sum;
        }
      }
    }
    __syncwarp();
    colC = colP; aC = aP; rC = rP; dC = dP; vC = vP;
    colP = colE; aP = aE; rP = rE; dP = dE;
  }
}
static int g_ilu_packed = 1; /* PETSCB200_ILU_PACKED=0: level-scheduled pipe kernels on the factor's own layout */

static int ilu_set_backoff(void)
{
  static int done = 0;
  if (!done) {
    const char *e = getenv("PETSCB200_ILU_BACKOFF_NS");
    if (e) {
      int v = atoi(e);
      B200_CUDA(cudaMemcpyToSymbol(g_ilu_backoff_ns, &v, sizeof(int)));
    }
    if ((e = getenv("PETSCB200_ILU_LOOKAHEAD")) && atoi(e) > 0) g_ilu_lookahead = atoi(e);
    if ((e = getenv("PETSCB200_ILU_BATCH")) && atoi(e) > 0) g_ilu_batch = atoi(e);
    if ((e = getenv("PETSCB200_ILU_PIPE"))) g_ilu_pipe = atoi(e);
    if ((e = getenv("PETSCB200_ILU_MARCH"))) g_ilu_march = atoi(e);
    if ((e = getenv("PETSCB200_ILU_PACKED"))) g_ilu_packed = atoi(e);
    done = 1;
  }
  return 0;
}

template <int G>
static int sweeps_launch(b200Handle h, b200IluPlan p, const double *b, double *x)
{
  if (ilu_set_backoff()) return B200_ERR_GPU;
  const int gridL = ilu_grid(h, p->maxwL, G), gridU = ilu_grid(h, p->maxwU, G);
  B200_CUDA(cudaMemsetAsync(p->d_ticket, 0, 64, h->stream));
  B200_CUDA(cudaMemsetAsync(p->d_tmp, 0xFF, sizeof(double) * (size_t)p->n, h->stream)); /* sentinel fill */
  B200_CUDA(cudaMemsetAsync(x, 0xFF, sizeof(double) * (size_t)p->n, h->stream));
  if (g_ilu_pipe) {
    /* co-resident persistent grid: cudaLaunchCooperativeKernel fails rather than deadlocks if the grid does not fit */
    static int occ = 0, coop = -1;
    if (!occ) {
      B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ilu_sweep_pipe_kernel<G, true>, ILU_TPB, 0));
      if (occ < 1) occ = 1;
      { /* experiments: PETSCB200_ILU_CTAS_PER_SM caps the resident grid of the level-scheduled sweeps */
        const char *e = getenv("PETSCB200_ILU_CTAS_PER_SM");
        if (e && atoi(e) > 0 && atoi(e) < occ) occ = atoi(e);
      }
      B200_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    }
    int gL = gridL, gU = gridU;
    if (gL > occ * h->num_sms) gL = occ * h->num_sms;
    if (gU > occ * h->num_sms) gU = occ * h->num_sms;
    int          *tkL = coop == 1 ? NULL : p->d_ticket, *tkU = coop == 1 ? NULL : p->d_ticket + 1;
    const double *rhsL = b, *rhsU = p->d_tmp;
    double       *outL = p->d_tmp, *outU = x;
    void *argsL[] = {&p->nslotL, &p->d_metaL, &p->d_bj, &p->d_ba, &rhsL, &outL, &tkL};
    void *argsU[] = {&p->nslotU, &p->d_metaU, &p->d_bj, &p->d_ba, &rhsU, &outU, &tkU};
    if (coop == 1) {
      B200_CUDA(cudaLaunchCooperativeKernel((void *)ilu_sweep_pipe_kernel<G, false>, dim3(gL), dim3(ILU_TPB), argsL, 0, h->stream));
      B200_CUDA(cudaLaunchCooperativeKernel((void *)ilu_sweep_pipe_kernel<G, true>, dim3(gU), dim3(ILU_TPB), argsU, 0, h->stream));
    } else {
      B200_CUDA(cudaLaunchKernel((void *)ilu_sweep_pipe_kernel<G, false>, dim3(gL), dim3(ILU_TPB), argsL, 0, h->stream));
      B200_CUDA(cudaLaunchKernel((void *)ilu_sweep_pipe_kernel<G, true>, dim3(gU), dim3(ILU_TPB), argsU, 0, h->stream));
    }
  } else {
    ilu_sweep_kernel<G, false><<<gridL, ILU_TPB, 0, h->stream>>>(p->nslotL, p->d_orderL, p->d_bi, p->d_bdiag, p->d_bj, p->d_ba, b, p->d_tmp, p->d_ticket, g_ilu_batch);
    B200_KERNEL_CHECK();
    ilu_sweep_kernel<G, true><<<gridU, ILU_TPB, 0, h->stream>>>(p->nslotU, p->d_orderU, p->d_bi, p->d_bdiag, p->d_bj, p->d_ba, p->d_tmp, x, p->d_ticket + 1, g_ilu_batch);
    B200_KERNEL_CHECK();
  }
  B200_LAUNCHED(2);
  return 0;
}

template <int G>
static int packed_launch(b200Handle h, b200IluPlan p, const double *b, double *x)
{
  static int occ = 0;
  if (!occ) {
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ilu_sweep_packed_kernel<G, true>, ILU_TPB, 0));
    if (occ < 1) occ = 1;
  }
  B200_CUDA(cudaMemsetAsync(p->d_tL, 0xFF, sizeof(double) * (size_t)p->nslotL, h->stream)); /* sentinel fill */
  B200_CUDA(cudaMemsetAsync(p->d_xU, 0xFF, sizeof(double) * (size_t)p->nslotU, h->stream));
  int gL = ilu_grid(h, p->maxwL, G), gU = ilu_grid(h, p->maxwU, G);
  if (gL > occ * h->num_sms) gL = occ * h->num_sms;
  if (gU > occ * h->num_sms) gU = occ * h->num_sms;
  const double *nuld = NULL, *rhsU = p->d_tL;
  const int    *nuli = NULL;
  double       *nulo = NULL;
  void *argsL[] = {&p->nslotL, &p->d_pkcolL, &p->d_pkvalL, &nuld, &p->d_orderL, &b, &p->d_tL, &nuli, &nulo};
  void *argsU[] = {&p->nslotU, &p->d_pkcolU, &p->d_pkvalU, &p->d_pkdinvU, &p->d_mapLU, &rhsU, &p->d_xU, &p->d_orderU, &x};
  B200_CUDA(cudaLaunchCooperativeKernel((void *)ilu_sweep_packed_kernel<G, false>, dim3(gL), dim3(ILU_TPB), argsL, 0, h->stream));
  B200_CUDA(cudaLaunchCooperativeKernel((void *)ilu_sweep_packed_kernel<G, true>, dim3(gU), dim3(ILU_TPB), argsU, 0, h->stream));
  B200_LAUNCHED(2);
  return 0;
}

static int march_launch(b200Handle h, b200IluPlan p, const double *b, double *x)
{
  B200_CUDA(cudaMemsetAsync(p->d_tmp, 0xFF, sizeof(double) * (size_t)p->n, h->stream)); /* sentinel fill */
  B200_CUDA(cudaMemsetAsync(x, 0xFF, sizeof(double) * (size_t)p->n, h->stream));
  int rc = b200_sweep_march(h, p->GS, 0, p->nsegslotL, p->d_segL, p->d_bi, p->d_bj, p->d_ba, b, p->d_tmp, p->nnz, NULL, NULL);
  if (rc) return rc;
  return b200_sweep_march(h, p->GS, 1, p->nsegslotU, p->d_segU, p->d_bdiag, p->d_bj, p->d_ba, p->d_tmp, x, p->nnz, NULL, NULL);
}

extern "C" int b200Ilu0Solve(b200Handle h, b200IluPlan p, const double *d_b, double *d_x)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(p->factored, B200_ERR_ORDER, "b200Ilu0Numeric must be called first");
  if (p->n == 0) return 0;
  B200_CHECK(d_b && d_x && d_b != d_x, B200_ERR_ARG_WRONG, "b and x must be distinct non-null vectors");
  B200_CHECK(p->epoch < 2147483000, B200_ERR_SUP, "epoch counter exhausted");
  if (ilu_set_backoff()) return B200_ERR_GPU;
  {
    int coop = 0;
    B200_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (g_ilu_march && coop == 1) return march_launch(h, p, d_b, d_x);
    if (g_ilu_packed && p->packed && coop == 1) {
      switch (p->G) {
      case 2: return packed_launch<2>(h, p, d_b, d_x);
      case 4: return packed_launch<4>(h, p, d_b, d_x);
      case 8: return packed_launch<8>(h, p, d_b, d_x);
      case 16: return packed_launch<16>(h, p, d_b, d_x);
      default: return packed_launch<32>(h, p, d_b, d_x);
      }
    }
  }
  switch (p->G) {
  case 2: return sweeps_launch<2>(h, p, d_b, d_x);
  case 4: return sweeps_launch<4>(h, p, d_b, d_x);
  case 8: return sweeps_launch<8>(h, p, d_b, d_x);
  case 16: return sweeps_launch<16>(h, p, d_b, d_x);
  default: return sweeps_launch<32>(h, p, d_b, d_x);
  }
}

extern "C" int b200Ilu0GetFactor(b200Handle h, b200IluPlan p, int *bi, int *bj, int *bdiag, double *ba)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (bi) memcpy(bi, p->h_bi, sizeof(int) * (size_t)(p->n + 1));
  if (bdiag) memcpy(bdiag, p->h_bdiag, sizeof(int) * (size_t)(p->n + 1));
  if (bj) memcpy(bj, p->h_bj, sizeof(int) * (size_t)p->nnz);
  if (ba && p->nnz) {
    B200_CUDA(cudaMemcpyAsync(ba, p->d_ba, sizeof(double) * (size_t)p->nnz, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

extern "C" int b200Ilu0GetInfo(b200IluPlan p, int *nlevL, int *nlevU, int64_t *nnz)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  if (nlevL) *nlevL = p->nlevL;
  if (nlevU) *nlevU = p->nlevU;
  if (nnz) *nnz = p->nnz;
  return 0;
}

/* the marching sweeps' schedule: lanes per row, segment slots (padded) and segment dependency levels of both sweeps */
extern "C" int b200Ilu0GetSegmentInfo(b200IluPlan p, int *lanes, int *nslotL, int *nlevL, int *nslotU, int *nlevU)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  if (lanes) *lanes = p->GS;
  if (nslotL) *nslotL = p->nsegslotL;
  if (nlevL) *nlevL = p->nseglevL;
  if (nslotU) *nslotU = p->nsegslotU;
  if (nlevU) *nlevU = p->nseglevU;
  return 0;
}
