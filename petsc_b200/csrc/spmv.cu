/*
 * spmv.cu -- CSR SpMV for MATSEQAIJ on sm_100a (HBM-bound gather; no tensor cores).
 *
 * Replaces the cusparseSpMV(CSR_ALG1) call site of the reference (aijcusparse.cu:2424-2568) and restates
 * MatMult_SeqAIJ / MatMultAdd_SeqAIJ (aij.c:1444-1499, 1606-1655) for the GPU.
 *
 * Kernel design (row-binned, TMA-staged):
 *   - The matrix is cut into row tiles of R consecutive rows.  R is the largest power of two for which every tile's
 *     nonzero count fits the shared-memory stage (b200CsrPlanCreate measures this on the device).
 *   - A persistent CTA walks its tiles through an S-stage shared-memory ring.  For every tile one elected thread issues
 *     three 1-D TMA bulk copies (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes): the row-pointer
 *     slice, the column-index slice and the value slice, all 16-byte granular, completing on one mbarrier.  val/col are
 *     therefore read from HBM exactly once, fully coalesced, with no register staging, and S-1 tiles are always in
 *     flight per CTA while one is being consumed.
 *   - Consumers: G lanes per row (G = 1,2,4,...,32 chosen from the mean row length).  Each lane walks its share of the
 *     row in shared memory, gathers x through the read-only path (x is the only operand with reuse: it lives in L2/L1),
 *     and the G partial sums are combined with warp shuffles.
 *   - G == 1 is the PARITY mode: one thread sums a row strictly left to right with __dmul_rn/__dadd_rn, i.e. exactly
 *     the arithmetic of the generic PetscSparseDensePlusDot branch (aij.h:609-614) in the reference's -O2 x86-64 build.
 *     y is then bit-identical to MatMult_SeqAIJ.  It is also the fastest mode for stencil matrices (<= ~12 nnz/row).
 *   - Epilogues fused into the same kernel: y = A x; z = y + A x (sum starts at y[r], as aij.c:1639-1652 does);
 *     w = dinv .* (A x) (PCApply_Jacobi fused, saves one n-vector write and read per Krylov iteration).
 *   - Tiles whose nonzeros do not fit a stage (a few very long rows) take an in-kernel fallback that streams the rows
 *     straight from global memory with the same G-lane decomposition.
 *
 * Algorithmic bytes per call: nnz*(8+4) + m*(4+8+8).
 */
#include "b200_internal.h"
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>

#define SPMV_TPB 256
#define SPMV_MAX_STAGES 4

struct b200CsrPlan_s {
  int        m, n;
  int64_t    nnz;
  const int *d_rowptr, *d_colidx;
  int        lanes;       /* G */
  int        rows_tile;   /* R */
  int        cap;         /* nnz capacity of one stage */
  int        stages;
  int        ctas_per_sm;
  int        grid;
  int        smem;
  int        max_row_nnz;
  int        num_sms;
  int        user_lanes, user_rows, user_stages, user_ctas;
  int        max_tile_nnz[16]; /* for R = 8 << k */
  int        hints;            /* bit0: CSR streams evict_first, bit1: x evict_last, bit2: persisting L2 window on x */
  int        l2_persist_max;   /* of the device the plan was made on */
  int        tree_sum;         /* 1 (default): FMA + shuffle tree for rows with several lanes (rounding differs from the reference in
                                  the last bits; one lane per row is always exact); 0: the reference's left-to-right FMA-free
                                  order for every lane count */
  int        hints_auto;
  double     span_bytes;       /* mean bytes of x spanned by one row ((last col - first col) * 8): how scattered the gather is */
  int        vec_lanes;        /* > 0: rows are handled by the streaming CSR-vector kernel with this many lanes per row */
  int       *d_longrows;       /* rows longer than SPMV_LONG_ROW (power-law tails), handled by the long-row bin: */
  int        nlong;
  int4      *d_longseg;        /* their entries cut into segments of SPMV_LONG_ROW: (row slot q, k0, k1, index of the row's first segment) */
  int       *d_longfirst;      /* [nlong+1] first segment of every long row */
  double    *d_longpart;       /* one partial sum per segment */
  int        nlongseg;
  struct CsrBlocks *blk;       /* column-blocked copy (b200CsrPlanSetColumnBlocks) or NULL */
};

/* Column-blocked layout for matrices whose gathered vector does not stay in the L2 (random CSR, SURVEY config 5): the columns
   are cut into nb ranges, every range is a CSR matrix of its own over ALL rows (entries of a row inside a range are
   contiguous because rows are sorted), and y = A x becomes nb passes y += A_b x that each touch only 8 n / nb bytes of x.
   The passes run in column order and each pass continues the row sum where the previous one stopped, so with an exact
   per-pass summation the result is still the reference's left-to-right sum.  Values are kept packed in block order
   (b200CsrPlanPackValues after every value change). */
#define SPMV_MAX_BLOCKS 64
struct CsrBlocks {
  int         nb, ms;                    /* blocks; row-pointer stride (m+1 rounded up to 4) */
  int        *d_rowptr;                  /* nb * ms */
  int        *d_col, *d_perm;            /* blocked columns; original position of every blocked entry */
  double     *d_val;                     /* packed values */
  int64_t     off[SPMV_MAX_BLOCKS + 1];  /* element offset of every block (multiples of 4) */
  int64_t     tot;                       /* allocated elements */
  b200CsrPlan sub[SPMV_MAX_BLOCKS];
  int         packed;
};

/* ------------------------------------------------------------------ PTX helpers: mbarrier + 1-D TMA bulk copy */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void     mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void     mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_LOOP:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra WAIT_DONE;\n"
    "bra WAIT_LOOP;\n"
    "WAIT_DONE:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

/* L2 eviction-priority hints: the CSR streams (val/col/rowptr) are read exactly once -> evict_first; x is the only
   operand with reuse (every entry is gathered ~nnz/row times within a window of a few grid planes) -> evict_last. */
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ double ldg_hint(const double *p, uint64_t pol)
{
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}

/* ------------------------------------------------------------------ shared-memory stage layout */
struct StageLayout {
  int rp_off, col_off, val_off, stage_bytes;
};
__host__ __device__ inline StageLayout stage_layout(int R, int cap)
{
  StageLayout L;
  L.val_off     = 0;                                   /* (cap+4) doubles */
  L.col_off     = (cap + 4) * 8;                       /* (cap+4) ints    */
  L.rp_off      = L.col_off + (cap + 4) * 4;           /* (R+4) ints      */
  L.stage_bytes = (L.rp_off + (R + 4) * 4 + 127) & ~127;
  return L;
}

/* ------------------------------------------------------------------ the kernel */
/* ORD (G > 1 only): the G lanes of a row still load and multiply G consecutive entries at a time (coalesced), but the
   products are then added into the row sum one after the other in column order (G width-G shuffles per chunk, every lane of
   the group keeps the same running sum) -- MatMult_SeqAIJ's left-to-right, FMA-free association at full bandwidth.  ORD = 0
   is the FMA + shuffle-tree variant (rounding differs from the reference in the last bits). */
template <int G, int HINTS, int ORD>
__global__ void __launch_bounds__(SPMV_TPB) csr_spmv_tile_kernel(int m, int R, int cap, int S, const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val, const double *__restrict__ x, const double *yin, const double *__restrict__ dinv, double *yout, double *yplain)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t                           full_bar[SPMV_MAX_STAGES];
  const StageLayout                             L      = stage_layout(R, cap);
  const int                                     ntiles = (m + R - 1) / R;
  const int                                     tid    = threadIdx.x;
  const uint64_t pol_stream = (HINTS & 1) ? policy_evict_first() : 0;
  const uint64_t pol_x      = (HINTS & 2) ? policy_evict_last() : 0;
  auto tma = [&](void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if (HINTS & 1) tma_load_1d_hint(dst, src, bytes, bar, pol_stream);
    else tma_load_1d(dst, src, bytes, bar);
  };
  auto ldx = [&](const double *p) -> double { return (HINTS & 2) ? ldg_hint(p, pol_x) : __ldg(p); };

  if (tid == 0) {
    for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  /* producer: issue the three bulk copies of tile t into stage s (or an empty arrival for an oversize tile) */
  auto issue = [&](int t, int s, int k0, int k1) {
    unsigned char *st   = smem + (size_t)s * L.stage_bytes;
    const int      r0   = t * R;
    const int      rows = min(R, m - r0);
    const uint32_t rpcnt = (uint32_t)((rows + 1 + 3) & ~3);
    if (k1 - k0 > cap) { /* fallback tile: only the row pointers are staged, the rows stream from global memory */
      mbar_expect_tx(&full_bar[s], rpcnt * 4u);
      tma(st + L.rp_off, rowptr + r0, rpcnt * 4u, &full_bar[s]);
      return;
    }
    const int      k0a   = k0 & ~3;                            /* 16-byte alignment for both arrays */
    const int64_t  kend  = ((int64_t)k1 + 3) & ~(int64_t)3;    /* <= roundup4(nnz): inside the allocation pad */
    const uint32_t cnt   = (uint32_t)(kend - k0a);
    const uint32_t bytes = cnt * 12u + rpcnt * 4u;
    mbar_expect_tx(&full_bar[s], bytes);
    tma(st + L.rp_off, rowptr + r0, rpcnt * 4u, &full_bar[s]);
    if (cnt) {
      tma(st + L.col_off, colidx + k0a, cnt * 4u, &full_bar[s]);
      tma(st + L.val_off, val + k0a, cnt * 8u, &full_bar[s]);
    }
  };

  /* prologue: fill the ring */
  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) {
        int r0 = t * R, r1 = min(r0 + R, m);
        issue(t, s, __ldg(rowptr + r0), __ldg(rowptr + r1));
      }
    }
  }

  constexpr int RPP = SPMV_TPB / G; /* rows per pass */
  const int     grp = tid / G, gl = tid % G;

  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
    const int      s      = it % S;
    const uint32_t parity = (uint32_t)((it / S) & 1);
    /* bounds of the tile that will reuse this stage: loaded early so the latency hides behind the compute */
    int       nk0 = 0, nk1 = 0;
    const int tn = t + S * gridDim.x;
    if (tid == 0 && tn < ntiles) {
      int r0 = tn * R, r1 = min(r0 + R, m);
      nk0 = __ldg(rowptr + r0);
      nk1 = __ldg(rowptr + r1);
    }
    mbar_wait(&full_bar[s], parity);

    unsigned char *st   = smem + (size_t)s * L.stage_bytes;
    const int     *rp   = reinterpret_cast<const int *>(st + L.rp_off);
    const int     *cs   = reinterpret_cast<const int *>(st + L.col_off);
    const double  *vs   = reinterpret_cast<const double *>(st + L.val_off);
    const int      r0   = t * R;
    const int      rows = min(R, m - r0);
    const int      k0   = rp[0];
    const int      k1   = rp[rows];
    const bool     staged = (k1 - k0 <= cap);
    const int      k0a  = k0 & ~3;

    /* the row loop bound is warp-uniform (lanes without a row stay in the loop and only join the shuffles) */
    for (int base = 0; base < rows; base += RPP) {
      const int  lr  = base + grp;
      const bool act = lr < rows;
      const int  r   = r0 + lr;
      const int  ks  = act ? rp[lr] : 0, ke = act ? rp[lr + 1] : 0;
      double     sum;
      if (G == 1) {
        /* parity mode: strict left-to-right, FMA-free (aij.h:609-614) */
        sum = (act && yin) ? yin[r] : 0.0;
        if (staged) {
          /* four gathers in flight per thread, then the same left-to-right adds (order unchanged => still bit-exact) */
          int k = ks - k0a;
          for (; k + 4 <= ke - k0a; k += 4) {
            const double x0 = ldx(x + cs[k]), x1 = ldx(x + cs[k + 1]), x2 = ldx(x + cs[k + 2]), x3 = ldx(x + cs[k + 3]);
            sum = __dadd_rn(sum, __dmul_rn(vs[k], x0));
            sum = __dadd_rn(sum, __dmul_rn(vs[k + 1], x1));
            sum = __dadd_rn(sum, __dmul_rn(vs[k + 2], x2));
            sum = __dadd_rn(sum, __dmul_rn(vs[k + 3], x3));
          }
          for (; k < ke - k0a; k++) sum = __dadd_rn(sum, __dmul_rn(vs[k], ldx(x + cs[k])));
        } else {
          for (int k = ks; k < ke; k++) sum = __dadd_rn(sum, __dmul_rn(__ldg(val + k), __ldg(x + __ldg(colidx + k))));
        }
      } else if (ORD) {
        sum = (act && yin) ? yin[r] : 0.0;
        const int len = ke - ks; /* 0 for lanes without a row */
        int       nch = (len + G - 1) / G;
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) nch = max(nch, __shfl_xor_sync(0xffffffffu, nch, o)); /* warp-uniform trip count */
        const int kb = staged ? ks - k0a : ks;
        for (int c = 0; c < nch; c += 4) { /* four chunks of G entries: four independent gathers in flight per lane */
          double p[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int  kk = (c + u) * G + gl;
            const bool v  = kk < len;
            if (staged) p[u] = v ? __dmul_rn(vs[kb + kk], ldx(x + cs[kb + kk])) : 0.0;
            else p[u] = v ? __dmul_rn(__ldg(val + kb + kk), __ldg(x + __ldg(colidx + kb + kk))) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int cnt = len - (c + u) * G; /* entries of this chunk that exist (<= 0: none) */
#pragma unroll
            for (int l = 0; l < G; l++) {
              const double pl = __shfl_sync(0xffffffffu, p[u], l, G);
              if (l < cnt) sum = __dadd_rn(sum, pl);
            }
          }
        }
      } else {
        sum = 0.0;
        if (staged) {
          int k = ks - k0a + gl;
          for (; k + 3 * G < ke - k0a; k += 4 * G) { /* four independent gathers in flight per lane */
            const double x0 = ldx(x + cs[k]), x1 = ldx(x + cs[k + G]), x2 = ldx(x + cs[k + 2 * G]), x3 = ldx(x + cs[k + 3 * G]);
            sum = fma(vs[k], x0, sum);
            sum = fma(vs[k + G], x1, sum);
            sum = fma(vs[k + 2 * G], x2, sum);
            sum = fma(vs[k + 3 * G], x3, sum);
          }
          for (; k < ke - k0a; k += G) sum = fma(vs[k], ldx(x + cs[k]), sum);
        } else {
          for (int k = ks + gl; k < ke; k += G) sum = fma(__ldg(val + k), __ldg(x + __ldg(colidx + k)), sum);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (act && yin) sum += yin[r];
      }
      if (act && (G == 1 || gl == 0)) {
        if (yplain) yplain[r] = sum;
        yout[r] = dinv ? __dmul_rn(sum, __ldg(dinv + r)) : sum;
      }
    }
    __syncthreads(); /* every thread is done reading stage s */
    if (tid == 0 && tn < ntiles) issue(tn, s, nk0, nk1);
  }
}

/* ------------------------------------------------------------------ analysis kernels */
__global__ void tile_nnz_max_kernel(int m, const int *__restrict__ rowptr, int *out /* [16] */, int *maxrow)
{
  /* for R = 8<<k : max over tiles of rowptr[min((t+1)R,m)] - rowptr[tR]; also the max row length */
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int           loc[16];
  int           mr = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) loc[k] = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    int a = rowptr[r];
    mr    = max(mr, rowptr[r + 1] - a);
    if ((r & 7) == 0) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int64_t R = (int64_t)8 << k;
        if ((r & (R - 1)) == 0) {
          int64_t e = r + R;
          if (e > m) e = m;
          loc[k] = max(loc[k], rowptr[e] - a);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; k++) {
    int v = loc[k];
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(&out[k], v);
  }
  for (int o = 16; o > 0; o >>= 1) mr = max(mr, __shfl_xor_sync(0xffffffffu, mr, o));
  if ((threadIdx.x & 31) == 0 && mr) atomicMax(maxrow, mr);
}

__global__ void get_diagonal_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val, double *diag, int *diagpos)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    int    pos = -1;
    double d   = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      int c = colidx[k];
      if (c == (int)r) {
        pos = k;
        d   = val[k];
        break;
      }
      if (c > (int)r) break; /* columns are sorted (MatAssemblyEnd_SeqAIJ) */
    }
    diag[r] = d;
    if (diagpos) diagpos[r] = pos;
  }
}

__global__ void jacobi_invert_kernel(int64_t n, const double *__restrict__ d, double *dinv, int *nzero)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int           z      = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = d[i];
    if (v == 0.0) {
      z++;
      dinv[i] = 1.0; /* jacobi.c:253-266 */
    } else dinv[i] = 1.0 / v;
  }
  if (z) atomicAdd(nzero, z);
}

/* ------------------------------------------------------------------ streaming CSR-vector kernel (scattered / very long rows)
   The tile kernel above stages whole row tiles in shared memory with TMA and then gathers x: ideal when the gather window is
   cache-resident (stencils: 1.09 / 0.98 of the measured HBM peak), but on matrices whose columns are SCATTERED over a vector far
   larger than the L2 every gather is a DRAM access (~1 us), and a single-stage tile keeps too few of them in flight (random CSR
   d = 32: 65 G gathers/s, cuSPARSE CSR_ALG1 reaches 140 G/s); rows too long for a shared-memory stage (d = 512) fell back to a
   slow in-kernel path.  This kernel is the classic CSR-vector scheme tuned for latency hiding: G lanes per row read val/col
   coalesced straight from global memory (streaming, evict-first), four independent gathers per lane in flight, FMA
   accumulation, shuffle reduction.  The row sum is therefore NOT the reference's left-to-right order (same contract as the
   tile kernel with several lanes per row: equal to MatMult_SeqAIJ to rounding, <= 1e-12); -mat_b200_spmv_ordered /
   b200CsrPlanSetSummation(plan, 0) and one-lane plans keep using the tile kernel.  Epilogues as in the tile kernel. */
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int *p)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
#define SPMV_LONG_ROW 8192 /* entries: beyond this a row gets a whole CTA (row-binning for power-law matrices: G is per bin, not per matrix) */
template <int G>
__global__ void __launch_bounds__(256) csr_spmv_vector_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val, const double *__restrict__ x,
                                                              const double *__restrict__ yin, const double *__restrict__ dinv, double *__restrict__ yout, double *__restrict__ yplain, int skip_long)
{
  const int      lane = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int64_t  stride = ((int64_t)gridDim.x * blockDim.x) / G;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < m; r += stride) {
    const int k0 = rowptr[r], k1 = rowptr[r + 1];
    if (skip_long && k1 - k0 > SPMV_LONG_ROW) continue; /* handled by csr_spmv_longrow_kernel */
    double    s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int       k = k0 + lane;
    for (; k + 3 * G < k1; k += 4 * G) { /* four independent gathers in flight per lane */
      const int    c0 = ld_stream_s32(colidx + k), c1 = ld_stream_s32(colidx + k + G), c2 = ld_stream_s32(colidx + k + 2 * G), c3 = ld_stream_s32(colidx + k + 3 * G);
      const double a0 = ld_stream_f64(val + k), a1 = ld_stream_f64(val + k + G), a2 = ld_stream_f64(val + k + 2 * G), a3 = ld_stream_f64(val + k + 3 * G);
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      s0 = fma(a0, x0, s0); s1 = fma(a1, x1, s1); s2 = fma(a2, x2, s2); s3 = fma(a3, x3, s3);
    }
    for (; k < k1; k += G) s0 = fma(ld_stream_f64(val + k), __ldg(x + ld_stream_s32(colidx + k)), s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o, G);
    if (lane == 0) {
      if (yin) s += yin[r];
      if (yplain) yplain[r] = s;
      yout[r] = dinv ? s * dinv[r] : s;
    }
  }
}
/* the long-row bin: the entries of a long row are cut into segments of SPMV_LONG_ROW; one CTA per SEGMENT (256 threads x 4
   independent gathers, block reduction in fixed order) leaves a partial sum, a second small kernel adds the partials of a row in
   segment order and applies the epilogue -- deterministic, and a 600 k-entry row is spread over 74 CTAs instead of one */
__global__ void __launch_bounds__(256) csr_spmv_longseg_kernel(int nseg, const int4 *__restrict__ seg, const int *__restrict__ colidx, const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ part)
{
  __shared__ double red[8];
  for (int q = blockIdx.x; q < nseg; q += gridDim.x) {
    const int4 sg = seg[q];
    double     s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int        k = sg.y + threadIdx.x;
    for (; k + 3 * 256 < sg.z; k += 4 * 256) {
      const int    c0 = ld_stream_s32(colidx + k), c1 = ld_stream_s32(colidx + k + 256), c2 = ld_stream_s32(colidx + k + 512), c3 = ld_stream_s32(colidx + k + 768);
      const double a0 = ld_stream_f64(val + k), a1 = ld_stream_f64(val + k + 256), a2 = ld_stream_f64(val + k + 512), a3 = ld_stream_f64(val + k + 768);
      s0 = fma(a0, __ldg(x + c0), s0); s1 = fma(a1, __ldg(x + c1), s1); s2 = fma(a2, __ldg(x + c2), s2); s3 = fma(a3, __ldg(x + c3), s3);
    }
    for (; k < sg.z; k += 256) s0 = fma(ld_stream_f64(val + k), __ldg(x + ld_stream_s32(colidx + k)), s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) part[q] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    __syncthreads();
  }
}
__global__ void csr_spmv_longfinal_kernel(int nlong, const int *__restrict__ rows, const int *__restrict__ first, const double *__restrict__ part, const double *__restrict__ yin, const double *__restrict__ dinv,
                                          double *__restrict__ yout, double *__restrict__ yplain)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nlong) return;
  const int r = rows[q];
  double    t = 0.0;
  for (int k = first[q]; k < first[q + 1]; k++) t += part[k]; /* segment order */
  if (yin) t += yin[r];
  if (yplain) yplain[r] = t;
  yout[r] = dinv ? t * dinv[r] : t;
}
__global__ void csr_longrow_ext_kernel(int nlong, const int *__restrict__ rows, const int *__restrict__ rowptr, int2 *ext)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nlong) ext[q] = make_int2(rowptr[rows[q]], rowptr[rows[q] + 1]);
}
__global__ void csr_longrow_list_kernel(int m, const int *__restrict__ rowptr, int *list, int *count)
{
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride)
    if (rowptr[r + 1] - rowptr[r] > SPMV_LONG_ROW) list[atomicAdd(count, 1)] = r; /* order is irrelevant: every row is computed on its own */
}
template <int G>
static int spmv_vector_launch(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  int64_t g = ((int64_t)p->m * G + 255) / 256;
  if (g > (int64_t)h->num_sms * 32) g = (int64_t)h->num_sms * 32;
  csr_spmv_vector_kernel<G><<<(int)g, 256, 0, h->stream>>>(p->m, p->d_rowptr, p->d_colidx, val, x, yin, dinv, yout, yplain, p->nlong > 0);
  B200_LAUNCHED(1);
  if (p->nlong > 0) {
    int gl = p->nlongseg < h->num_sms * 8 ? p->nlongseg : h->num_sms * 8;
    csr_spmv_longseg_kernel<<<gl, 256, 0, h->stream>>>(p->nlongseg, p->d_longseg, p->d_colidx, val, x, p->d_longpart);
    csr_spmv_longfinal_kernel<<<(p->nlong + 127) / 128, 128, 0, h->stream>>>(p->nlong, p->d_longrows, p->d_longfirst, p->d_longpart, yin, dinv, yout, yplain);
    B200_LAUNCHED(2);
  }
  B200_KERNEL_CHECK();
  return 0;
}
__global__ void csr_span_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, double *sum);

/* ------------------------------------------------------------------ plan */
static int pick_lanes(double avg)
{
  if (avg <= 12.0) return 1;
  if (avg <= 24.0) return 2;
  if (avg <= 48.0) return 4;
  if (avg <= 96.0) return 8;
  if (avg <= 192.0) return 16;
  return 32;
}

static int roundup4(int v) { return (v + 3) & ~3; }

static int plan_configure(b200CsrPlan p)
{
  double avg = p->m ? (double)p->nnz / p->m : 0.0;
  int    G   = p->user_lanes ? p->user_lanes : pick_lanes(avg);
  int    S   = p->user_stages ? p->user_stages : 1;
  int    cps = p->user_ctas ? p->user_ctas : 8; /* measured on B200 (profiles/round1_notes.md): 8 resident single-stage CTAs (1.95 ms) beat 6 (2.02), 4 x 2 stages (2.17) and 2 x 2 (2.79 ms): occupancy hides the TMA latency better than a deeper ring */
  if (S > SPMV_MAX_STAGES) S = SPMV_MAX_STAGES;
  /* shared-memory budget per CTA: 227 KB per SM, 1 KB reserved per resident CTA, static smem for the barriers */
  const int budget = (227 * 1024) / cps - 1024 - 256;
  int       minR   = SPMV_TPB / G; /* at least one full pass */
  if (minR < 8) minR = 8;
  int R = 0, cap = 0;
  for (int k = 8; k >= 0 && !R; k--) { /* largest R = 8<<k whose worst tile fits one stage */
    int Rk = 8 << k;
    if (p->user_rows ? Rk != p->user_rows : Rk < minR) continue;
    int capk = roundup4(p->max_tile_nnz[k]);
    if (capk < 64) capk = 64;
    if ((int64_t)stage_layout(Rk, capk).stage_bytes * S <= budget || p->user_rows) {
      R   = Rk;
      cap = capk;
    }
  }
  if (!R) { /* user asked for an unsupported R, or even one pass of rows overflows: smallest R, fallback tiles */
    R   = minR;
    cap = 1 << 30;
  }
  /* clamp the capacity to what fits; tiles above it take the in-kernel global-memory path */
  {
    int fit = ((budget / S - (R + 4) * 4 - 128) / 12 - 4) & ~3;
    if (fit < 64) fit = 64;
    if (cap > fit) cap = fit;
  }
  StageLayout L  = stage_layout(R, cap);
  p->lanes       = G;
  { /* streaming CSR-vector kernel for scattered gathers (mean row span beyond ~16 MB of x) and for rows too long for a stage */
    const char *e = getenv("PETSCB200_SPMV_VECTOR"); /* experiments: lanes per row, 0 = never */
    int         v = 0;
    if (p->span_bytes > (double)(16 << 20) && avg >= 2.0) v = avg < 6.0 ? 2 : (avg < 12.0 ? 4 : (avg < 24.0 ? 8 : (avg < 64.0 ? 16 : 32)));
    if (avg > 192.0) v = 32;
    if (p->nlong > 0 && !v) v = avg < 6.0 ? 2 : (avg < 12.0 ? 4 : (avg < 24.0 ? 8 : (avg < 64.0 ? 16 : 32))); /* power-law tail: binned kernels instead of the tile kernel's in-kernel fallback */
    if (e) v = atoi(e);
    if (p->user_lanes == 1) v = 0; /* the caller asked for the parity layout */
    p->vec_lanes = (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
  }
  /* L2 hints (measured, profiles/round1_notes.md): stencil-like matrices (G = 1): x evict_last only (2); matrices with
     scattered columns: also stream val/col as evict_first (3) so the gathered x keeps more of the L2 */
  if (p->hints < 0 || p->hints_auto) {
    p->hints      = (G == 1) ? 2 : 3;
    /* scattered columns and an x that fits the L2 set-aside: pin x with a persisting access-policy window (bit 2) */
    if (G > 1 && p->l2_persist_max > 0 && (int64_t)p->n * 8 >= (4 << 20) && (int64_t)p->n * 8 <= (int64_t)p->l2_persist_max) p->hints |= 4;
    p->hints_auto = 1;
  }
  p->rows_tile   = R;
  p->cap         = cap;
  p->stages      = S;
  p->ctas_per_sm = cps;
  p->smem        = L.stage_bytes * S;
  int ntiles     = (p->m + R - 1) / R;
  int grid       = p->num_sms * cps;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  p->grid = grid;
  return 0;
}

extern "C" int b200CsrPlanCreate(b200Handle h, int m, int n, int64_t nnz, const int *d_rowptr, const int *d_colidx, b200CsrPlan *plan)
{
  B200_CHECK(h && plan, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(m >= 0 && n >= 0 && nnz >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  B200_CHECK(nnz <= 2147483647LL - 8, B200_ERR_SUP, "nnz %lld does not fit 32-bit PetscInt row pointers (needs --with-64-bit-indices)", (long long)nnz);
  B200_CHECK(!m || (d_rowptr && (nnz == 0 || d_colidx)), B200_ERR_ARG_NULL, "null CSR arrays");
  b200CsrPlan p = (b200CsrPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->m = m; p->n = n; p->nnz = nnz; p->d_rowptr = d_rowptr; p->d_colidx = d_colidx; p->num_sms = h->num_sms; p->l2_persist_max = h->l2_persist_max;
  p->hints = -1; /* auto, see plan_configure */
  /* measured (profiles/round1_notes.md): the ordered sums cost 1.0x (d <= 128 random), 1.4x (d = 512) and 1.7x (27-point) of the
     tree variant, so the fast variant stays the default and the exact one is a switch (b200CsrPlanSetSummation(plan, 0)) */
  p->tree_sum = 1;
  if (m > 0) {
    int *d_stats;
    int  hstats[17];
    B200_CUDA(cudaMalloc(&d_stats, sizeof(int) * 17));
    B200_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(int) * 17, h->stream));
    int g = (int)(((int64_t)m + 255) / 256);
    if (g > h->num_sms * 8) g = h->num_sms * 8;
    tile_nnz_max_kernel<<<g, 256, 0, h->stream>>>(m, d_rowptr, d_stats, d_stats + 16);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
    B200_CUDA(cudaMemcpyAsync(hstats, d_stats, sizeof hstats, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_CUDA(cudaFree(d_stats));
    for (int k = 0; k < 16; k++) p->max_tile_nnz[k] = hstats[k];
    p->max_row_nnz = hstats[16];
    if (nnz > 0 && (int64_t)n * 8 > (16 << 20)) { /* a gather can only be scattered if x is large */
      double *d_sum = NULL, span = 0.0;
      B200_CUDA(cudaMalloc(&d_sum, sizeof(double)));
      B200_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), h->stream));
      csr_span_kernel<<<g, 256, 0, h->stream>>>(m, d_rowptr, d_colidx, d_sum);
      B200_LAUNCHED(1);
      B200_CUDA(cudaMemcpyAsync(&span, d_sum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      cudaFree(d_sum);
      p->span_bytes = span / m * 8.0;
    }
    if (p->max_row_nnz > SPMV_LONG_ROW) { /* row-binning: the tail of very long rows gets its own kernel */
      int *d_cnt = NULL, cnt = 0;
      B200_CUDA(cudaMalloc(&d_cnt, sizeof(int)));
      B200_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
      const int64_t cap = nnz / SPMV_LONG_ROW + 1; /* there cannot be more rows that long */
      B200_CUDA(cudaMalloc(&p->d_longrows, sizeof(int) * (size_t)cap));
      csr_longrow_list_kernel<<<g, 256, 0, h->stream>>>(m, d_rowptr, p->d_longrows, d_cnt);
      B200_LAUNCHED(1);
      B200_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      cudaFree(d_cnt);
      p->nlong = cnt;
      if (cnt > 0) { /* cut the long rows into segments (host: a handful of rows) */
        int2 *d_ext = NULL, *ext = (int2 *)malloc(sizeof(int2) * (size_t)cnt);
        B200_CUDA(cudaMalloc(&d_ext, sizeof(int2) * (size_t)cnt));
        csr_longrow_ext_kernel<<<(cnt + 127) / 128, 128, 0, h->stream>>>(cnt, p->d_longrows, d_rowptr, d_ext);
        B200_LAUNCHED(1);
        B200_CUDA(cudaMemcpyAsync(ext, d_ext, sizeof(int2) * (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(d_ext);
        int64_t nseg = 0;
        for (int q = 0; q < cnt; q++) nseg += ((int64_t)(ext[q].y - ext[q].x) + SPMV_LONG_ROW - 1) / SPMV_LONG_ROW;
        int4 *seg   = (int4 *)malloc(sizeof(int4) * (size_t)(nseg + 1));
        int  *first = (int *)malloc(sizeof(int) * (size_t)(cnt + 1));
        int   w     = 0;
        for (int q = 0; q < cnt; q++) {
          first[q] = w;
          for (int k0 = ext[q].x; k0 < ext[q].y; k0 += SPMV_LONG_ROW) seg[w++] = make_int4(q, k0, k0 + SPMV_LONG_ROW < ext[q].y ? k0 + SPMV_LONG_ROW : ext[q].y, first[q]);
        }
        first[cnt]  = w;
        p->nlongseg = w;
        B200_CUDA(cudaMalloc(&p->d_longseg, sizeof(int4) * (size_t)(w + 1)));
        B200_CUDA(cudaMalloc(&p->d_longfirst, sizeof(int) * (size_t)(cnt + 1)));
        B200_CUDA(cudaMalloc(&p->d_longpart, sizeof(double) * (size_t)(w + 1)));
        B200_CUDA(cudaMemcpyAsync(p->d_longseg, seg, sizeof(int4) * (size_t)w, cudaMemcpyHostToDevice, h->stream));
        B200_CUDA(cudaMemcpyAsync(p->d_longfirst, first, sizeof(int) * (size_t)(cnt + 1), cudaMemcpyHostToDevice, h->stream));
        B200_CUDA(cudaStreamSynchronize(h->stream));
        free(ext); free(seg); free(first);
      }
    }
  }
  plan_configure(p);
  *plan = p;
  return 0;
}

static void csr_blocks_free(struct CsrBlocks *B)
{
  if (!B) return;
  for (int b = 0; b < B->nb; b++)
    if (B->sub[b]) b200CsrPlanDestroy(B->sub[b]);
  cudaFree(B->d_rowptr);
  cudaFree(B->d_col);
  cudaFree(B->d_perm);
  cudaFree(B->d_val);
  free(B);
}

extern "C" int b200CsrPlanDestroy(b200CsrPlan plan)
{
  if (plan) {
    csr_blocks_free(plan->blk);
    cudaFree(plan->d_longrows); cudaFree(plan->d_longseg); cudaFree(plan->d_longfirst); cudaFree(plan->d_longpart);
  }
  free(plan);
  return 0;
}

extern "C" int b200CsrPlanSetLayout(b200CsrPlan p, int lanes, int rows_per_tile, int stages, int ctas_per_sm)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  B200_CHECK(lanes == 0 || lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32, B200_ERR_ARG_OUTOFRANGE, "lanes_per_row must be 0,1,2,4,8,16,32");
  B200_CHECK(rows_per_tile == 0 || (rows_per_tile >= 8 && rows_per_tile <= 2048 && (rows_per_tile & (rows_per_tile - 1)) == 0), B200_ERR_ARG_OUTOFRANGE, "rows_per_tile must be 0 or a power of two in [8,2048]");
  B200_CHECK(stages >= 0 && stages <= SPMV_MAX_STAGES && ctas_per_sm >= 0 && ctas_per_sm <= 8, B200_ERR_ARG_OUTOFRANGE, "stages/ctas out of range");
  p->user_lanes = lanes; p->user_rows = rows_per_tile; p->user_stages = stages; p->user_ctas = ctas_per_sm;
  if (p->blk) /* the column blocks follow the lane choice; their tiles are sized from their own row lengths */
    for (int b = 0; b < p->blk->nb; b++) {
      int rc = b200CsrPlanSetLayout(p->blk->sub[b], lanes, 0, 0, 0);
      if (rc) return rc;
    }
  return plan_configure(p);
}

extern "C" int b200CsrPlanSetCacheHints(b200CsrPlan p, int hints)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  B200_CHECK(hints >= 0 && hints <= 7, B200_ERR_ARG_OUTOFRANGE, "hints must be 0..7");
  p->hints      = hints;
  p->hints_auto = 0;
  return 0;
}

extern "C" int b200CsrPlanSetSummation(b200CsrPlan p, int tree)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  B200_CHECK(tree == 0 || tree == 1, B200_ERR_ARG_OUTOFRANGE, "summation mode must be 0 (reference order) or 1 (tree)");
  p->tree_sum = tree;
  return 0;
}

/* ------------------------------------------------------------------ column-blocked layout (set-up kernels) */
__global__ void blk_count_kernel(int m, int ms, int nb, int cw, const int *__restrict__ rowptr, const int *__restrict__ colidx, int *brow)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < ms; r += stride) {
    int k = r < m ? rowptr[r] : 0;
    const int ke = r < m ? rowptr[r + 1] : 0;
    for (int b = 0; b < nb; b++) {
      const int64_t bound = b == nb - 1 ? (int64_t)1 << 40 : (int64_t)(b + 1) * cw;
      int           cnt   = 0;
      while (k < ke && colidx[k] < bound) { k++; cnt++; }
      brow[(int64_t)b * ms + r] = cnt; /* rows >= m count 0: the exclusive scan leaves the block total at index m */
    }
  }
}
__global__ void blk_fill_kernel(int m, int ms, int nb, const int *__restrict__ rowptr, const int *__restrict__ colidx, const int *__restrict__ brow, const long long *__restrict__ off, int *col, int *perm)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    int k = rowptr[r];
    for (int b = 0; b < nb; b++) {
      const int     s = brow[(int64_t)b * ms + r], e = brow[(int64_t)b * ms + r + 1];
      const int64_t base = off[b] + s;
      for (int t = 0; t < e - s; t++, k++) {
        col[base + t]  = colidx[k];
        perm[base + t] = k;
      }
    }
  }
}
__global__ void blk_pack_kernel(int64_t tot, const int *__restrict__ perm, const double *__restrict__ val, double *out)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += stride) out[k] = val[perm[k]];
}

/* nblocks <= 1 removes the blocked copy.  The plan keeps using the caller's rowptr/colidx for its own (unblocked) layout. */
extern "C" int b200CsrPlanSetColumnBlocks(b200Handle h, b200CsrPlan p, int nblocks)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(nblocks >= 0 && nblocks <= SPMV_MAX_BLOCKS, B200_ERR_ARG_OUTOFRANGE, "number of column blocks must be in [0,%d]", SPMV_MAX_BLOCKS);
  csr_blocks_free(p->blk);
  p->blk = NULL;
  if (nblocks <= 1 || p->m == 0 || p->nnz == 0) return 0;
  struct CsrBlocks *B = (struct CsrBlocks *)calloc(1, sizeof(*B));
  B200_CHECK(B, B200_ERR_MEM, "out of host memory");
  const int m = p->m, ms = (m + 1 + 3) & ~3, nb = nblocks, cw = (p->n + nb - 1) / nb;
  B->nb = nb; B->ms = ms;
  int        rc = 0;
  void      *tmp = NULL;
  long long *d_off = NULL;
  int       *tot = (int *)malloc(sizeof(int) * (size_t)nb);
#define BLK_CUDA(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      b200_set_error(B200_ERR_GPU, "cuda error %d (%s) : %s at %s:%d", (int)e_, cudaGetErrorName(e_), cudaGetErrorString(e_), __FILE__, __LINE__); \
      rc = B200_ERR_GPU; \
      goto done; \
    } \
  } while (0)
  {
    int64_t g = ((int64_t)ms + 255) / 256;
    if (g > (int64_t)h->num_sms * 16) g = (int64_t)h->num_sms * 16;
    BLK_CUDA(cudaMalloc(&B->d_rowptr, sizeof(int) * (size_t)nb * ms));
    blk_count_kernel<<<(int)g, 256, 0, h->stream>>>(m, ms, nb, cw, p->d_rowptr, p->d_colidx, B->d_rowptr);
    B200_LAUNCHED(1);
    BLK_CUDA(cudaPeekAtLastError());
    size_t tmp_bytes = 0;
    BLK_CUDA(cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, B->d_rowptr, B->d_rowptr, m + 1, h->stream));
    BLK_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    for (int b = 0; b < nb; b++) {
      BLK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, B->d_rowptr + (size_t)b * ms, B->d_rowptr + (size_t)b * ms, m + 1, h->stream));
      BLK_CUDA(cudaMemcpyAsync(&tot[b], B->d_rowptr + (size_t)b * ms + m, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    B200_LAUNCHED(nb);
    BLK_CUDA(cudaStreamSynchronize(h->stream));
    B->off[0] = 0;
    for (int b = 0; b < nb; b++) B->off[b + 1] = B->off[b] + (((int64_t)tot[b] + 3) & ~(int64_t)3) + 4; /* 16-byte aligned starts, slack for the kernel's 16-byte over-read */
    B->tot = B->off[nb];
    BLK_CUDA(cudaMalloc(&B->d_col, sizeof(int) * (size_t)B->tot + B200_ALLOC_PAD));
    BLK_CUDA(cudaMalloc(&B->d_perm, sizeof(int) * (size_t)B->tot + B200_ALLOC_PAD));
    BLK_CUDA(cudaMalloc(&B->d_val, sizeof(double) * (size_t)B->tot + B200_ALLOC_PAD));
    BLK_CUDA(cudaMemsetAsync(B->d_col, 0, sizeof(int) * (size_t)B->tot, h->stream));
    BLK_CUDA(cudaMemsetAsync(B->d_perm, 0, sizeof(int) * (size_t)B->tot, h->stream));
    BLK_CUDA(cudaMalloc(&d_off, sizeof(long long) * (size_t)(nb + 1)));
    {
      long long hoff[SPMV_MAX_BLOCKS + 1];
      for (int b = 0; b <= nb; b++) hoff[b] = (long long)B->off[b];
      BLK_CUDA(cudaMemcpyAsync(d_off, hoff, sizeof(long long) * (size_t)(nb + 1), cudaMemcpyHostToDevice, h->stream));
      blk_fill_kernel<<<(int)g, 256, 0, h->stream>>>(m, ms, nb, p->d_rowptr, p->d_colidx, B->d_rowptr, d_off, B->d_col, B->d_perm);
      B200_LAUNCHED(1);
      BLK_CUDA(cudaPeekAtLastError());
      BLK_CUDA(cudaStreamSynchronize(h->stream)); /* hoff is on this stack frame */
    }
    for (int b = 0; b < nb && !rc; b++) {
      rc = b200CsrPlanCreate(h, m, p->n, (int64_t)tot[b], B->d_rowptr + (size_t)b * ms, B->d_col + B->off[b], &B->sub[b]);
      if (!rc && p->user_lanes) rc = b200CsrPlanSetLayout(B->sub[b], p->user_lanes, 0, 0, 0);
    }
  }
done:
  cudaFree(tmp);
  cudaFree(d_off);
  free(tot);
  if (rc) {
    csr_blocks_free(B);
    return rc;
  }
  p->blk = B;
  return 0;
#undef BLK_CUDA
}

/* Automatic choice of the column blocks: blocking pays when the gathered vector does not stay in the L2 next
   to the streamed matrix (n*8 well above ~40 MB of the 126 MB L2), the rows are long enough to amortise the extra row-pointer /
   y traffic of every pass (>= 16 entries), and the columns of a row are SCATTERED (mean column span of a row, measured by a
   reduction kernel, beyond the L2 window) -- stencil-like matrices gather from a narrow window whatever n is.  Measured on
   random CSR n = 10 M: 2 blocks already bring d = 32 from 5.1 to 3.3 ms and d = 128 from 20.4 to 11.7 ms
   (profiles/round1_configs_5_blocks_pipecg.json); more blocks only add passes.  *nblocks_out = 0 when blocking is not used. */
__global__ void csr_span_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, double *sum)
{
  const int stride = gridDim.x * blockDim.x;
  double    s      = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    const int k0 = rowptr[r], k1 = rowptr[r + 1];
    if (k1 > k0) s += (double)(colidx[k1 - 1] - colidx[k0]); /* columns are sorted within a row */
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(sum, s);
}

extern "C" int b200CsrPlanAutoColumnBlocks(b200Handle h, b200CsrPlan p, int *nblocks_out)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (nblocks_out) *nblocks_out = 0;
  const double  avg     = p->m ? (double)p->nnz / p->m : 0.0;
  const int64_t xbytes  = (int64_t)p->n * 8, window = 40LL << 20;
  if (p->m == 0 || p->nnz == 0 || avg < 16.0 || xbytes <= (48LL << 20)) return b200CsrPlanSetColumnBlocks(h, p, 0);
  const double span = p->span_bytes; /* mean bytes of x spanned by one row, measured at plan creation */
  if (span <= (double)window) return b200CsrPlanSetColumnBlocks(h, p, 0);
  int nb = (int)((xbytes + window - 1) / window);
  if (nb < 2) nb = 2;
  if (nb > SPMV_MAX_BLOCKS) nb = SPMV_MAX_BLOCKS;
  int rc = b200CsrPlanSetColumnBlocks(h, p, nb);
  if (!rc && nblocks_out) *nblocks_out = nb;
  return rc;
}

/* gathers the values into block order; call after every change of the matrix values (the blocked SpMV uses this copy) */
extern "C" int b200CsrPlanPackValues(b200Handle h, b200CsrPlan p, const double *d_val)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->blk) return 0;
  B200_CHECK(d_val, B200_ERR_ARG_NULL, "null value array");
  int64_t g = (p->blk->tot + 255) / 256;
  if (g > (int64_t)h->num_sms * 16) g = (int64_t)h->num_sms * 16;
  blk_pack_kernel<<<(int)g, 256, 0, h->stream>>>(p->blk->tot, p->blk->d_perm, d_val, p->blk->d_val);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  p->blk->packed = 1;
  return 0;
}

extern "C" int b200CsrPlanGetLayout(b200CsrPlan p, int *lanes, int *rows, int *stages, int *grid, int *smem, int *maxrow)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  if (lanes) *lanes = p->lanes;
  if (rows) *rows = p->rows_tile;
  if (stages) *stages = p->stages;
  if (grid) *grid = p->grid;
  if (smem) *smem = p->smem;
  if (maxrow) *maxrow = p->max_row_nnz;
  return 0;
}

template <int G, int HINTS, int ORD>
static int spmv_launch_gho(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  static int configured = 0;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(csr_spmv_tile_kernel<G, HINTS, ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    configured = 1;
  }
  csr_spmv_tile_kernel<G, HINTS, ORD><<<p->grid, SPMV_TPB, p->smem, h->stream>>>(p->m, p->rows_tile, p->cap, p->stages, p->d_rowptr, p->d_colidx, val, x, yin, dinv, yout, yplain);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}
template <int G, int HINTS>
static int spmv_launch_gh(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  if (G > 1 && !p->tree_sum) return spmv_launch_gho<G, HINTS, (G > 1) ? 1 : 0>(h, p, val, x, yin, dinv, yout, yplain);
  return spmv_launch_gho<G, HINTS, 0>(h, p, val, x, yin, dinv, yout, yplain);
}
template <int G>
static int spmv_launch_g(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  switch (p->hints & 3) {
  case 0: return spmv_launch_gh<G, 0>(h, p, val, x, yin, dinv, yout, yplain);
  case 1: return spmv_launch_gh<G, 1>(h, p, val, x, yin, dinv, yout, yplain);
  case 2: return spmv_launch_gh<G, 2>(h, p, val, x, yin, dinv, yout, yplain);
  default: return spmv_launch_gh<G, 3>(h, p, val, x, yin, dinv, yout, yplain);
  }
}

static int spmv_launch_lanes(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain);

/* hints bit 2: the gathered vector x is given a persisting access-policy window for the duration of the launch (the CSR
   streams then miss as "streaming" and cannot push x out of the L2); the set-aside is raised once per handle */
static int spmv_launch(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain);

/* nb passes in column order; the partial row sums live in yplain (fused-Jacobi form) or in yout itself */
static int spmv_launch_blocked(b200Handle h, b200CsrPlan p, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  struct CsrBlocks *B = p->blk;
  B200_CHECK(B->packed, B200_ERR_ORDER, "b200CsrPlanPackValues must be called after the matrix values changed");
  B200_CHECK(yout && x, B200_ERR_ARG_NULL, "null vector pointer");
  double *acc = yplain ? yplain : yout;
  for (int b = 0; b < B->nb; b++) {
    b200CsrPlan   s    = B->sub[b];
    const double *in   = b == 0 ? yin : acc;
    const double *vb   = B->d_val + B->off[b];
    const bool    last = b == B->nb - 1;
    int           rc;
    s->tree_sum = p->tree_sum;
    if (!last) rc = spmv_launch(h, s, vb, x, in, NULL, acc, NULL);
    else rc = spmv_launch(h, s, vb, x, in, dinv, yout, dinv ? yplain : NULL);
    if (rc) return rc;
  }
  return 0;
}

static int spmv_launch(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null handle/plan");
  if (p->blk && p->m) return spmv_launch_blocked(h, p, x, yin, dinv, yout, yplain);
  if (!(p->hints & 4) || h->l2_persist_max <= 0 || h->l2_window_max <= 0 || !x) return spmv_launch_lanes(h, p, val, x, yin, dinv, yout, yplain);
  if (!h->l2_persist_set) {
    B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)h->l2_persist_max));
    h->l2_persist_set = 1;
  }
  size_t bytes = (size_t)p->n * sizeof(double);
  if (bytes > (size_t)h->l2_window_max) bytes = (size_t)h->l2_window_max;
  cudaStreamAttrValue at;
  memset(&at, 0, sizeof at);
  at.accessPolicyWindow.base_ptr  = (void *)x;
  at.accessPolicyWindow.num_bytes = bytes;
  at.accessPolicyWindow.hitRatio  = bytes <= (size_t)h->l2_persist_max ? 1.0f : (float)h->l2_persist_max / (float)bytes;
  at.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
  at.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
  B200_CUDA(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &at));
  const int rc = spmv_launch_lanes(h, p, val, x, yin, dinv, yout, yplain);
  at.accessPolicyWindow.num_bytes = 0; /* later kernels on this stream run without a window */
  B200_CUDA(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &at));
  return rc;
}

static int spmv_launch_lanes(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *yin, const double *dinv, double *yout, double *yplain)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null handle/plan");
  if (p->m == 0) return 0;
  B200_CHECK(yout && x && (val || p->nnz == 0), B200_ERR_ARG_NULL, "null vector/value pointer");
  if (p->vec_lanes && p->tree_sum) {
    switch (p->vec_lanes) {
    case 2: return spmv_vector_launch<2>(h, p, val, x, yin, dinv, yout, yplain);
    case 4: return spmv_vector_launch<4>(h, p, val, x, yin, dinv, yout, yplain);
    case 8: return spmv_vector_launch<8>(h, p, val, x, yin, dinv, yout, yplain);
    case 16: return spmv_vector_launch<16>(h, p, val, x, yin, dinv, yout, yplain);
    default: return spmv_vector_launch<32>(h, p, val, x, yin, dinv, yout, yplain);
    }
  }
  B200_CHECK((((uintptr_t)val) & 15) == 0 && (((uintptr_t)p->d_colidx) & 15) == 0 && (((uintptr_t)p->d_rowptr) & 15) == 0, B200_ERR_ARG_WRONG, "CSR arrays must be 16-byte aligned (allocate with b200Malloc)");
  switch (p->lanes) {
  case 1: return spmv_launch_g<1>(h, p, val, x, yin, dinv, yout, yplain);
  case 2: return spmv_launch_g<2>(h, p, val, x, yin, dinv, yout, yplain);
  case 4: return spmv_launch_g<4>(h, p, val, x, yin, dinv, yout, yplain);
  case 8: return spmv_launch_g<8>(h, p, val, x, yin, dinv, yout, yplain);
  case 16: return spmv_launch_g<16>(h, p, val, x, yin, dinv, yout, yplain);
  case 32: return spmv_launch_g<32>(h, p, val, x, yin, dinv, yout, yplain);
  }
  B200_CHECK(0, B200_ERR_ARG_WRONGSTATE, "bad plan");
}

extern "C" int b200CsrSpMV(b200Handle h, b200CsrPlan p, const double *val, const double *x, double *y) { return spmv_launch(h, p, val, x, NULL, NULL, y, NULL); }
extern "C" int b200CsrSpMVAdd(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *y, double *z)
{
  B200_CHECK(y, B200_ERR_ARG_NULL, "null y");
  return spmv_launch(h, p, val, x, y, NULL, z, NULL);
}
extern "C" int b200CsrSpMVJacobi(b200Handle h, b200CsrPlan p, const double *val, const double *x, const double *dinv, double *w, double *y)
{
  B200_CHECK(dinv, B200_ERR_ARG_NULL, "null dinv");
  return spmv_launch(h, p, val, x, NULL, dinv, w, y);
}

extern "C" int b200CsrGetDiagonal(b200Handle h, int m, const int *rowptr, const int *colidx, const double *val, double *diag, int *diagpos)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (m <= 0) return 0;
  B200_CHECK(rowptr && diag, B200_ERR_ARG_NULL, "null pointer");
  int g = (m + 255) / 256;
  if (g > h->num_sms * 8) g = h->num_sms * 8;
  get_diagonal_kernel<<<g, 256, 0, h->stream>>>(m, rowptr, colidx, val, diag, diagpos);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

extern "C" int b200JacobiInvertDiagonal(b200Handle h, int64_t n, const double *diag, double *dinv, int *nzero_host)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (nzero_host) *nzero_host = 0;
  if (n <= 0) return 0;
  B200_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  int64_t g = (n + 255) / 256;
  if (g > h->num_sms * 8) g = h->num_sms * 8;
  jacobi_invert_kernel<<<(int)g, 256, 0, h->stream>>>(n, diag, dinv, h->d_flag);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  if (nzero_host) {
    B200_CUDA(cudaMemcpyAsync(h->h_flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    *nzero_host = *h->h_flag;
  }
  return 0;
}

/* ------------------------------------------------------------------ benchmark operator generator (device side) */
/* number of stencil entries in rows [0, r) of the 7-point operator, closed form (no scan needed) */
__host__ __device__ inline int64_t lap7_prefix(int nx, int ny, int nz, int64_t r)
{
  const int64_t nxy = (int64_t)nx * ny;
  if (r <= 0) return 0;
  int64_t z = r / nxy, rem = r % nxy, y = rem / nx, x = rem % nx;
  /* entries = 7*r - (#rows with x==0) - (#x==nx-1) - (#y==0) - (#y==ny-1) - (#z==0) - (#z==nz-1) among rows < r */
  int64_t full_planes = z, full_lines = z * ny + y;
  int64_t cx0 = full_lines + (x > 0 ? 1 : 0);                    /* rows with x == 0 */
  int64_t cx1 = full_lines + 0;                                   /* rows with x == nx-1: only in completed lines */
  int64_t cy0 = full_planes * nx + (y > 0 ? nx : x);              /* y == 0 */
  int64_t cy1 = full_planes * nx + (y == ny - 1 ? x : 0);         /* y == ny-1 */
  int64_t cz0 = z > 0 ? nxy : rem;                                /* z == 0 */
  int64_t cz1 = z > nz - 1 ? nxy : (z == nz - 1 ? rem : 0);       /* z == nz-1 (z == nz when r is one past the end) */
  if (nx == 1) cx1 = cx0;
  return 7 * r - cx0 - cx1 - cy0 - cy1 - cz0 - cz1;
}

__global__ void lap7_fill_kernel(int nx, int ny, int nz, int64_t r0, int64_t r1, int64_t base, int *rowptr, int *colidx, double *val)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nxy    = (int64_t)nx * ny;
  for (int64_t r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += stride) {
    int     x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / nxy);
    int64_t k = lap7_prefix(nx, ny, nz, r) - base;
    rowptr[r - r0] = (int)k;
    if (z > 0) { colidx[k] = (int)(r - nxy); val[k++] = -1.0; }
    if (y > 0) { colidx[k] = (int)(r - nx); val[k++] = -1.0; }
    if (x > 0) { colidx[k] = (int)(r - 1); val[k++] = -1.0; }
    colidx[k] = (int)r; val[k++] = 6.0;
    if (x < nx - 1) { colidx[k] = (int)(r + 1); val[k++] = -1.0; }
    if (y < ny - 1) { colidx[k] = (int)(r + nx); val[k++] = -1.0; }
    if (z < nz - 1) { colidx[k] = (int)(r + nxy); val[k++] = -1.0; }
    if (r == r1 - 1) rowptr[r1 - r0] = (int)k;
  }
}

extern "C" int b200GenLaplace7Nnz(int nx, int ny, int nz, int64_t r0, int64_t r1, int64_t *nnz)
{
  *nnz = lap7_prefix(nx, ny, nz, r1) - lap7_prefix(nx, ny, nz, r0);
  return 0;
}

extern "C" int b200GenLaplace7(b200Handle h, int nx, int ny, int nz, int64_t r0, int64_t r1, int *rowptr, int *colidx, double *val)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(nx > 0 && ny > 0 && nz > 0 && r0 >= 0 && r1 >= r0 && r1 <= (int64_t)nx * ny * nz, B200_ERR_ARG_OUTOFRANGE, "bad grid/row range");
  B200_CHECK((int64_t)nx * ny * nz <= 2147483647LL, B200_ERR_SUP, "global size exceeds 32-bit PetscInt");
  if (r1 == r0) {
    B200_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int), h->stream));
    return 0;
  }
  int64_t g = (r1 - r0 + 255) / 256;
  if (g > h->num_sms * 16) g = h->num_sms * 16;
  lap7_fill_kernel<<<(int)g, 256, 0, h->stream>>>(nx, ny, nz, r0, r1, lap7_prefix(nx, ny, nz, r0), rowptr, colidx, val);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* ------------------------------------------------------------------ compressed-row multadd (off-diagonal block of MATMPIAIJ) */
/* MatMultAdd_SeqAIJ, compressed-row branch (aij.c:1626-1640): only rows that own entries are touched,
   z[ridx[i]] = y[ridx[i]] + sum, strict left-to-right, FMA-free.  In place (z == y) the untouched rows need no copy. */
__global__ void csr_multadd_compressed_kernel(int nrows, const int *__restrict__ cr_i, const int *__restrict__ ridx, const int *__restrict__ colidx, const double *__restrict__ val, const double *__restrict__ x, const double *y, double *z)
{
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
    const int r   = ridx[i];
    double    sum = y[r];
    for (int k = cr_i[i]; k < cr_i[i + 1]; k++) sum = __dadd_rn(sum, __dmul_rn(val[k], __ldg(x + colidx[k])));
    z[r] = sum;
  }
}

extern "C" int b200CsrSpMVAddCompressed(b200Handle h, int nrows_c, const int *d_cr_i, const int *d_rindex, const int *d_colidx, const double *d_val, const double *d_x, const double *d_y, double *d_z)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (nrows_c <= 0) return 0;
  B200_CHECK(d_cr_i && d_rindex && d_colidx && d_val && d_x && d_y && d_z, B200_ERR_ARG_NULL, "null pointer");
  int g = (nrows_c + 255) / 256;
  if (g > h->num_sms * 8) g = h->num_sms * 8;
  csr_multadd_compressed_kernel<<<g, 256, 0, h->stream>>>(nrows_c, d_cr_i, d_rindex, d_colidx, d_val, d_x, d_y, d_z);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* MatMult_MPIAIJ + PCApply_Jacobi fused for the rows that own off-diagonal entries (mpiaij.c:1057-1060 + jacobi.c:354):
   the diagonal-block kernel has already written w = dinv .* (A_d x) for every row; for the nrows_c halo rows this kernel
   recomputes the diagonal-block row sum (left to right from 0.0, exactly as the diagonal-block product does), continues it
   with the off-diagonal entries (the sum MatMultAdd_SeqAIJ starts at y[r], aij.c:1639-1652) and overwrites
   w[r] = dinv[r] * sum -- the reference's association without ever materialising y. */
__global__ void csr_multadd_compressed_jacobi_kernel(int nrows, const int *__restrict__ cr_i, const int *__restrict__ ridx, const int *__restrict__ bj, const double *__restrict__ ba, const double *__restrict__ lvec,
                                                     const int *__restrict__ ai, const int *__restrict__ aj, const double *__restrict__ aa, const double *__restrict__ x, const double *__restrict__ dinv, double *w)
{
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
    const int r   = ridx[i];
    double    sum = 0.0;
    for (int k = ai[r]; k < ai[r + 1]; k++) sum = __dadd_rn(sum, __dmul_rn(aa[k], __ldg(x + aj[k])));
    for (int k = cr_i[i]; k < cr_i[i + 1]; k++) sum = __dadd_rn(sum, __dmul_rn(ba[k], __ldg(lvec + bj[k])));
    w[r] = __dmul_rn(sum, dinv[r]);
  }
}

extern "C" int b200CsrSpMVAddCompressedJacobi(b200Handle h, int nrows_c, const int *d_cr_i, const int *d_rindex, const int *d_bj, const double *d_ba, const double *d_lvec,
                                              const int *d_ai, const int *d_aj, const double *d_aa, const double *d_x, const double *d_dinv, double *d_w)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (nrows_c <= 0) return 0;
  B200_CHECK(d_cr_i && d_rindex && d_bj && d_ba && d_lvec && d_ai && d_aj && d_aa && d_x && d_dinv && d_w, B200_ERR_ARG_NULL, "null pointer");
  int g = (nrows_c + 127) / 128;
  if (g > h->num_sms * 16) g = h->num_sms * 16;
  csr_multadd_compressed_jacobi_kernel<<<g, 128, 0, h->stream>>>(nrows_c, d_cr_i, d_rindex, d_bj, d_ba, d_lvec, d_ai, d_aj, d_aa, d_x, d_dinv, d_w);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* compressed-row analysis (MatCheckCompressedRow, src/mat/utils/compressedrow.c): list of non-empty rows */
__global__ void count_nonempty_kernel(int m, const int *__restrict__ rowptr, int *count)
{
  const int stride = gridDim.x * blockDim.x;
  int       c      = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) c += (rowptr[r + 1] > rowptr[r]);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

extern "C" int b200CsrCountNonemptyRows(b200Handle h, int m, const int *d_rowptr, int *count_host)
{
  B200_CHECK(h && count_host, B200_ERR_ARG_NULL, "null argument");
  *count_host = 0;
  if (m <= 0) return 0;
  B200_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  int g = (m + 255) / 256;
  if (g > h->num_sms * 8) g = h->num_sms * 8;
  count_nonempty_kernel<<<g, 256, 0, h->stream>>>(m, d_rowptr, h->d_flag);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  B200_CUDA(cudaMemcpyAsync(h->h_flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *count_host = *h->h_flag;
  return 0;
}

/* ------------------------------------------------------------------ CSR validation (MatAssemblyEnd_SeqAIJ invariants) */
__global__ void csr_validate_kernel(int m, int n, const int *__restrict__ rowptr, const int *__restrict__ colidx, int *bad /* [2]: row+1 (min), kind */)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    int k0 = rowptr[r], k1 = rowptr[r + 1], kind = 0;
    if (k1 < k0) kind = 3;
    for (int k = k0; k < k1 && !kind; k++) {
      int c = colidx[k];
      if (c < 0 || c >= n) kind = 1;
      else if (k > k0 && c <= colidx[k - 1]) kind = 2;
    }
    if (kind) {
      int old = atomicCAS(&bad[0], 0, (int)r + 1);
      if (old == 0 || (int)r + 1 < old) {
        atomicMin(&bad[0], (int)r + 1);
        bad[1] = kind;
      }
    }
  }
}

/* *bad_row = -1 when the CSR is valid; otherwise a row that violates (kind 1: column out of range, 2: columns not
   strictly increasing, 3: row pointer decreasing) */
extern "C" int b200CsrValidate(b200Handle h, int m, int n, const int *d_rowptr, const int *d_colidx, int *bad_row, int *kind)
{
  B200_CHECK(h && bad_row && kind, B200_ERR_ARG_NULL, "null argument");
  *bad_row = -1;
  *kind    = 0;
  if (m <= 0) return 0;
  B200_CUDA(cudaMemsetAsync(h->d_flag, 0, 2 * sizeof(int), h->stream));
  int64_t g = ((int64_t)m + 255) / 256;
  if (g > h->num_sms * 8) g = h->num_sms * 8;
  csr_validate_kernel<<<(int)g, 256, 0, h->stream>>>(m, n, d_rowptr, d_colidx, h->d_flag);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  B200_CUDA(cudaMemcpyAsync(h->h_flag, h->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (h->h_flag[0]) {
    *bad_row = h->h_flag[0] - 1;
    *kind    = h->h_flag[1];
  }
  return 0;
}
