/*
 * blas1.cu -- BLAS-1 kernels of the Krylov loop for sm_100a (HBM-bound streaming kernels, no tensor cores).
 *
 * Replaces, per SURVEY.md 2.1: cuBLAS axpy/dot/nrm2/scal (vecseqcupm_impl.hpp:519,1125,1749,1473), MDot_kernel<N<=8> +
 * sum_kernel (:1148-1240), MAXPY_kernel<N<=8> (:959-988), thrust pointwise ops (:202-251).
 *
 * Design
 *  - 128-bit (double2) loads/stores on the 16-byte-aligned fast path, grid-stride, all loads of an iteration issued
 *    before the first use (memory-level parallelism), grid = SMs x resident CTAs.
 *  - VecMDot: ONE kernel for any nv <= 32, x is read once, the nv dot products accumulate in registers; the
 *    cross-CTA stage is atomic-free and deterministic: per-CTA partials + a ticket counter, the last CTA to arrive
 *    sums the partials in fixed order with warp shuffles and writes the results to device memory AND to mapped pinned
 *    host memory (so the host needs one stream sync, no extra memcpy).
 *  - VecMAXPY: ONE pass (x read and written once) with the association of VecMAXPY_Seq (dvec2.c:658-693,
 *    petscaxpy.h:125-150) and __dmul_rn/__dadd_rn (no FMA contraction): bit-identical to the CPU reference.
 *    Optionally accumulates ||x_new||^2 in the same pass (fused MAXPY+norm for VecNormalize in KSPGMRESCycle).
 */
#include "b200_internal.h"

#define TPB 256

struct PtrPack {
  const double *p[B200_MAX_NV];
};
struct AlphaPack {
  double a[B200_MAX_NV];
};

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

static int ew_grid(b200Handle h, int64_t nwork)
{
  int64_t blocks = (nwork + TPB - 1) / TPB;
  int64_t cap    = (int64_t)h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

/* ------------------------------------------------------------------ elementwise */
enum { OP_SET, OP_COPY, OP_SCALE, OP_AXPY, OP_AYPX, OP_AXPBY, OP_WAXPY, OP_PMULT, OP_PDIV, OP_RECIP, OP_SHIFT };

template <int OP>
__device__ __forceinline__ double ew_apply(double a, double b, double x, double y)
{
  /* x = first input, y = second input (or old output for in-place ops) */
  if (OP == OP_SET) return a;
  if (OP == OP_COPY) return x;
  if (OP == OP_SCALE) return a * x;
  if (OP == OP_AXPY) return y + a * x;   /* y += a x */
  if (OP == OP_AYPX) return x + a * y;   /* y = x + a y */
  if (OP == OP_AXPBY) return a * x + b * y;
  if (OP == OP_WAXPY) return a * x + y;
  if (OP == OP_PMULT) return __dmul_rn(x, y);
  if (OP == OP_PDIV) return y != 0.0 ? x / y : 0.0;
  if (OP == OP_RECIP) return x != 0.0 ? 1.0 / x : 0.0;
  if (OP == OP_SHIFT) return x + a;
  return 0.0;
}

template <int OP> struct ew_traits {
  static constexpr int nin = (OP == OP_SET) ? 0 : (OP == OP_COPY || OP == OP_SCALE || OP == OP_RECIP || OP == OP_SHIFT) ? 1 : 2;
};

/* in0 = x, in1 = y (second operand; for in-place ops in1 == out) */
template <int OP, int UNR>
__global__ void __launch_bounds__(TPB) ew_kernel_v2(int64_t n, double a, double b, const double *__restrict__ in0, const double *in1, double *out)
{
  constexpr int nin    = ew_traits<OP>::nin;
  const int64_t nvec   = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * TPB;
  int64_t       i      = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const double2 *x2 = reinterpret_cast<const double2 *>(in0);
  const double2 *y2 = reinterpret_cast<const double2 *>(in1);
  double2       *o2 = reinterpret_cast<double2 *>(out);
  for (; i < nvec; i += stride * UNR) {
    double2 xv[UNR], yv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      int64_t k = i + u * stride;
      if (k < nvec) {
        if (nin >= 1) xv[u] = x2[k];
        if (nin >= 2) yv[u] = y2[k];
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      int64_t k = i + u * stride;
      if (k < nvec) {
        double2 r;
        r.x = ew_apply<OP>(a, b, xv[u].x, yv[u].x);
        r.y = ew_apply<OP>(a, b, xv[u].y, yv[u].y);
        o2[k] = r;
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t k = n - 1;
    out[k]    = ew_apply<OP>(a, b, nin >= 1 ? in0[k] : 0.0, nin >= 2 ? in1[k] : 0.0);
  }
}

template <int OP>
__global__ void __launch_bounds__(TPB) ew_kernel_s(int64_t n, double a, double b, const double *in0, const double *in1, double *out)
{
  constexpr int nin    = ew_traits<OP>::nin;
  const int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x; k < n; k += stride) out[k] = ew_apply<OP>(a, b, nin >= 1 ? in0[k] : 0.0, nin >= 2 ? in1[k] : 0.0);
}

template <int OP>
static int ew_launch(b200Handle h, int64_t n, double a, double b, const double *in0, const double *in1, double *out)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative length");
  if (!n) return 0;
  constexpr int nin = ew_traits<OP>::nin;
  B200_CHECK(out && (nin < 1 || in0) && (nin < 2 || in1), B200_ERR_ARG_NULL, "null vector pointer");
  bool al = aligned16(out) && (nin < 1 || aligned16(in0)) && (nin < 2 || aligned16(in1));
  if (al) {
    constexpr int UNR = 4;
    int           g   = ew_grid(h, ((n >> 1) + UNR - 1) / UNR);
    ew_kernel_v2<OP, UNR><<<g, TPB, 0, h->stream>>>(n, a, b, in0, in1, out);
  } else {
    ew_kernel_s<OP><<<ew_grid(h, n), TPB, 0, h->stream>>>(n, a, b, in0, in1, out);
  }
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

extern "C" int b200VecSet(b200Handle h, int64_t n, double alpha, double *x) { return ew_launch<OP_SET>(h, n, alpha, 0, NULL, NULL, x); }
extern "C" int b200VecCopy(b200Handle h, int64_t n, const double *x, double *y)
{
  if (x == y) return 0;
  return ew_launch<OP_COPY>(h, n, 0, 0, x, NULL, y);
}
extern "C" int b200VecScale(b200Handle h, int64_t n, double alpha, double *x) { return ew_launch<OP_SCALE>(h, n, alpha, 0, x, NULL, x); }
extern "C" int b200VecAXPY(b200Handle h, int64_t n, double alpha, const double *x, double *y) { return ew_launch<OP_AXPY>(h, n, alpha, 0, x, y, y); }
extern "C" int b200VecAYPX(b200Handle h, int64_t n, double alpha, const double *x, double *y) { return ew_launch<OP_AYPX>(h, n, alpha, 0, x, y, y); }
extern "C" int b200VecAXPBY(b200Handle h, int64_t n, double alpha, double beta, const double *x, double *y) { return ew_launch<OP_AXPBY>(h, n, alpha, beta, x, y, y); }
extern "C" int b200VecWAXPY(b200Handle h, int64_t n, double alpha, const double *x, const double *y, double *w) { return ew_launch<OP_WAXPY>(h, n, alpha, 0, x, y, w); }
extern "C" int b200VecPointwiseMult(b200Handle h, int64_t n, const double *x, const double *y, double *w) { return ew_launch<OP_PMULT>(h, n, 0, 0, x, y, w); }
extern "C" int b200VecPointwiseDivide(b200Handle h, int64_t n, const double *x, const double *y, double *w) { return ew_launch<OP_PDIV>(h, n, 0, 0, x, y, w); }
extern "C" int b200VecReciprocal(b200Handle h, int64_t n, double *x) { return ew_launch<OP_RECIP>(h, n, 0, 0, x, NULL, x); }
extern "C" int b200VecShift(b200Handle h, int64_t n, double s, double *x) { return ew_launch<OP_SHIFT>(h, n, s, 0, x, NULL, x); }

/* ------------------------------------------------------------------ reductions: shared final stage */
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

/* Block-level + grid-level reduction of NV per-thread accumulators.  RED: 0 = sum, 1 = max.
   partials layout [NV][gridDim.x]; the last CTA (ticket counter) reduces them in a fixed order. */
template <int NV, int RED>
__device__ __forceinline__ void grid_reduce(double (&acc)[NV], double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  __shared__ double   s_part[NV][TPB / 32];
  __shared__ unsigned s_last;
  const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double v = RED ? warp_max(acc[j]) : warp_sum(acc[j]);
    if (lane == 0) s_part[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = s_part[threadIdx.x][0];
#pragma unroll
    for (int w = 1; w < TPB / 32; w++) v = RED ? fmax(v, s_part[threadIdx.x][w]) : v + s_part[threadIdx.x][w];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int j = warp; j < NV; j += TPB / 32) {
      double v = RED ? -INFINITY : 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) {
        double p = __ldcg(&partials[(size_t)j * gridDim.x + b]);
        v        = RED ? fmax(v, p) : v + p;
      }
      v = RED ? warp_max(v) : warp_sum(v);
      if (lane == 0) {
        d_result[j] = v;
        if (h_result) h_result[j] = v;
      }
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

/* ------------------------------------------------------------------ MDot */
/* Generalised final stage: this CTA holds NVT accumulators per thread for the vectors [jbase, jbase+NVT); it is the
   vb-th of nvb "virtual" CTAs walking x.  partials layout [nv][nvb]; the last CTA of the whole grid reduces them. */
template <int NVT>
__device__ __forceinline__ void grid_reduce_x(double (&acc)[NVT], int jbase, unsigned vb, unsigned nvb, int nv_total, double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  __shared__ double   s_part[NVT][TPB / 32];
  __shared__ unsigned s_last;
  const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NVT; j++) {
    double v = warp_sum(acc[j]);
    if (lane == 0) s_part[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NVT) {
    double v = s_part[threadIdx.x][0];
#pragma unroll
    for (int w = 1; w < TPB / 32; w++) v += s_part[threadIdx.x][w];
    partials[(size_t)(jbase + threadIdx.x) * nvb + vb] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int j = warp; j < nv_total; j += TPB / 32) {
      double v = 0.0;
      for (unsigned b = lane; b < nvb; b += 32) v += __ldcg(&partials[(size_t)j * nvb + b]);
      v = warp_sum(v);
      if (lane == 0) {
        d_result[j] = v;
        if (h_result) h_result[j] = v;
      }
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

/* Each thread owns NVT <= 16 accumulators.  SPLIT (nv > 16): the vectors are divided into two halves handled by the even
   and the odd CTAs, which walk the same elements of x side by side (x comes from HBM once; the partner's read hits L2).
   That bounds the register footprint (~64-80 registers -> 3-4 resident CTAs per SM) for every nv <= 32; without it
   ptxas allocates up to 148 registers for some nv and occupancy collapses (profiles/round1_notes.md). */
template <int NVT, int UNR, bool VEC2, bool SPLIT>
__global__ void __launch_bounds__(TPB, (NVT <= 12 ? 4 : 3)) mdot_kernel(int64_t n, const double *__restrict__ x, PtrPack y, int nv_total, double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  const int      sub   = SPLIT ? (int)(blockIdx.x & 1u) : 0;
  const unsigned vb    = SPLIT ? (blockIdx.x >> 1) : blockIdx.x;
  const unsigned nvb   = SPLIT ? (gridDim.x >> 1) : gridDim.x;
  const int      jbase = sub * NVT;
  double         acc[NVT];
#pragma unroll
  for (int j = 0; j < NVT; j++) acc[j] = 0.0;
  const int64_t stride = (int64_t)nvb * TPB;
  int64_t       i      = (int64_t)vb * TPB + threadIdx.x;
  if (VEC2) {
    const int64_t  nvec = n >> 1;
    const double2 *x2   = reinterpret_cast<const double2 *>(x);
    for (; i < nvec; i += stride * UNR) {
      double2 xv[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        int64_t k = i + u * stride;
        xv[u]     = k < nvec ? x2[k] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < NVT; j++) {
        const double2 *y2 = reinterpret_cast<const double2 *>(y.p[jbase + j]);
        double2        yv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int64_t k = i + u * stride;
          yv[u]     = k < nvec ? __ldg(&y2[k]) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          acc[j] = fma(xv[u].x, yv[u].x, acc[j]);
          acc[j] = fma(xv[u].y, yv[u].y, acc[j]);
        }
      }
    }
    if ((n & 1) && vb == 0 && threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < NVT; j++) acc[j] = fma(x[n - 1], y.p[jbase + j][n - 1], acc[j]);
    }
  } else {
    for (; i < n; i += stride) {
      double xv = x[i];
#pragma unroll
      for (int j = 0; j < NVT; j++) acc[j] = fma(xv, y.p[jbase + j][i], acc[j]);
    }
  }
  grid_reduce_x<NVT>(acc, jbase, vb, nvb, nv_total, partials, counter, d_result, h_result);
}

template <int NVT, bool SPLIT>
static int mdot_launch_nv(b200Handle h, int64_t n, int nv, const double *x, const PtrPack &y, bool vec2, double *d_result, double *h_result)
{
  constexpr int UNR = NVT <= 2 ? 4 : (NVT <= 6 ? 2 : 1);
  static int    occ[2] = {0, 0};
  void (*kern)(int64_t, const double *, PtrPack, int, double *, unsigned int *, double *, double *) = vec2 ? mdot_kernel<NVT, UNR, true, SPLIT> : mdot_kernel<NVT, 1, false, SPLIT>;
  if (!occ[vec2]) {
    int o = 1;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, TPB, 0));
    occ[vec2] = o < 1 ? 1 : (o > 8 ? 8 : o);
  }
  int64_t work   = vec2 ? ((n >> 1) + UNR - 1) / UNR : n;
  int64_t blocks = (work + TPB - 1) / TPB; /* virtual CTAs */
  int64_t cap    = (int64_t)h->num_sms * occ[vec2];
  if (cap > B200_RED_MAXGRID) cap = B200_RED_MAXGRID;
  if (SPLIT) cap /= 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(int)(SPLIT ? 2 * blocks : blocks), TPB, 0, h->stream>>>(n, x, y, nv, h->d_partials, h->d_counter, d_result, h_result);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

static int mdot_dispatch(b200Handle h, int64_t n, int nv, const double *x, const double *const *yp, double *d_result, double *h_result)
{
  PtrPack y;
  bool    vec2 = aligned16(x);
  for (int j = 0; j < nv; j++) {
    y.p[j] = yp[j];
    vec2   = vec2 && aligned16(yp[j]);
  }
  for (int j = nv; j < B200_MAX_NV; j++) y.p[j] = yp[0]; /* padding slots of the split layout read a valid vector */
  switch (nv <= 16 ? nv : 100 + (nv + 1) / 2) {
#define C_(N) \
  case N: return mdot_launch_nv<N, false>(h, n, nv, x, y, vec2, d_result, h_result);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14) C_(15) C_(16)
#undef C_
#define S_(N) \
  case 100 + N: return mdot_launch_nv<N, true>(h, n, nv, x, y, vec2, d_result, h_result);
    S_(9) S_(10) S_(11) S_(12) S_(13) S_(14) S_(15) S_(16)
#undef S_
  }
  B200_CHECK(0, B200_ERR_ARG_OUTOFRANGE, "nv=%d out of range", nv);
}

/* device-result variant; chunks nv > 32 */
extern "C" int b200VecMDotAsync(b200Handle h, int64_t n, int nv, const double *x, const double *const *y, double *d_result)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(nv >= 0 && n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  if (!nv) return 0;
  B200_CHECK(x && y && d_result, B200_ERR_ARG_NULL, "null pointer");
  if (n == 0) {
    B200_CUDA(cudaMemsetAsync(d_result, 0, sizeof(double) * nv, h->stream));
    return 0;
  }
  for (int j0 = 0; j0 < nv; j0 += B200_MAX_NV) {
    int c  = nv - j0 < B200_MAX_NV ? nv - j0 : B200_MAX_NV;
    int rc = mdot_dispatch(h, n, c, x, y + j0, d_result + j0, NULL);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int b200VecMDot(b200Handle h, int64_t n, int nv, const double *x, const double *const *y, double *result)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(nv >= 0 && n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  if (!nv) return 0;
  B200_CHECK(x && y && result, B200_ERR_ARG_NULL, "null pointer");
  if (n == 0) {
    for (int j = 0; j < nv; j++) result[j] = 0.0;
    return 0;
  }
  for (int j0 = 0; j0 < nv; j0 += B200_MAX_NV) {
    int c  = nv - j0 < B200_MAX_NV ? nv - j0 : B200_MAX_NV;
    int rc = mdot_dispatch(h, n, c, x, y + j0, h->d_result, h->h_result_dev);
    if (rc) return rc;
    B200_CUDA(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < c; j++) result[j0 + j] = h->h_result[j];
  }
  return 0;
}

extern "C" int b200VecDot(b200Handle h, int64_t n, const double *x, const double *y, double *result)
{
  /* VecDot(x,y) = y^H x ; real scalars: symmetric */
  return b200VecMDot(h, n, 1, x, &y, result);
}

/* ------------------------------------------------------------------ generic single reductions (norms, sum, max, min) */
enum { R_SUMSQ, R_SUMABS, R_SUM, R_MAXABS, R_MAX, R_MIN };

template <int R>
__global__ void __launch_bounds__(TPB) reduce_kernel(int64_t n, const double *__restrict__ x, double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  constexpr bool ismax = (R == R_MAXABS || R == R_MAX || R == R_MIN);
  double         acc[1];
  acc[0] = ismax ? -INFINITY : 0.0;
  const int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
    double v = x[i];
    if (R == R_SUMSQ) acc[0] = fma(v, v, acc[0]);
    else if (R == R_SUMABS) acc[0] += fabs(v);
    else if (R == R_SUM) acc[0] += v;
    else if (R == R_MAXABS) acc[0] = fmax(acc[0], fabs(v));
    else if (R == R_MAX) acc[0] = fmax(acc[0], v);
    else acc[0] = fmax(acc[0], -v);
  }
  grid_reduce<1, ismax ? 1 : 0>(acc, partials, counter, d_result, h_result);
}

template <int R>
static int reduce_launch(b200Handle h, int64_t n, const double *x, double *result)
{
  B200_CHECK(h && result, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative length");
  if (!n) {
    *result = (R == R_MAX || R == R_MIN) ? -INFINITY : 0.0;
    return 0;
  }
  B200_CHECK(x, B200_ERR_ARG_NULL, "null vector");
  int g = ew_grid(h, (n + 3) / 4);
  if (g > B200_RED_MAXGRID) g = B200_RED_MAXGRID;
  reduce_kernel<R><<<g, TPB, 0, h->stream>>>(n, x, h->d_partials, h->d_counter, h->d_result, h->h_result_dev);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *result = h->h_result[0];
  return 0;
}

extern "C" int b200VecNorm2(b200Handle h, int64_t n, const double *x, double *result)
{
  /* VecNorm_Seq NORM_2 = sqrt(ddot(x,x)) (bvec2.c:201-205): same kernel as the dot product */
  int rc = b200VecMDot(h, n, 1, x, &x, result);
  if (!rc) *result = sqrt(*result);
  return rc;
}

extern "C" int b200VecNorm(b200Handle h, int64_t n, const double *x, int type, double *result)
{
  if (type == 1) return b200VecNorm2(h, n, x, result);
  if (type == 0) return reduce_launch<R_SUMABS>(h, n, x, result);
  if (type == 3) {
    int rc = reduce_launch<R_MAXABS>(h, n, x, result);
    if (!rc && n == 0) *result = 0.0;
    return rc;
  }
  B200_CHECK(0, B200_ERR_ARG_OUTOFRANGE, "unknown norm type %d", type);
}
extern "C" int b200VecSum(b200Handle h, int64_t n, const double *x, double *result) { return reduce_launch<R_SUM>(h, n, x, result); }

__global__ void first_index_kernel(int64_t n, const double *__restrict__ x, double val, long long *idx)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (x[i] == val) atomicMin(idx, (long long)i);
}

static int maxmin(b200Handle h, int64_t n, const double *x, int64_t *idx, double *result, bool ismax)
{
  int rc = ismax ? reduce_launch<R_MAX>(h, n, x, result) : reduce_launch<R_MIN>(h, n, x, result);
  if (rc) return rc;
  if (!ismax) *result = -*result;
  if (!n) {
    *result = ismax ? -1.7976931348623157e308 : 1.7976931348623157e308; /* PETSC_MIN_REAL / PETSC_MAX_REAL */
    if (idx) *idx = -1;
    return 0;
  }
  if (idx) {
    long long big = 0x7fffffffffffffffLL;
    B200_CUDA(cudaMemcpyAsync(h->d_idx, &big, sizeof big, cudaMemcpyHostToDevice, h->stream));
    first_index_kernel<<<ew_grid(h, n), TPB, 0, h->stream>>>(n, x, *result, h->d_idx);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
    B200_CUDA(cudaMemcpyAsync(h->h_flag, h->d_idx, sizeof big, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    *idx = *(long long *)h->h_flag;
  }
  return 0;
}
extern "C" int b200VecMax(b200Handle h, int64_t n, const double *x, int64_t *idx, double *r) { return maxmin(h, n, x, idx, r, true); }
extern "C" int b200VecMin(b200Handle h, int64_t n, const double *x, int64_t *idx, double *r) { return maxmin(h, n, x, idx, r, false); }

/* ------------------------------------------------------------------ MAXPY (+ fused norm) */
/* VecMAXPY_Seq (dvec2.c:658-693, petscaxpy.h:125-150): the nv&3 remainder group first, then groups of 4; inside a group
   the products are summed left to right and the group total is added to x.  No FMA contraction -> bit-identical.
   The groups are walked two at a time by a real (not unrolled) loop: 8 x 128-bit loads in flight per thread and ~56
   registers, i.e. 4 resident CTAs per SM for every nv (the fully unrolled form made ptxas allocate 102-128 registers
   for nv >= 27 -> 2 CTAs/SM and 69% of peak, profiles/round1_notes.md). */
template <int CNT>
__device__ __forceinline__ double maxpy_group(const double *a, const double (&v)[CNT])
{
  double t = __dmul_rn(a[0], v[0]);
#pragma unroll
  for (int k = 1; k < CNT; k++) t = __dadd_rn(t, __dmul_rn(a[k], v[k]));
  return t;
}

template <int NV, bool NORM, bool VEC2>
__global__ void __launch_bounds__(TPB, 4) maxpy_kernel(int64_t n, double *x, PtrPack y, AlphaPack al, double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  constexpr int G0     = NV & 3;          /* remainder group size */
  constexpr int NG     = (NV - G0) / 4;   /* groups of four */
  double        acc[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * TPB;
  int64_t       i      = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (VEC2) {
    const int64_t nvec = n >> 1;
    double2      *x2   = reinterpret_cast<double2 *>(x);
    for (; i < nvec; i += stride) {
      double2 xv = x2[i];
      if (G0) {
        double v0[G0 ? G0 : 1], v1[G0 ? G0 : 1];
#pragma unroll
        for (int k = 0; k < G0; k++) {
          double2 t = __ldg(reinterpret_cast<const double2 *>(y.p[k]) + i);
          v0[k] = t.x; v1[k] = t.y;
        }
        xv.x = __dadd_rn(xv.x, maxpy_group<(G0 ? G0 : 1)>(al.a, v0));
        xv.y = __dadd_rn(xv.y, maxpy_group<(G0 ? G0 : 1)>(al.a, v1));
      }
      int j = G0;
#pragma unroll 1
      for (int p = 0; p < NG / 2; p++, j += 8) {
        double v0[8], v1[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          double2 t = __ldg(reinterpret_cast<const double2 *>(y.p[j + k]) + i);
          v0[k] = t.x; v1[k] = t.y;
        }
        double a8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) a8[k] = al.a[j + k];
        const double(&v0a)[4] = *reinterpret_cast<const double(*)[4]>(v0);
        const double(&v0b)[4] = *reinterpret_cast<const double(*)[4]>(v0 + 4);
        const double(&v1a)[4] = *reinterpret_cast<const double(*)[4]>(v1);
        const double(&v1b)[4] = *reinterpret_cast<const double(*)[4]>(v1 + 4);
        xv.x = __dadd_rn(xv.x, maxpy_group<4>(a8, v0a));
        xv.y = __dadd_rn(xv.y, maxpy_group<4>(a8, v1a));
        xv.x = __dadd_rn(xv.x, maxpy_group<4>(a8 + 4, v0b));
        xv.y = __dadd_rn(xv.y, maxpy_group<4>(a8 + 4, v1b));
      }
      if (NG & 1) {
        double v0[4], v1[4], a4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double2 t = __ldg(reinterpret_cast<const double2 *>(y.p[NV - 4 + k]) + i);
          v0[k] = t.x; v1[k] = t.y;
          a4[k] = al.a[NV - 4 + k];
        }
        xv.x = __dadd_rn(xv.x, maxpy_group<4>(a4, v0));
        xv.y = __dadd_rn(xv.y, maxpy_group<4>(a4, v1));
      }
      x2[i] = xv;
      if (NORM) {
        acc[0] = fma(xv.x, xv.x, acc[0]);
        acc[0] = fma(xv.y, xv.y, acc[0]);
      }
    }
  }
  /* scalar path: odd tail of the vector path, or everything when some pointer is only 8-byte aligned */
  {
    int64_t s0 = VEC2 ? (n & ~(int64_t)1) : 0;
    int64_t k0 = VEC2 ? ((blockIdx.x == 0 && threadIdx.x == 0) ? s0 : n) : (int64_t)blockIdx.x * TPB + threadIdx.x;
    int64_t st = VEC2 ? 1 : stride;
    for (int64_t e = k0; e < n; e += st) {
      double xv = x[e];
      if (G0) {
        double v[G0 ? G0 : 1];
#pragma unroll
        for (int k = 0; k < G0; k++) v[k] = y.p[k][e];
        xv = __dadd_rn(xv, maxpy_group<(G0 ? G0 : 1)>(al.a, v));
      }
#pragma unroll 1
      for (int g = 0; g < NG; g++) {
        double v[4], a4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          v[k]  = y.p[G0 + 4 * g + k][e];
          a4[k] = al.a[G0 + 4 * g + k];
        }
        xv = __dadd_rn(xv, maxpy_group<4>(a4, v));
      }
      x[e] = xv;
      if (NORM) acc[0] = fma(xv, xv, acc[0]);
    }
  }
  if (NORM) grid_reduce<1, 0>(acc, partials, counter, d_result, h_result);
}

template <int NV>
static int maxpy_launch_nv(b200Handle h, int64_t n, double *x, const PtrPack &y, const AlphaPack &al, bool vec2, bool norm, double *d_result, double *h_result)
{
  typedef void (*kern_t)(int64_t, double *, PtrPack, AlphaPack, double *, unsigned int *, double *, double *);
  static int occ[4] = {0, 0, 0, 0};
  int        v      = (vec2 ? 1 : 0) + (norm ? 2 : 0);
  kern_t     kern   = vec2 ? (norm ? (kern_t)maxpy_kernel<NV, true, true> : (kern_t)maxpy_kernel<NV, false, true>) : (norm ? (kern_t)maxpy_kernel<NV, true, false> : (kern_t)maxpy_kernel<NV, false, false>);
  if (!occ[v]) {
    int o = 1;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, TPB, 0));
    occ[v] = o < 1 ? 1 : (o > 8 ? 8 : o);
  }
  int64_t work   = vec2 ? (n >> 1) : n;
  int64_t blocks = (work + TPB - 1) / TPB;
  int64_t cap    = (int64_t)h->num_sms * occ[v];
  if (cap > B200_RED_MAXGRID) cap = B200_RED_MAXGRID;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(int)blocks, TPB, 0, h->stream>>>(n, x, y, al, h->d_partials, h->d_counter, d_result, h_result);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

static int maxpy_dispatch(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *yp, double *x, bool norm, double *d_result, double *h_result)
{
  PtrPack   y;
  AlphaPack al;
  bool      vec2 = aligned16(x);
  for (int j = 0; j < nv; j++) {
    y.p[j]  = yp[j];
    al.a[j] = alpha[j];
    vec2    = vec2 && aligned16(yp[j]);
  }
  for (int j = nv; j < B200_MAX_NV; j++) {
    y.p[j]  = yp[0];
    al.a[j] = 0.0;
  }
  switch (nv) {
#define C_(N) \
  case N: return maxpy_launch_nv<N>(h, n, x, y, al, vec2, norm, d_result, h_result);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14) C_(15) C_(16)
    C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27) C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
  }
  B200_CHECK(0, B200_ERR_ARG_OUTOFRANGE, "nv=%d out of range", nv);
}

/* chunking for nv > 32 must respect the reference grouping: the remainder group (nv & 3) comes first, so the first
   chunk takes (nv & 3) + a multiple of 4 vectors and all later chunks are multiples of 4 */
static int maxpy_any(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *x, bool norm, double *d_result, double *h_result)
{
  int j0 = 0;
  while (j0 < nv) {
    int left = nv - j0, c;
    if (left <= B200_MAX_NV) c = left;
    else c = (j0 == 0) ? ((nv & 3) + ((B200_MAX_NV - (nv & 3)) & ~3)) : B200_MAX_NV;
    bool last = (j0 + c == nv);
    int  rc   = maxpy_dispatch(h, n, c, alpha + j0, y + j0, x, norm && last, d_result, h_result);
    if (rc) return rc;
    j0 += c;
  }
  return 0;
}

extern "C" int b200VecMAXPYAsync(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *x, double *d_sumsq)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(nv >= 0 && n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  if (n == 0 || nv == 0) {
    if (d_sumsq && n == 0) B200_CUDA(cudaMemsetAsync(d_sumsq, 0, sizeof(double), h->stream));
    if (d_sumsq && n > 0) { /* no update: plain sum of squares */
      const double *xp = x;
      return b200VecMDotAsync(h, n, 1, x, &xp, d_sumsq);
    }
    return 0;
  }
  B200_CHECK(x && y && alpha, B200_ERR_ARG_NULL, "null pointer");
  return maxpy_any(h, n, nv, alpha, y, x, d_sumsq != NULL, d_sumsq, NULL);
}

extern "C" int b200VecMAXPY(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *x, double *norm2_out)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(nv >= 0 && n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  if (n == 0 || nv == 0) {
    if (norm2_out) return b200VecNorm2(h, n, x, norm2_out);
    return 0;
  }
  B200_CHECK(x && y && alpha, B200_ERR_ARG_NULL, "null pointer");
  int rc = maxpy_any(h, n, nv, alpha, y, x, norm2_out != NULL, h->d_result, h->h_result_dev);
  if (rc) return rc;
  if (norm2_out) {
    B200_CUDA(cudaStreamSynchronize(h->stream));
    *norm2_out = sqrt(h->h_result[0]);
  }
  return 0;
}

/* ------------------------------------------------------------------ KSPPIPECG vector recurrences in one pass
   pipecg.c:124-141: four VecAYPX (or VecCopy in the first iteration) and four VecAXPY, here one kernel: 10 vectors read,
   8 written (18 n doubles of traffic instead of 24 n, one launch instead of eight).  Every entry goes through exactly the
   expressions of the separate kernels (x + a*y, y + a*x, same contraction), so the result is bit-identical to them. */
template <bool VEC2>
__global__ void __launch_bounds__(TPB) pipecg_update_kernel(int64_t n, double alpha, double beta, int first, const double *__restrict__ vn, const double *__restrict__ vm, double *u, double *w, double *z, double *q, double *p, double *s, double *x, double *r)
{
  const int64_t stride = (int64_t)gridDim.x * TPB;
  const double  na     = -alpha;
  auto one = [&](double nn, double mm, double &uu, double &ww, double &zz, double &qq, double &pp, double &ss, double &xx, double &rr) {
    if (first) {
      zz = nn; qq = mm; pp = uu; ss = ww;
    } else {
      zz = nn + beta * zz; /* VecAYPX(Z, beta, N) */
      qq = mm + beta * qq;
      pp = uu + beta * pp;
      ss = ww + beta * ss;
    }
    xx = xx + alpha * pp;  /* VecAXPY(X, alpha, P) */
    uu = uu + na * qq;     /* VecAXPY(U, -alpha, Q) */
    ww = ww + na * zz;
    rr = rr + na * ss;
  };
  if (VEC2) {
    const int64_t  nvec = n >> 1;
    const double2 *n2 = reinterpret_cast<const double2 *>(vn), *m2 = reinterpret_cast<const double2 *>(vm);
    double2       *u2 = reinterpret_cast<double2 *>(u), *w2 = reinterpret_cast<double2 *>(w), *z2 = reinterpret_cast<double2 *>(z), *q2 = reinterpret_cast<double2 *>(q);
    double2       *p2 = reinterpret_cast<double2 *>(p), *s2 = reinterpret_cast<double2 *>(s), *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < nvec; i += stride) {
      const double2 nn = n2[i], mm = m2[i];
      double2       uu = u2[i], ww = w2[i], xx = x2[i], rr = r2[i], zz, qq, pp, ss;
      if (first) zz = qq = pp = ss = make_double2(0.0, 0.0);
      else { zz = z2[i]; qq = q2[i]; pp = p2[i]; ss = s2[i]; }
      one(nn.x, mm.x, uu.x, ww.x, zz.x, qq.x, pp.x, ss.x, xx.x, rr.x);
      one(nn.y, mm.y, uu.y, ww.y, zz.y, qq.y, pp.y, ss.y, xx.y, rr.y);
      z2[i] = zz; q2[i] = qq; p2[i] = pp; s2[i] = ss; x2[i] = xx; u2[i] = uu; w2[i] = ww; r2[i] = rr;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const int64_t i = n - 1;
      double        uu = u[i], ww = w[i], xx = x[i], rr = r[i], zz = first ? 0.0 : z[i], qq = first ? 0.0 : q[i], pp = first ? 0.0 : p[i], ss = first ? 0.0 : s[i];
      one(vn[i], vm[i], uu, ww, zz, qq, pp, ss, xx, rr);
      z[i] = zz; q[i] = qq; p[i] = pp; s[i] = ss; x[i] = xx; u[i] = uu; w[i] = ww; r[i] = rr;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
      double uu = u[i], ww = w[i], xx = x[i], rr = r[i], zz = first ? 0.0 : z[i], qq = first ? 0.0 : q[i], pp = first ? 0.0 : p[i], ss = first ? 0.0 : s[i];
      one(vn[i], vm[i], uu, ww, zz, qq, pp, ss, xx, rr);
      z[i] = zz; q[i] = qq; p[i] = pp; s[i] = ss; x[i] = xx; u[i] = uu; w[i] = ww; r[i] = rr;
    }
  }
}

extern "C" int b200VecPipeCGUpdate(b200Handle h, int64_t n, double alpha, double beta, int first, const double *d_n, const double *d_m, double *d_u, double *d_w, double *d_z, double *d_q, double *d_p, double *d_s, double *d_x, double *d_r)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CHECK(n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative length");
  if (!n) return 0;
  B200_CHECK(d_n && d_m && d_u && d_w && d_z && d_q && d_p && d_s && d_x && d_r, B200_ERR_ARG_NULL, "null vector");
  const bool vec2 = aligned16(d_n) && aligned16(d_m) && aligned16(d_u) && aligned16(d_w) && aligned16(d_z) && aligned16(d_q) && aligned16(d_p) && aligned16(d_s) && aligned16(d_x) && aligned16(d_r);
  const int  g    = ew_grid(h, vec2 ? (n >> 1) : n);
  if (vec2) pipecg_update_kernel<true><<<g, TPB, 0, h->stream>>>(n, alpha, beta, first, d_n, d_m, d_u, d_w, d_z, d_q, d_p, d_s, d_x, d_r);
  else pipecg_update_kernel<false><<<g, TPB, 0, h->stream>>>(n, alpha, beta, first, d_n, d_m, d_u, d_w, d_z, d_q, d_p, d_s, d_x, d_r);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* ------------------------------------------------------------------ fused AXPY + dot */
template <bool VEC2>
__global__ void __launch_bounds__(TPB) axpy_dot_kernel(int64_t n, double a, const double *__restrict__ x, double *y, const double *z, double *partials, unsigned int *counter, double *d_result, double *h_result)
{
  double        acc[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * TPB;
  int64_t       i      = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const bool    self   = (z == y);
  if (VEC2) {
    const int64_t  nvec = n >> 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(x), *z2 = reinterpret_cast<const double2 *>(z);
    double2       *y2 = reinterpret_cast<double2 *>(y);
    for (; i < nvec; i += stride) {
      double2 xv = x2[i], yv = y2[i], zv;
      if (!self) zv = z2[i];
      yv.x = yv.x + a * xv.x;
      yv.y = yv.y + a * xv.y;
      if (self) zv = yv;
      y2[i]  = yv;
      acc[0] = fma(yv.x, zv.x, acc[0]);
      acc[0] = fma(yv.y, zv.y, acc[0]);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double yv = y[n - 1] + a * x[n - 1];
      y[n - 1]  = yv;
      acc[0]    = fma(yv, self ? yv : z[n - 1], acc[0]);
    }
  } else {
    for (; i < n; i += stride) {
      double yv = y[i] + a * x[i];
      y[i]      = yv;
      acc[0]    = fma(yv, self ? yv : z[i], acc[0]);
    }
  }
  grid_reduce<1, 0>(acc, partials, counter, d_result, h_result);
}

extern "C" int b200VecAXPYDot(b200Handle h, int64_t n, double alpha, const double *x, double *y, const double *z, double *result)
{
  B200_CHECK(h && result, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative length");
  if (!n) {
    *result = 0.0;
    return 0;
  }
  B200_CHECK(x && y && z, B200_ERR_ARG_NULL, "null vector");
  bool vec2 = aligned16(x) && aligned16(y) && aligned16(z);
  int  g    = ew_grid(h, vec2 ? (n >> 1) : n);
  if (g > B200_RED_MAXGRID) g = B200_RED_MAXGRID;
  if (vec2) axpy_dot_kernel<true><<<g, TPB, 0, h->stream>>>(n, alpha, x, y, z, h->d_partials, h->d_counter, h->d_result, h->h_result_dev);
  else axpy_dot_kernel<false><<<g, TPB, 0, h->stream>>>(n, alpha, x, y, z, h->d_partials, h->d_counter, h->d_result, h->h_result_dev);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *result = h->h_result[0];
  return 0;
}
