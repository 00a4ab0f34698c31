/*
 * sf.cu -- indexed scatter with an operation on device memory: the local part of a PetscSF broadcast / reduction.
 *
 * Replaces the host loops the reference runs for a star forest whose roots and leaves live in the same process
 * (PetscSFLinkScatterLocal, sfpack.c:1082, with the ScatterAnd<Op> bodies of sfpack.c:190-222): for i = 0 .. n-1 in this order
 *     dst[didx[i]*bs + c]  =  dst[didx[i]*bs + c]  <op>  src[sidx[i]*bs + c],      c = 0 .. bs-1.
 * A broadcast has dst = leaf data, src = root data; a reduction swaps them, so several entries may hit the same destination
 * (one root with many leaves).  The reference applies them in index order; so does this file: the plan groups the entries by
 * destination with a STABLE counting sort on the host (setup time) and the kernel walks each group sequentially -- no atomics, and
 * floating-point sums come out in the reference's association, bit for bit.  Destinations that are hit once (every VecScatter
 * with distinct "to" indices) take the one-thread-per-entry kernel, with the entries re-ordered by destination at plan time
 * ("pull" order: coalesced writes, gathered reads; a permutation of a range then needs no destination indices at all).
 *
 * HBM-bound gather/scatter: per entry 8 B of indices (none when a side is contiguous), bs*8 B read, bs*8 B written (+ bs*8 B read
 * for a non-REPLACE op).  Not on the Krylov inner loop; it is what VecScatterBegin/End cost on device vectors.
 */
#include "b200_internal.h"
#include <stdlib.h>
#include <string.h>

struct b200IndexedPlan_s {
  int64_t n;        /* entries */
  int64_t ngroups;  /* distinct destinations */
  int     grouped;  /* some destination is hit more than once: use the grouped kernel */
  int     src_contig, dst_contig; /* sidx[i] = s0 + i / didx[i] = d0 + i: no index loads */
  int     s0, d0;
  int64_t src_extent, dst_extent; /* 1 + largest index: bounds the buffers the kernels touch */
  int     k_src_contig, k_dst_contig, k_s0, k_d0; /* what the ungrouped kernel sees: entries re-ordered by destination ("pull" order) */
  int    *d_sidx, *d_didx;        /* [n] ungrouped kernel, pull order */
  int    *d_gdst, *d_goff, *d_gsrc; /* grouped: destination of group g, entries goff[g]..goff[g+1] of gsrc (in entry order) */
};

template <typename T, int OP> __device__ __forceinline__ T sf_apply(T s, T t);
/* PetscMax(a,b) = a < b ? b : a and PetscMin(a,b) = a < b ? a : b (petscmath.h), applied as s = op(s, t) (sfpack.c:19) */
template <> __device__ __forceinline__ double sf_apply<double, B200_SF_REPLACE>(double s, double t) { (void)s; return t; }
template <> __device__ __forceinline__ double sf_apply<double, B200_SF_SUM>(double s, double t) { return __dadd_rn(s, t); }
template <> __device__ __forceinline__ double sf_apply<double, B200_SF_PROD>(double s, double t) { return __dmul_rn(s, t); }
template <> __device__ __forceinline__ double sf_apply<double, B200_SF_MAX>(double s, double t) { return (s < t) ? t : s; }
template <> __device__ __forceinline__ double sf_apply<double, B200_SF_MIN>(double s, double t) { return (s < t) ? s : t; }
template <> __device__ __forceinline__ int sf_apply<int, B200_SF_REPLACE>(int s, int t) { (void)s; return t; }
template <> __device__ __forceinline__ int sf_apply<int, B200_SF_SUM>(int s, int t) { return (int)((unsigned)s + (unsigned)t); }
template <> __device__ __forceinline__ int sf_apply<int, B200_SF_PROD>(int s, int t) { return (int)((unsigned)s * (unsigned)t); }
template <> __device__ __forceinline__ int sf_apply<int, B200_SF_MAX>(int s, int t) { return (s < t) ? t : s; }
template <> __device__ __forceinline__ int sf_apply<int, B200_SF_MIN>(int s, int t) { return (s < t) ? s : t; }

/* one thread per (entry, component): destinations are distinct */
template <typename T, int OP, bool BS1> __global__ void __launch_bounds__(256) sf_scatter_kernel(int64_t n, int bs, const int *__restrict__ sidx, int s0, const int *__restrict__ didx, int d0, const T *__restrict__ src, T *dst)
{
  const int64_t total = n * bs, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = BS1 ? e : e / bs; /* no 64-bit division on the scalar (bs = 1) path */
    const int     c = BS1 ? 0 : (int)(e - i * bs);
    const int64_t s = (int64_t)(sidx ? sidx[i] : s0 + (int)i) * bs + c;
    const int64_t t = (int64_t)(didx ? didx[i] : d0 + (int)i) * bs + c;
    const T       u = src[s];
    if (OP == B200_SF_REPLACE) dst[t] = u;
    else dst[t] = sf_apply<T, OP>(dst[t], u);
  }
}

/* one thread per (destination group, component): the entries of a group are applied in entry order */
template <typename T, int OP, bool BS1> __global__ void __launch_bounds__(256) sf_scatter_grouped_kernel(int64_t ng, int bs, const int *__restrict__ gdst, const int *__restrict__ goff, const int *__restrict__ gsrc, const T *__restrict__ src, T *dst)
{
  const int64_t total = ng * bs, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t g = BS1 ? e : e / bs;
    const int     c = BS1 ? 0 : (int)(e - g * bs);
    const int64_t t = (int64_t)gdst[g] * bs + c;
    T             v = dst[t];
    for (int k = goff[g]; k < goff[g + 1]; k++) v = sf_apply<T, OP>(v, src[(int64_t)gsrc[k] * bs + c]);
    dst[t] = v;
  }
}

/* host part of the plan (no device needed): extents, contiguity, and -- when some destination is hit more than once -- the stable
   grouping by destination.  gdst/goff/gsrc are malloc'ed (b200HostFree) and NULL when *grouped == 0. */
extern "C" int b200IndexedGroupHost(int64_t n, const int *sidx, int s0, const int *didx, int d0, int *src_contig, int *dst_contig, int64_t *src_extent, int64_t *dst_extent, int *grouped, int64_t *ngroups, int **gdst_out, int **goff_out, int **gsrc_out)
{
  B200_CHECK(src_contig && dst_contig && src_extent && dst_extent && grouped && ngroups && gdst_out && goff_out && gsrc_out, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 0 && n < 2147483647LL, B200_ERR_ARG_OUTOFRANGE, "entry count %lld out of range", (long long)n);
  *gdst_out = *goff_out = *gsrc_out = NULL;
  *grouped = 0;
  int     sc = 1, dc = 1;
  int64_t smax = -1, dmax = -1;
  for (int64_t i = 0; i < n; i++) {
    const int s = sidx ? sidx[i] : s0 + (int)i, d = didx ? didx[i] : d0 + (int)i;
    B200_CHECK(s >= 0 && d >= 0, B200_ERR_ARG_OUTOFRANGE, "negative index at entry %lld", (long long)i);
    if (sidx && s != sidx[0] + (int)i) sc = 0;
    if (didx && d != didx[0] + (int)i) dc = 0;
    if (s > smax) smax = s;
    if (d > dmax) dmax = d;
  }
  *src_contig = sc;
  *dst_contig = dc;
  *src_extent = smax + 1;
  *dst_extent = dmax + 1;
  *ngroups    = n;
  if (dc || !n) return 0; /* contiguous destinations are distinct */
  int *cnt = (int *)calloc((size_t)dmax + 2, sizeof(int));
  B200_CHECK(cnt, B200_ERR_MEM, "out of host memory");
  for (int64_t i = 0; i < n; i++)
    if (++cnt[didx[i] + 1] > 1) *grouped = 1;
  if (!*grouped) {
    free(cnt);
    return 0;
  }
  /* stable counting sort by destination: groups in increasing destination, entries of a group in entry order */
  int64_t ng = 0;
  for (int64_t d = 0; d <= dmax; d++) ng += (cnt[d + 1] > 0);
  int *gdst = (int *)malloc(sizeof(int) * (size_t)ng), *goff = (int *)malloc(sizeof(int) * ((size_t)ng + 1)), *gsrc = (int *)malloc(sizeof(int) * (size_t)n);
  int *pos = (int *)malloc(sizeof(int) * ((size_t)dmax + 1)); /* next free slot of destination d */
  if (!gdst || !goff || !gsrc || !pos) {
    free(gdst); free(goff); free(gsrc); free(pos); free(cnt);
    B200_CHECK(0, B200_ERR_MEM, "out of host memory");
  }
  int64_t g = 0;
  int     off = 0;
  for (int64_t d = 0; d <= dmax; d++) {
    pos[d] = off;
    if (cnt[d + 1] > 0) {
      gdst[g] = (int)d;
      goff[g] = off;
      g++;
      off += cnt[d + 1];
    }
  }
  goff[ng] = off;
  for (int64_t i = 0; i < n; i++) gsrc[pos[didx[i]]++] = sidx ? sidx[i] : s0 + (int)i;
  free(pos);
  free(cnt);
  *ngroups  = ng;
  *gdst_out = gdst;
  *goff_out = goff;
  *gsrc_out = gsrc;
  return 0;
}

static int sf_upload(b200Handle h, int **d, const int *src, size_t count)
{
  int rc = b200Malloc(h, (void **)d, sizeof(int) * count);
  return rc ? rc : b200MemcpyHtoD(h, *d, src, sizeof(int) * count);
}

extern "C" int b200IndexedPlanCreate(b200Handle h, int64_t n, const int *sidx, int s0, const int *didx, int d0, b200IndexedPlan *plan_out)
{
  B200_CHECK(h && plan_out, B200_ERR_ARG_NULL, "null argument");
  int *gdst = NULL, *goff = NULL, *gsrc = NULL;
  b200IndexedPlan p = (b200IndexedPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  int rc = b200IndexedGroupHost(n, sidx, s0, didx, d0, &p->src_contig, &p->dst_contig, &p->src_extent, &p->dst_extent, &p->grouped, &p->ngroups, &gdst, &goff, &gsrc);
  if (rc) {
    free(p);
    return rc;
  }
  p->n  = n;
  p->s0 = (sidx && n) ? sidx[0] : s0;
  p->d0 = (didx && n) ? didx[0] : d0;
  p->k_src_contig = p->src_contig;
  p->k_dst_contig = p->dst_contig;
  p->k_s0         = p->s0;
  p->k_d0         = p->d0;
  if (!p->grouped && !p->dst_contig && n) {
    /* distinct, non-contiguous destinations: the order of the entries does not matter, so give the kernel "pull" order -- entries
       sorted by destination.  Consecutive threads then WRITE consecutive (or increasing) addresses and gather their sources;
       scattered 8-byte read-modify-writes measured 5x slower than gathers of the same index set (profiles/round2_notes.md).
       A permutation of a range (the usual VecScatter "to" set) needs no destination indices at all after the sort. */
    const int64_t ext = p->dst_extent;
    int *inv = (int *)malloc(sizeof(int) * (size_t)ext), *sd = (int *)malloc(sizeof(int) * (size_t)n), *ss = (int *)malloc(sizeof(int) * (size_t)n);
    if (!inv || !sd || !ss) {
      free(inv); free(sd); free(ss); free(p);
      B200_CHECK(0, B200_ERR_MEM, "out of host memory");
    }
    for (int64_t d = 0; d < ext; d++) inv[d] = -1;
    for (int64_t i = 0; i < n; i++) inv[didx[i]] = (int)i;
    int64_t k = 0;
    int     sc = 1, dc = 1;
    for (int64_t d = 0; d < ext; d++)
      if (inv[d] >= 0) {
        sd[k] = (int)d;
        ss[k] = sidx ? sidx[inv[d]] : s0 + inv[d];
        if (sd[k] != sd[0] + (int)k) dc = 0;
        if (ss[k] != ss[0] + (int)k) sc = 0;
        k++;
      }
    p->k_src_contig = sc;
    p->k_dst_contig = dc;
    p->k_s0         = ss[0];
    p->k_d0         = sd[0];
    if (!sc) rc = sf_upload(h, &p->d_sidx, ss, (size_t)n);
    if (!rc && !dc) rc = sf_upload(h, &p->d_didx, sd, (size_t)n);
    free(inv); free(sd); free(ss);
  } else if (!p->grouped) {
    if (!p->src_contig && n) rc = sf_upload(h, &p->d_sidx, sidx, (size_t)n);
  } else {
    rc = sf_upload(h, &p->d_gdst, gdst, (size_t)p->ngroups);
    if (!rc) rc = sf_upload(h, &p->d_goff, goff, (size_t)p->ngroups + 1);
    if (!rc) rc = sf_upload(h, &p->d_gsrc, gsrc, (size_t)n);
  }
  free(gdst);
  free(goff);
  free(gsrc);
  if (rc) {
    b200IndexedPlanDestroy(h, p);
    return rc;
  }
  *plan_out = p;
  return 0;
}

extern "C" int b200IndexedPlanDestroy(b200Handle h, b200IndexedPlan p)
{
  if (!p) return 0;
  int rc = 0, r;
  if (p->d_sidx && (r = b200Free(h, p->d_sidx))) rc = r;
  if (p->d_didx && (r = b200Free(h, p->d_didx))) rc = r;
  if (p->d_gdst && (r = b200Free(h, p->d_gdst))) rc = r;
  if (p->d_goff && (r = b200Free(h, p->d_goff))) rc = r;
  if (p->d_gsrc && (r = b200Free(h, p->d_gsrc))) rc = r;
  free(p);
  return rc;
}

extern "C" int b200IndexedPlanGetInfo(b200IndexedPlan p, int64_t *n, int64_t *ngroups, int *grouped, int *src_contig, int *dst_contig, int64_t *src_extent, int64_t *dst_extent)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null plan");
  if (n) *n = p->n;
  if (ngroups) *ngroups = p->ngroups;
  if (grouped) *grouped = p->grouped;
  if (src_contig) *src_contig = p->src_contig;
  if (dst_contig) *dst_contig = p->dst_contig;
  if (src_extent) *src_extent = p->src_extent;
  if (dst_extent) *dst_extent = p->dst_extent;
  return 0;
}

template <typename T, int OP> static int sf_launch(b200Handle h, b200IndexedPlan p, int bs, const T *src, T *dst)
{
  const int64_t units = (p->grouped ? p->ngroups : p->n) * bs;
  if (!units) return 0;
  int64_t g = (units + 255) / 256;
  if (g > (int64_t)h->num_sms * 16) g = (int64_t)h->num_sms * 16;
  const int *si = p->k_src_contig ? NULL : p->d_sidx, *di = p->k_dst_contig ? NULL : p->d_didx;
  if (p->grouped && bs == 1) sf_scatter_grouped_kernel<T, OP, true><<<(int)g, 256, 0, h->stream>>>(p->ngroups, bs, p->d_gdst, p->d_goff, p->d_gsrc, src, dst);
  else if (p->grouped) sf_scatter_grouped_kernel<T, OP, false><<<(int)g, 256, 0, h->stream>>>(p->ngroups, bs, p->d_gdst, p->d_goff, p->d_gsrc, src, dst);
  else if (bs == 1) sf_scatter_kernel<T, OP, true><<<(int)g, 256, 0, h->stream>>>(p->n, bs, si, p->k_s0, di, p->k_d0, src, dst);
  else sf_scatter_kernel<T, OP, false><<<(int)g, 256, 0, h->stream>>>(p->n, bs, si, p->k_s0, di, p->k_d0, src, dst);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

template <typename T> static int sf_dispatch(b200Handle h, b200IndexedPlan p, int bs, int op, const T *src, T *dst)
{
  switch (op) {
  case B200_SF_REPLACE: return sf_launch<T, B200_SF_REPLACE>(h, p, bs, src, dst);
  case B200_SF_SUM: return sf_launch<T, B200_SF_SUM>(h, p, bs, src, dst);
  case B200_SF_PROD: return sf_launch<T, B200_SF_PROD>(h, p, bs, src, dst);
  case B200_SF_MAX: return sf_launch<T, B200_SF_MAX>(h, p, bs, src, dst);
  case B200_SF_MIN: return sf_launch<T, B200_SF_MIN>(h, p, bs, src, dst);
  }
  B200_CHECK(0, B200_ERR_SUP, "operation %d is not one of REPLACE/SUM/PROD/MAX/MIN", op);
}

extern "C" int b200IndexedOp(b200Handle h, b200IndexedPlan p, int dtype, int bs, int op, const void *d_src, void *d_dst)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(bs >= 1, B200_ERR_ARG_OUTOFRANGE, "block size %d", bs);
  if (!p->n) return 0;
  B200_CHECK(d_src && d_dst, B200_ERR_ARG_NULL, "null data pointer");
  B200_CHECK(d_src != d_dst, B200_ERR_SUP, "in-place indexed operation (source == destination) has sequential semantics: not supported on the device");
  if (dtype == B200_SF_F64) return sf_dispatch<double>(h, p, bs, op, (const double *)d_src, (double *)d_dst);
  if (dtype == B200_SF_I32) return sf_dispatch<int>(h, p, bs, op, (const int *)d_src, (int *)d_dst);
  B200_CHECK(0, B200_ERR_SUP, "data type %d is not B200_SF_F64 / B200_SF_I32", dtype);
}
