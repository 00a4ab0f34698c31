/*
 * assemble.cu -- device-side COO assembly and the transposed product.
 *
 *   b200CooPlan*          MatSetPreallocationCOO_SeqAIJ / MatSetValuesCOO_SeqAIJ (reference: src/mat/impls/aij/seq/aij.c:4524-4732;
 *                         device version of the value pass: aijcusparse.cu MatSetValuesCOO_SeqAIJCUSPARSE).  The reference sorts the
 *                         (i,j) pairs on the HOST even in its device back ends; here the sort, the unique pass and the CSR pattern are
 *                         built on the device from device-resident index arrays (64-bit key radix sort, cub), so a matrix that
 *                         is generated on the GPU never visits the host.
 *                         Summation order of repeated (i,j) pairs: the device plan adds them in their order in the user's
 *                         array (stable sort) -- the order MatSetValues(ADD_VALUES) would give.  The reference's order is an
 *                         artefact of its unstable quicksort (sorti.c:198-240); with <= 2 repeats both give bit-identical
 *                         values, with more they differ by rounding.  b200CooPlanCreateFromMaps takes the reference's own
 *                         jmap/perm (MatCOOStruct_SeqAIJ, aij.h:170-176) and then reproduces its values bit for bit: that is what
 *                         the PETSc plugin uses by default.
 *   b200CsrTranspose*     MatMultTranspose_SeqAIJ / MatMultTransposeAdd_SeqAIJ (aij.c:1383-1440): y[c] accumulates x[i]*a[k] in
 *                         increasing row order.  An explicit transposed pattern (column-major order of A's entries, stable in
 *                         the row index) turns this into the row-ordered FMA-free sum the SpMV kernel already computes, so the
 *                         result is bit-identical to the reference with the SpMV plan's exact summation (one lane per row, or
 *                         b200CsrPlanSetSummation(plan, 0)) and equal to rounding at the SpMV roofline otherwise; the value permutation is
 *                         re-applied (one gather pass) only when A's values change.
 *
 * Setup-time helpers use cub (CUDA toolkit header library) for sort / scan.
 */
#include "b200_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>

static inline int asm_grid(b200Handle h, int64_t n)
{
  int64_t g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > (int64_t)h->num_sms * 16) g = (int64_t)h->num_sms * 16;
  return (int)g;
}
static inline int bits_for(uint64_t v) /* number of bits needed to represent v */
{
  int b = 0;
  while (v) { b++; v >>= 1; }
  return b ? b : 1;
}

/* =================================================================================================================== */
/* COO                                                                                                                 */
/* =================================================================================================================== */
struct b200CooPlan_s {
  int     M, N;
  int64_t coo_n, atot, nnz;
  int    *d_rowptr, *d_colidx; /* CSR pattern (NULL for plans made from maps) */
  int    *d_jmap;              /* [nnz+1] */
  int    *d_perm;              /* [atot]  */
};

/* key = row:col; dropped entries (negative row or column, aij.c:4547-4556) get row M so that they sort behind every valid one */
__global__ void coo_keys_kernel(int64_t n, int M, int N, const int *__restrict__ ci, const int *__restrict__ cj, unsigned long long *keys, int *idx, int *bad)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int i = ci[k], j = cj[k];
    unsigned long long key;
    if (i < 0 || j < 0) key = (unsigned long long)(unsigned)M << 32;
    else {
      if (i >= M) atomicOr(bad, 1);
      if (j >= N) atomicOr(bad, 2);
      key = ((unsigned long long)(unsigned)i << 32) | (unsigned)j;
    }
    keys[k] = key;
    idx[k]  = (int)k;
  }
}

/* head[k] = 1 where a new (row,col) pair starts among the valid entries; atot = number of valid entries */
__global__ void coo_heads_kernel(int64_t n, int M, const unsigned long long *__restrict__ keys, int *head, long long *atot)
{
  const int64_t            stride  = (int64_t)gridDim.x * blockDim.x;
  const unsigned long long invalid = (unsigned long long)(unsigned)M << 32;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += stride) {
    if (k == n) { head[k] = 0; continue; } /* the scan over n+1 items leaves nnz in position n */
    const unsigned long long key   = keys[k];
    const bool               valid = key < invalid;
    head[k] = valid && (k == 0 || keys[k - 1] != key);
    if (valid && (k == n - 1 || keys[k + 1] >= invalid)) *atot = k + 1;
  }
}

/* for every head: column, jmap, and the row pointers of the rows that start at (or are empty before) this nonzero */
__global__ void coo_fill_kernel(int64_t n, int M, const unsigned long long *__restrict__ keys, const int *__restrict__ head, const int *__restrict__ q_of, int *rowptr, int *colidx, int *jmap)
{
  const int64_t            stride  = (int64_t)gridDim.x * blockDim.x;
  const unsigned long long invalid = (unsigned long long)(unsigned)M << 32;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    if (!head[k]) continue;
    const unsigned long long key = keys[k];
    const int                q = q_of[k], row = (int)(key >> 32);
    colidx[q] = (int)(key & 0xffffffffu);
    jmap[q]   = (int)k;
    const int rprev = k ? (int)(keys[k - 1] >> 32) : -1; /* row of the previous nonzero (same row: nothing to do) */
    for (int r = rprev + 1; r <= row; r++) rowptr[r] = q;
    /* last nonzero: close the remaining rows */
    if (k == n - 1 || keys[k + 1] >= invalid) {
      /* handled by the tail kernel (needs nnz) */
    }
  }
}
__global__ void coo_tail_kernel(int M, int last_row, int nnz, int atot, int *rowptr, int *jmap)
{
  const int stride = gridDim.x * blockDim.x;
  for (int r = last_row + 1 + blockIdx.x * blockDim.x + threadIdx.x; r <= M; r += stride) rowptr[r] = nnz;
  if (blockIdx.x == 0 && threadIdx.x == 0) jmap[nnz] = atot;
}

/* MatSetValuesCOO_SeqAIJ (aij.c:4724-4728): one thread per unique nonzero, repeats added left to right starting from 0.0 */
__global__ void coo_setvalues_kernel(int64_t nnz, const int *__restrict__ jmap, const int *__restrict__ perm, const double *__restrict__ v, int insert, double *a)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
    const int ks = jmap[q], ke = jmap[q + 1];
    double    sum = 0.0;
    for (int k = ks; k < ke; k++) sum = __dadd_rn(sum, v[perm[k]]);
    a[q] = __dadd_rn(insert ? 0.0 : a[q], sum);
  }
}

extern "C" int b200CooPlanDestroy(b200CooPlan p)
{
  if (!p) return 0;
  cudaFree(p->d_rowptr);
  cudaFree(p->d_colidx);
  cudaFree(p->d_jmap);
  cudaFree(p->d_perm);
  free(p);
  return 0;
}

extern "C" int b200CooPlanCreate(b200Handle h, int M, int N, int64_t coo_n, const int *d_coo_i, const int *d_coo_j, b200CooPlan *plan)
{
  B200_CHECK(h && plan, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(M >= 0 && N >= 0 && coo_n >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  B200_CHECK(coo_n <= 2147483647LL - 8, B200_ERR_SUP, "COO count %lld needs 64-bit permutation indices (not built)", (long long)coo_n);
  B200_CHECK(!coo_n || (d_coo_i && d_coo_j), B200_ERR_ARG_NULL, "null COO index arrays");
  b200CooPlan p = (b200CooPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->M = M; p->N = N; p->coo_n = coo_n;
  unsigned long long *keys = NULL, *keys2 = NULL;
  int                *idx = NULL, *idx2 = NULL, *head = NULL, *qof = NULL, *bad = NULL;
  long long          *atot = NULL;
  void               *tmp = NULL;
  int                 rc = 0;
#define COO_CUDA(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      b200_set_error(B200_ERR_GPU, "cuda error %d (%s) : %s at %s:%d", (int)e_, cudaGetErrorName(e_), cudaGetErrorString(e_), __FILE__, __LINE__); \
      rc = B200_ERR_GPU; \
      goto done; \
    } \
  } while (0)
  {
    const size_t n1 = (size_t)coo_n + 1;
    COO_CUDA(cudaMalloc(&keys, sizeof(*keys) * n1));
    COO_CUDA(cudaMalloc(&keys2, sizeof(*keys2) * n1));
    COO_CUDA(cudaMalloc(&idx, sizeof(int) * n1));
    COO_CUDA(cudaMalloc(&idx2, sizeof(int) * n1));
    COO_CUDA(cudaMalloc(&bad, sizeof(int) * 4));
    COO_CUDA(cudaMalloc(&atot, sizeof(long long)));
    COO_CUDA(cudaMemsetAsync(bad, 0, sizeof(int) * 4, h->stream));
    COO_CUDA(cudaMemsetAsync(atot, 0, sizeof(long long), h->stream));
    COO_CUDA(cudaMalloc(&p->d_rowptr, sizeof(int) * ((size_t)M + 1)));
    if (coo_n) {
      coo_keys_kernel<<<asm_grid(h, coo_n), 256, 0, h->stream>>>(coo_n, M, N, d_coo_i, d_coo_j, keys, idx, bad);
      B200_LAUNCHED(1);
      COO_CUDA(cudaPeekAtLastError());
      int hbad = 0;
      COO_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      COO_CUDA(cudaStreamSynchronize(h->stream));
      if (hbad) {
        /* aij.c:4561 / :4631 */
        b200_set_error(B200_ERR_ARG_OUTOFRANGE, "COO %s index is >= the matrix %s size", (hbad & 1) ? "row" : "column", (hbad & 1) ? "row" : "column");
        rc = B200_ERR_ARG_OUTOFRANGE;
        goto done;
      }
      /* stable LSD radix sort on the significant bits of row:col only */
      size_t    tmp_bytes = 0;
      const int end_bit   = 32 + bits_for((uint64_t)M);
      COO_CUDA(cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, keys, keys2, idx, idx2, (int)coo_n, 0, end_bit, h->stream));
      COO_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
      COO_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, idx2, (int)coo_n, 0, end_bit, h->stream));
      B200_LAUNCHED(1);
      COO_CUDA(cudaFree(tmp));
      tmp = NULL;
      COO_CUDA(cudaFree(keys));
      keys = NULL;
      COO_CUDA(cudaFree(idx));
      idx = NULL;
    }
    COO_CUDA(cudaMalloc(&head, sizeof(int) * n1));
    COO_CUDA(cudaMalloc(&qof, sizeof(int) * n1));
    coo_heads_kernel<<<asm_grid(h, coo_n + 1), 256, 0, h->stream>>>(coo_n, M, keys2, head, atot);
    B200_LAUNCHED(1);
    COO_CUDA(cudaPeekAtLastError());
    size_t tmp_bytes = 0;
    COO_CUDA(cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, head, qof, (int)n1, h->stream));
    COO_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    COO_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, head, qof, (int)n1, h->stream));
    B200_LAUNCHED(1);
    int       hnnz = 0;
    long long hatot = 0;
    COO_CUDA(cudaMemcpyAsync(&hnnz, qof + coo_n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    COO_CUDA(cudaMemcpyAsync(&hatot, atot, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    COO_CUDA(cudaStreamSynchronize(h->stream));
    p->nnz  = hnnz;
    p->atot = hatot;
    COO_CUDA(cudaMalloc(&p->d_colidx, sizeof(int) * ((size_t)hnnz + 1)));
    COO_CUDA(cudaMalloc(&p->d_jmap, sizeof(int) * ((size_t)hnnz + 1)));
    int last_row = -1;
    if (hatot) {
      unsigned long long lastkey = 0;
      COO_CUDA(cudaMemcpyAsync(&lastkey, keys2 + (hatot - 1), sizeof(lastkey), cudaMemcpyDeviceToHost, h->stream));
      COO_CUDA(cudaStreamSynchronize(h->stream));
      last_row = (int)(lastkey >> 32);
      coo_fill_kernel<<<asm_grid(h, coo_n), 256, 0, h->stream>>>(coo_n, M, keys2, head, qof, p->d_rowptr, p->d_colidx, p->d_jmap);
      B200_LAUNCHED(1);
      COO_CUDA(cudaPeekAtLastError());
    }
    coo_tail_kernel<<<asm_grid(h, M + 1), 256, 0, h->stream>>>(M, last_row, hnnz, (int)hatot, p->d_rowptr, p->d_jmap);
    B200_LAUNCHED(1);
    COO_CUDA(cudaPeekAtLastError());
    /* perm = the sorted original positions of the valid entries */
    COO_CUDA(cudaMalloc(&p->d_perm, sizeof(int) * ((size_t)hatot + 1)));
    if (hatot) COO_CUDA(cudaMemcpyAsync(p->d_perm, idx2, sizeof(int) * (size_t)hatot, cudaMemcpyDeviceToDevice, h->stream));
    COO_CUDA(cudaStreamSynchronize(h->stream));
  }
done:
  cudaFree(keys); cudaFree(keys2); cudaFree(idx); cudaFree(idx2); cudaFree(head); cudaFree(qof); cudaFree(bad); cudaFree(atot); cudaFree(tmp);
  if (rc) {
    b200CooPlanDestroy(p);
    return rc;
  }
  *plan = p;
  return 0;
#undef COO_CUDA
}

extern "C" int b200CooPlanCreateFromMaps(b200Handle h, int64_t nnz, int64_t atot, const int64_t *h_jmap, const int64_t *h_perm, b200CooPlan *plan)
{
  B200_CHECK(h && plan && h_jmap && (h_perm || !atot), B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(nnz >= 0 && atot >= 0 && atot <= 2147483647LL - 8 && nnz <= atot, B200_ERR_ARG_OUTOFRANGE, "bad COO map sizes");
  b200CooPlan p = (b200CooPlan)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->M = p->N = -1; p->coo_n = atot; p->atot = atot; p->nnz = nnz;
  int *t = (int *)malloc(sizeof(int) * (size_t)((nnz + 1 > atot ? nnz + 1 : atot) + 1));
  if (!t) { free(p); B200_CHECK(0, B200_ERR_MEM, "out of host memory"); }
  cudaError_t e = cudaMalloc(&p->d_jmap, sizeof(int) * ((size_t)nnz + 1));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_perm, sizeof(int) * ((size_t)atot + 1));
  if (e == cudaSuccess) {
    for (int64_t q = 0; q <= nnz; q++) t[q] = (int)h_jmap[q];
    e = cudaMemcpy(p->d_jmap, t, sizeof(int) * ((size_t)nnz + 1), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && atot) {
    for (int64_t k = 0; k < atot; k++) t[k] = (int)h_perm[k];
    e = cudaMemcpy(p->d_perm, t, sizeof(int) * (size_t)atot, cudaMemcpyHostToDevice);
  }
  free(t);
  if (e != cudaSuccess) {
    b200CooPlanDestroy(p);
    b200_set_error(B200_ERR_GPU, "cuda error %d (%s) while copying the COO maps", (int)e, cudaGetErrorName(e));
    return B200_ERR_GPU;
  }
  *plan = p;
  return 0;
}

extern "C" int b200CooPlanGetCsr(b200CooPlan p, int64_t *nnz, int64_t *atot, const int **d_rowptr, const int **d_colidx)
{
  B200_CHECK(p, B200_ERR_ARG_NULL, "null argument");
  if (nnz) *nnz = p->nnz;
  if (atot) *atot = p->atot;
  if (d_rowptr) *d_rowptr = p->d_rowptr;
  if (d_colidx) *d_colidx = p->d_colidx;
  return 0;
}

extern "C" int b200CooPlanGetMaps(b200Handle h, b200CooPlan p, int *h_jmap, int *h_perm)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (h_jmap) B200_CUDA(cudaMemcpyAsync(h_jmap, p->d_jmap, sizeof(int) * ((size_t)p->nnz + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h_perm && p->atot) B200_CUDA(cudaMemcpyAsync(h_perm, p->d_perm, sizeof(int) * (size_t)p->atot, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int b200CooSetValues(b200Handle h, b200CooPlan p, const double *d_v, int insert, double *d_a)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->nnz) return 0;
  B200_CHECK(d_a && (d_v || !p->atot), B200_ERR_ARG_NULL, "null value array");
  coo_setvalues_kernel<<<asm_grid(h, p->nnz), 256, 0, h->stream>>>(p->nnz, p->d_jmap, p->d_perm, d_v, insert, d_a);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* =================================================================================================================== */
/* transposed product                                                                                                  */
/* =================================================================================================================== */
struct b200CsrTranspose_s {
  int         m, n; /* of A; the transposed pattern has n rows */
  int64_t     nnz;
  int        *d_tptr, *d_trow, *d_tperm;
  double     *d_at;
  b200CsrPlan plan; /* SpMV plan on the transposed pattern */
};

__global__ void tr_expand_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, unsigned int *keys, int *pos, int *rowof)
{
  /* one warp per row: keys = column, pos = position in A's arrays, rowof[position] = row */
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int     lane = threadIdx.x & 31;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < m; r += nw) {
    const int ks = rowptr[r], ke = rowptr[r + 1];
    for (int k = ks + lane; k < ke; k += 32) {
      keys[k]  = (unsigned)colidx[k];
      pos[k]   = k;
      rowof[k] = (int)r;
    }
  }
}
/* after the stable sort by column: trow[k] = row of the k-th transposed entry; tptr from the column boundaries */
__global__ void tr_fill_kernel(int64_t nnz, int n, const unsigned int *__restrict__ keys, const int *__restrict__ tperm, const int *__restrict__ rowof, int *tptr, int *trow)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    trow[k] = rowof[tperm[k]];
    const int c = (int)keys[k], cprev = k ? (int)keys[k - 1] : -1;
    for (int cc = cprev + 1; cc <= c; cc++) tptr[cc] = (int)k;
    if (k == nnz - 1)
      for (int cc = c + 1; cc <= n; cc++) tptr[cc] = (int)nnz;
  }
}
__global__ void tr_zero_ptr_kernel(int n, int *tptr)
{
  const int stride = gridDim.x * blockDim.x;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= n; c += stride) tptr[c] = 0;
}
__global__ void tr_gather_kernel(int64_t nnz, const int *__restrict__ tperm, const double *__restrict__ a, double *at)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) at[k] = a[tperm[k]];
}

extern "C" int b200CsrTransposeDestroy(b200CsrTranspose T)
{
  if (!T) return 0;
  if (T->plan) b200CsrPlanDestroy(T->plan);
  cudaFree(T->d_tptr);
  cudaFree(T->d_trow);
  cudaFree(T->d_tperm);
  cudaFree(T->d_at);
  free(T);
  return 0;
}

extern "C" int b200CsrTransposeCreate(b200Handle h, int m, int n, int64_t nnz, const int *d_rowptr, const int *d_colidx, b200CsrTranspose *out)
{
  B200_CHECK(h && out, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(m >= 0 && n >= 0 && nnz >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  B200_CHECK(nnz <= 2147483647LL - 8, B200_ERR_SUP, "nnz %lld does not fit 32-bit PetscInt row pointers", (long long)nnz);
  B200_CHECK(!m || (d_rowptr && (nnz == 0 || d_colidx)), B200_ERR_ARG_NULL, "null CSR arrays");
  b200CsrTranspose T = (b200CsrTranspose)calloc(1, sizeof(*T));
  B200_CHECK(T, B200_ERR_MEM, "out of host memory");
  T->m = m; T->n = n; T->nnz = nnz;
  unsigned int *keys = NULL, *keys2 = NULL;
  int          *pos = NULL, *rowof = NULL;
  void         *tmp = NULL;
  int           rc = 0;
#define TR_CUDA(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      b200_set_error(B200_ERR_GPU, "cuda error %d (%s) : %s at %s:%d", (int)e_, cudaGetErrorName(e_), cudaGetErrorString(e_), __FILE__, __LINE__); \
      rc = B200_ERR_GPU; \
      goto done; \
    } \
  } while (0)
  {
    const size_t z1 = (size_t)nnz + 1;
    TR_CUDA(cudaMalloc(&T->d_tptr, sizeof(int) * ((size_t)n + 1)));
    TR_CUDA(cudaMalloc(&T->d_trow, sizeof(int) * z1));
    TR_CUDA(cudaMalloc(&T->d_tperm, sizeof(int) * z1));
    TR_CUDA(cudaMalloc(&T->d_at, sizeof(double) * z1));
    tr_zero_ptr_kernel<<<asm_grid(h, n + 1), 256, 0, h->stream>>>(n, T->d_tptr);
    B200_LAUNCHED(1);
    TR_CUDA(cudaPeekAtLastError());
    if (nnz) {
      TR_CUDA(cudaMalloc(&keys, sizeof(unsigned) * z1));
      TR_CUDA(cudaMalloc(&keys2, sizeof(unsigned) * z1));
      TR_CUDA(cudaMalloc(&pos, sizeof(int) * z1));
      TR_CUDA(cudaMalloc(&rowof, sizeof(int) * z1));
      tr_expand_kernel<<<asm_grid(h, (int64_t)m * 32), 256, 0, h->stream>>>(m, d_rowptr, d_colidx, keys, pos, rowof);
      B200_LAUNCHED(1);
      TR_CUDA(cudaPeekAtLastError());
      size_t    tmp_bytes = 0;
      const int end_bit   = bits_for((uint64_t)(n > 0 ? n - 1 : 0));
      TR_CUDA(cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, keys, keys2, pos, T->d_tperm, (int)nnz, 0, end_bit, h->stream));
      TR_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
      TR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, pos, T->d_tperm, (int)nnz, 0, end_bit, h->stream));
      B200_LAUNCHED(1);
      tr_fill_kernel<<<asm_grid(h, nnz), 256, 0, h->stream>>>(nnz, n, keys2, T->d_tperm, rowof, T->d_tptr, T->d_trow);
      B200_LAUNCHED(1);
      TR_CUDA(cudaPeekAtLastError());
    }
    TR_CUDA(cudaStreamSynchronize(h->stream));
    rc = b200CsrPlanCreate(h, n, m, nnz, T->d_tptr, T->d_trow, &T->plan);
  }
done:
  cudaFree(keys); cudaFree(keys2); cudaFree(pos); cudaFree(rowof); cudaFree(tmp);
  if (rc) {
    b200CsrTransposeDestroy(T);
    return rc;
  }
  *out = T;
  return 0;
#undef TR_CUDA
}

extern "C" int b200CsrTransposeSetValues(b200Handle h, b200CsrTranspose T, const double *d_a)
{
  B200_CHECK(h && T, B200_ERR_ARG_NULL, "null argument");
  if (!T->nnz) return 0;
  B200_CHECK(d_a, B200_ERR_ARG_NULL, "null value array");
  tr_gather_kernel<<<asm_grid(h, T->nnz), 256, 0, h->stream>>>(T->nnz, T->d_tperm, d_a, T->d_at);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}

/* y = A^T x  (d_z == NULL)  or  y = z + A^T x ; x has m entries, y and z have n */
extern "C" int b200CsrTransposeSpMV(b200Handle h, b200CsrTranspose T, const double *d_x, const double *d_z, double *d_y)
{
  B200_CHECK(h && T, B200_ERR_ARG_NULL, "null argument");
  if (d_z) return b200CsrSpMVAdd(h, T->plan, T->d_at, d_x, d_z, d_y);
  return b200CsrSpMV(h, T->plan, T->d_at, d_x, d_y);
}

/* the SpMV plan of the transposed pattern (n rows): b200CsrPlanSetLayout(plan, 1, ...) selects the bit-exact parity mode */
extern "C" int b200CsrTransposeGetPlan(b200CsrTranspose T, b200CsrPlan *plan)
{
  B200_CHECK(T && plan, B200_ERR_ARG_NULL, "null argument");
  *plan = T->plan;
  return 0;
}

extern "C" int b200CsrTransposeGet(b200Handle h, b200CsrTranspose T, int *h_tptr, int *h_trow, int *h_tperm)
{
  B200_CHECK(h && T, B200_ERR_ARG_NULL, "null argument");
  if (h_tptr) B200_CUDA(cudaMemcpyAsync(h_tptr, T->d_tptr, sizeof(int) * ((size_t)T->n + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h_trow && T->nnz) B200_CUDA(cudaMemcpyAsync(h_trow, T->d_trow, sizeof(int) * (size_t)T->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (h_tperm && T->nnz) B200_CUDA(cudaMemcpyAsync(h_tperm, T->d_tperm, sizeof(int) * (size_t)T->nnz, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
