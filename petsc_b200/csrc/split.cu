/*
 * split.cu -- device-side column split of a row block into the MATMPIAIJ diagonal / off-diagonal blocks.
 *
 * Setup-time helper (not on the Krylov hot path): MatSetUpMultiply_MPIAIJ's first step (mmaij.c:25-61) separates the
 * entries whose column is owned locally (block A, columns renumbered to local) from the rest (block B).  For matrices
 * that were generated or assembled on the device this avoids a round trip of the whole CSR through the host: only the
 * (tiny) B block travels to the host, where garray is sorted/uniqued exactly as the reference does.
 * The row-pointer exclusive scan uses cub::DeviceScan (CUDA toolkit header library).
 */
#include "b200_internal.h"
#include <cub/device/device_scan.cuh>

__global__ void split_count_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, int cstart, int cend, int *cntA, int *cntB)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= m; r += stride) {
    int a = 0, b = 0;
    if (r < m)
      for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
        int c = colidx[k];
        if (c >= cstart && c < cend) a++;
        else b++;
      }
    cntA[r] = a; /* entry m is 0: the exclusive scan over m+1 items leaves the totals in position m */
    cntB[r] = b;
  }
}

__global__ void split_fill_kernel(int m, const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val, int cstart, int cend, const int *__restrict__ Ai, int *Aj, double *Aa, const int *__restrict__ Bi, int *Bj, double *Ba)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    int ka = Ai[r], kb = Bi[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
      int c = colidx[k];
      if (c >= cstart && c < cend) {
        Aj[ka]   = c - cstart;
        Aa[ka++] = val[k];
      } else {
        Bj[kb]   = c; /* still global: renumbered into garray on the host */
        Ba[kb++] = val[k];
      }
    }
  }
}

extern "C" int b200CsrSplitColumns(b200Handle h, int m, const int *d_i, const int *d_j, const double *d_a, int cstart, int cend, int **d_Ai, int **d_Aj, double **d_Aa, int64_t *nzA, int **d_Bi, int **d_Bj, double **d_Ba, int64_t *nzB)
{
  B200_CHECK(h && d_Ai && d_Aj && d_Aa && d_Bi && d_Bj && d_Ba && nzA && nzB, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(m >= 0, B200_ERR_ARG_OUTOFRANGE, "negative size");
  int *Ai = NULL, *Bi = NULL, *Aj = NULL, *Bj = NULL;
  double *Aa = NULL, *Ba = NULL;
  int rc;
  if ((rc = b200Malloc(h, (void **)&Ai, sizeof(int) * ((size_t)m + 1)))) return rc;
  if ((rc = b200Malloc(h, (void **)&Bi, sizeof(int) * ((size_t)m + 1)))) return rc;
  int64_t g = ((int64_t)m + 256) / 256;
  if (g > h->num_sms * 16) g = h->num_sms * 16;
  split_count_kernel<<<(int)g, 256, 0, h->stream>>>(m, d_i, d_j, cstart, cend, Ai, Bi);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  void  *tmp = NULL;
  size_t tmp_bytes = 0;
  B200_CUDA(cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, Ai, Ai, m + 1, h->stream));
  B200_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, Ai, Ai, m + 1, h->stream));
  B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, Bi, Bi, m + 1, h->stream));
  int tot[2] = {0, 0};
  B200_CUDA(cudaMemcpyAsync(&tot[0], Ai + m, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(&tot[1], Bi + m, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_CUDA(cudaFree(tmp));
  if ((rc = b200Malloc(h, (void **)&Aj, sizeof(int) * ((size_t)tot[0] + 1)))) return rc;
  if ((rc = b200Malloc(h, (void **)&Aa, sizeof(double) * ((size_t)tot[0] + 1)))) return rc;
  if ((rc = b200Malloc(h, (void **)&Bj, sizeof(int) * ((size_t)tot[1] + 1)))) return rc;
  if ((rc = b200Malloc(h, (void **)&Ba, sizeof(double) * ((size_t)tot[1] + 1)))) return rc;
  if (m) {
    split_fill_kernel<<<(int)g, 256, 0, h->stream>>>(m, d_i, d_j, d_a, cstart, cend, Ai, Aj, Aa, Bi, Bj, Ba);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
  }
  *d_Ai = Ai; *d_Aj = Aj; *d_Aa = Aa; *nzA = tot[0];
  *d_Bi = Bi; *d_Bj = Bj; *d_Ba = Ba; *nzB = tot[1];
  return 0;
}

/* ------------------------------------------------------------------ benchmark operator generators (SURVEY 8d inputs) */
/* 27-point n^3 operator of bench_kspsolve.c:115-303 (h = 1/(n-1); centre 44h/13, face -3h/13, edge -3h/26, corner -h/13),
   rows in lexicographic order, columns sorted */
__global__ void lap27_count_kernel(int n, int *cnt)
{
  const int64_t N = (int64_t)n * n * n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= N; r += stride) {
    int c = 0;
    if (r < N) {
      int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / ((int64_t)n * n));
      c = (1 + (x > 0) + (x < n - 1)) * (1 + (y > 0) + (y < n - 1)) * (1 + (z > 0) + (z < n - 1));
    }
    cnt[r] = c;
  }
}
__global__ void lap27_fill_kernel(int n, const int *__restrict__ rowptr, int *colidx, double *val)
{
  const int64_t N = (int64_t)n * n * n, n2 = (int64_t)n * n, stride = (int64_t)gridDim.x * blockDim.x;
  const double  h = 1.0 / (n - 1);
  const double  w[4] = {44.0 / 13 * h, -3.0 / 13 * h, -3.0 / 26 * h, -1.0 / 13 * h};
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += stride) {
    int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / n2), k = rowptr[r];
    for (int dz = -1; dz <= 1; dz++) {
      if (z + dz < 0 || z + dz >= n) continue;
      for (int dy = -1; dy <= 1; dy++) {
        if (y + dy < 0 || y + dy >= n) continue;
        for (int dx = -1; dx <= 1; dx++) {
          if (x + dx < 0 || x + dx >= n) continue;
          colidx[k] = (int)(r + dx + (int64_t)n * dy + n2 * dz);
          val[k++]  = w[(dx != 0) + (dy != 0) + (dz != 0)];
        }
      }
    }
  }
}
extern "C" int b200GenLaplace27Nnz(int n, int64_t *nnz)
{
  int64_t t = 3 * (int64_t)n - 2;
  *nnz      = t * t * t;
  return 0;
}
extern "C" int b200GenLaplace27(b200Handle h, int n, int *d_rowptr, int *d_colidx, double *d_val)
{
  B200_CHECK(h && d_rowptr && d_colidx && d_val, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n >= 2, B200_ERR_ARG_OUTOFRANGE, "n must be >= 2");
  const int64_t N = (int64_t)n * n * n, t = 3 * (int64_t)n - 2;
  B200_CHECK(t * t * t < 2147483647LL - 8, B200_ERR_SUP, "operator exceeds 32-bit PetscInt");
  int64_t g = (N + 256) / 256;
  if (g > h->num_sms * 16) g = h->num_sms * 16;
  lap27_count_kernel<<<(int)g, 256, 0, h->stream>>>(n, d_rowptr);
  B200_KERNEL_CHECK();
  void  *tmp = NULL;
  size_t tmp_bytes = 0;
  B200_CUDA(cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, d_rowptr, d_rowptr, (int)(N + 1), h->stream));
  B200_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_rowptr, d_rowptr, (int)(N + 1), h->stream));
  lap27_fill_kernel<<<(int)g, 256, 0, h->stream>>>(n, d_rowptr, d_colidx, d_val);
  B200_KERNEL_CHECK();
  B200_LAUNCHED(2);
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_CUDA(cudaFree(tmp));
  return 0;
}

/* random CSR with a fixed row length d (config 5): column k of a row is drawn from the k-th of d equal strata of
   [0, ncols) -- sorted and distinct by construction; values in (-1, 1); both from a counter-based hash of (seed, row, k) */
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void random_csr_kernel(int n, int ncols, int d, uint64_t seed, int *rowptr, int *colidx, double *val)
{
  const int64_t total = (int64_t)n * d, stride = (int64_t)gridDim.x * blockDim.x;
  const int     width = ncols / d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t r = e / d;
    const int     k = (int)(e % d);
    uint64_t      u = mix64(seed ^ mix64((uint64_t)e));
    colidx[e]       = k * width + (int)(u % (uint64_t)width);
    val[e]          = ((double)(mix64(u) >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
    if (k == 0) rowptr[r] = (int)(r * d);
    if (e == total - 1) rowptr[n] = (int)total;
  }
}
extern "C" int b200GenRandomCsr(b200Handle h, int n, int ncols, int d, uint64_t seed, int *d_rowptr, int *d_colidx, double *d_val)
{
  B200_CHECK(h && d_rowptr && d_colidx && d_val, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(n > 0 && d > 0 && ncols >= d, B200_ERR_ARG_OUTOFRANGE, "need n > 0 and ncols >= d > 0");
  B200_CHECK((int64_t)n * d < 2147483647LL - 8, B200_ERR_SUP, "n*d exceeds 32-bit PetscInt");
  random_csr_kernel<<<h->num_sms * 16, 256, 0, h->stream>>>(n, ncols, d, seed, d_rowptr, d_colidx, d_val);
  B200_LAUNCHED(1);
  B200_KERNEL_CHECK();
  return 0;
}
