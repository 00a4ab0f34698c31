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
