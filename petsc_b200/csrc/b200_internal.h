/* Internal declarations shared by the .cu translation units of libpetscb200.so. */
#ifndef B200_INTERNAL_H
#define B200_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/petscb200.h"

#define B200_MAX_NV      32   /* vectors per MDot/MAXPY launch; larger nv is chunked */
#define B200_RED_MAXGRID 1184 /* 148 SMs x 8 */
#define B200_ALLOC_PAD   256

struct b200Handle_s {
  int          device;
  int          num_sms;
  int          l2_persist_max;  /* cudaDevAttrMaxPersistingL2CacheSize (bytes) */
  int          l2_window_max;   /* cudaDevAttrMaxAccessPolicyWindowSize (bytes) */
  int          l2_persist_set;  /* cudaLimitPersistingL2CacheSize raised by this handle */
  cudaStream_t stream;      /* current stream */
  cudaStream_t own_stream;  /* created with the handle */
  cudaStream_t halo_stream; /* second stream for the NCCL halo exchange */
  cudaEvent_t  ev_main, ev_halo;
  /* reduction workspace */
  double       *d_partials; /* [B200_MAX_NV+1][B200_RED_MAXGRID] */
  unsigned int *d_counter;
  double       *d_result;   /* [B200_MAX_NV+2] device copy of results */
  double       *h_result;   /* pinned + mapped: final stage writes here directly */
  double       *h_result_dev; /* device alias of h_result */
  int          *d_flag, *h_flag; /* small int results (zero counts, indices) */
  long long    *d_idx;
  /* NCCL */
  void *nccl_comm;
  int   rank, nranks;
};

void b200_set_error(int code, const char *fmt, ...);
extern long long g_b200_launches;

#define B200_CUDA(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      b200_set_error(B200_ERR_GPU, "cuda error %d (%s) : %s at %s:%d", (int)e_, cudaGetErrorName(e_), cudaGetErrorString(e_), __FILE__, __LINE__); \
      return B200_ERR_GPU; \
    } \
  } while (0)

#define B200_CHECK(cond, code, ...) \
  do { \
    if (!(cond)) { \
      b200_set_error(code, __VA_ARGS__); \
      return code; \
    } \
  } while (0)

#define B200_LAUNCHED(n) (g_b200_launches += (n))
#define B200_KERNEL_CHECK() B200_CUDA(cudaPeekAtLastError())


/* segment-marching triangular sweeps (ilu.cu), shared by ILU(0) and ICC(0): see sweep_march_kernel for the modes */
#include <vector_types.h>
int2 *b200_build_segments(int n, int mode, const int *ext, const int *bj, int rpw, int minlen, int maxlen, int *nslot_out, int *nlev_out);
int   b200_sweep_march(b200Handle h, int G, int mode, int nslot, const int2 *segs, const int *ext, const int *bj, const double *ba, const double *rhs, double *out, int64_t nnz, const double *dinv, double *out2);
#endif
