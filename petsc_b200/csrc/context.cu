/* Handle, streams, memory: the part of the C ABI that replaces PetscCUBLASGetHandle/PetscGetCurrentCUDAStream
   (include/petscdevice_cuda.h:180-183) and the cudaMalloc/cudaMemcpy calls of the reference's device mirrors. */
#include "b200_internal.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[1024] = "";
long long            g_b200_launches = 0;

void b200_set_error(int code, const char *fmt, ...)
{
  va_list ap;
  int     k = snprintf(g_err, sizeof g_err, "[petscb200 error %d] ", code);
  va_start(ap, fmt);
  vsnprintf(g_err + k, sizeof g_err - (size_t)k, fmt, ap);
  va_end(ap);
}

extern "C" const char *b200GetLastErrorString(void) { return g_err; }
extern "C" const char *b200Version(void) { return "petscb200 0.1 (sm_100a)"; }
extern "C" long long   b200KernelLaunchCount(void) { return g_b200_launches; }
/* bytes moved by b200Memcpy{HtoD,DtoH}[Async] since load: what a PETSc plugin reports through PetscLogCpuToGpu/GpuToCpu and
   what the tests use to prove that a KSPSolve moves no vector over PCIe per iteration */
static long long g_b200_h2d = 0, g_b200_d2h = 0;
extern "C" int b200TransferCounters(long long *h2d_bytes, long long *d2h_bytes)
{
  if (h2d_bytes) *h2d_bytes = g_b200_h2d;
  if (d2h_bytes) *d2h_bytes = g_b200_d2h;
  return 0;
}

extern "C" int b200DeviceCount(int *n)
{
  B200_CUDA(cudaGetDeviceCount(n));
  return 0;
}

extern "C" int b200Create(b200Handle *hp, int device)
{
  B200_CHECK(hp, B200_ERR_ARG_NULL, "null handle pointer");
  if (device < 0) B200_CUDA(cudaGetDevice(&device));
  else B200_CUDA(cudaSetDevice(device));
  b200Handle h = (b200Handle)calloc(1, sizeof(*h));
  B200_CHECK(h, B200_ERR_MEM, "out of host memory");
  h->device = device;
  h->nranks = 1;
  B200_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  B200_CUDA(cudaDeviceGetAttribute(&h->l2_persist_max, cudaDevAttrMaxPersistingL2CacheSize, device));
  B200_CUDA(cudaDeviceGetAttribute(&h->l2_window_max, cudaDevAttrMaxAccessPolicyWindowSize, device));
  B200_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&h->halo_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  B200_CUDA(cudaMalloc(&h->d_partials, sizeof(double) * (B200_MAX_NV + 2) * B200_RED_MAXGRID));
  B200_CUDA(cudaMalloc(&h->d_counter, 256));
  B200_CUDA(cudaMemset(h->d_counter, 0, 256));
  B200_CUDA(cudaMalloc(&h->d_result, sizeof(double) * (B200_MAX_NV + 8)));
  B200_CUDA(cudaMalloc(&h->d_flag, 256));
  B200_CUDA(cudaMalloc(&h->d_idx, 256));
  B200_CUDA(cudaHostAlloc(&h->h_result, sizeof(double) * (B200_MAX_NV + 8), cudaHostAllocMapped));
  B200_CUDA(cudaHostGetDevicePointer(&h->h_result_dev, h->h_result, 0));
  B200_CUDA(cudaHostAlloc(&h->h_flag, 256, cudaHostAllocDefault));
  *hp = h;
  return 0;
}

extern "C" int b200CommDestroy(b200Handle h);

extern "C" int b200Destroy(b200Handle h)
{
  if (!h) return 0;
  b200CommDestroy(h);
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_partials); cudaFree(h->d_counter); cudaFree(h->d_result); cudaFree(h->d_flag); cudaFree(h->d_idx);
  cudaFreeHost(h->h_result); cudaFreeHost(h->h_flag);
  cudaEventDestroy(h->ev_main); cudaEventDestroy(h->ev_halo);
  cudaStreamDestroy(h->own_stream); cudaStreamDestroy(h->halo_stream);
  free(h);
  return 0;
}

extern "C" int b200SetStream(b200Handle h, void *s)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return 0;
}
extern "C" int b200GetStream(b200Handle h, void **s)
{
  B200_CHECK(h && s, B200_ERR_ARG_NULL, "null argument");
  *s = (void *)h->stream;
  return 0;
}
extern "C" int b200Synchronize(b200Handle h)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int b200DeviceSynchronize(void)
{
  B200_CUDA(cudaDeviceSynchronize());
  return 0;
}
extern "C" int b200GetDevice(b200Handle h, int *d)
{
  B200_CHECK(h && d, B200_ERR_ARG_NULL, "null argument");
  *d = h->device;
  return 0;
}

extern "C" int b200Malloc(b200Handle h, void **p, size_t bytes)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  cudaError_t e = cudaMalloc(p, bytes + B200_ALLOC_PAD);
  if (e != cudaSuccess) {
    b200_set_error(B200_ERR_MEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    *p = NULL;
    return B200_ERR_MEM;
  }
  /* the pad is read (never used) by 16-byte-granular bulk copies: keep it defined */
  B200_CUDA(cudaMemsetAsync((char *)*p + bytes, 0, B200_ALLOC_PAD, h->stream));
  return 0;
}
extern "C" int b200Free(b200Handle h, void *p)
{
  (void)h;
  if (p) B200_CUDA(cudaFree(p));
  return 0;
}
extern "C" int b200MallocHost(void **p, size_t bytes)
{
  B200_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
/* pinned host memory that kernels can write: *d_ptr is the device alias of *h_ptr.  A kernel that leaves its scalar result here
   saves the host one cudaMemcpy per reduction: the host reads *h_ptr after synchronising the stream (cf. the mapped result
   slots of the reduction kernels). */
extern "C" int b200MallocMapped(void **h_ptr, void **d_ptr, size_t bytes)
{
  B200_CUDA(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocMapped));
  B200_CUDA(cudaHostGetDevicePointer(d_ptr, *h_ptr, 0));
  return 0;
}
extern "C" int b200FreeHost(void *p)
{
  if (p) B200_CUDA(cudaFreeHost(p));
  return 0;
}
extern "C" int b200MemcpyHtoD(b200Handle h, void *d, const void *s, size_t bytes)
{
  if (!bytes) return 0;
  g_b200_h2d += (long long)bytes;
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int b200MemcpyDtoH(b200Handle h, void *d, const void *s, size_t bytes)
{
  if (!bytes) return 0;
  g_b200_d2h += (long long)bytes;
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int b200MemcpyHtoDAsync(b200Handle h, void *d, const void *s, size_t bytes)
{
  if (!bytes) return 0;
  g_b200_h2d += (long long)bytes;
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}
extern "C" int b200MemcpyDtoHAsync(b200Handle h, void *d, const void *s, size_t bytes)
{
  if (!bytes) return 0;
  g_b200_d2h += (long long)bytes;
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
extern "C" int b200MemcpyDtoD(b200Handle h, void *d, const void *s, size_t bytes)
{
  if (!bytes) return 0;
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}
extern "C" int b200Memset(b200Handle h, void *d, int byte, size_t bytes)
{
  if (!bytes) return 0;
  B200_CUDA(cudaMemsetAsync(d, byte, bytes, h->stream));
  return 0;
}
extern "C" int b200MemGetInfo(size_t *f, size_t *t)
{
  B200_CUDA(cudaMemGetInfo(f, t));
  return 0;
}

/* ------------------------------------------------------------------ events */
struct b200Event_s {
  cudaEvent_t ev;
};
extern "C" int b200EventCreate(b200Event *ev)
{
  B200_CHECK(ev, B200_ERR_ARG_NULL, "null argument");
  b200Event e = (b200Event)calloc(1, sizeof(*e));
  B200_CHECK(e, B200_ERR_MEM, "out of host memory");
  B200_CUDA(cudaEventCreate(&e->ev));
  *ev = e;
  return 0;
}
extern "C" int b200EventDestroy(b200Event ev)
{
  if (ev) {
    cudaEventDestroy(ev->ev);
    free(ev);
  }
  return 0;
}
extern "C" int b200EventRecord(b200Handle h, b200Event ev)
{
  B200_CHECK(h && ev, B200_ERR_ARG_NULL, "null argument");
  B200_CUDA(cudaEventRecord(ev->ev, h->stream));
  return 0;
}
extern "C" int b200EventSynchronize(b200Event ev)
{
  B200_CHECK(ev, B200_ERR_ARG_NULL, "null event");
  B200_CUDA(cudaEventSynchronize(ev->ev));
  return 0;
}
extern "C" int b200EventElapsedMs(b200Event a, b200Event b, double *ms)
{
  B200_CHECK(a && b && ms, B200_ERR_ARG_NULL, "null argument");
  float f = 0;
  B200_CUDA(cudaEventSynchronize(b->ev));
  B200_CUDA(cudaEventElapsedTime(&f, a->ev, b->ev));
  *ms = f;
  return 0;
}

/* PetscGetMemType analogue (include/petscdevice.h, used by MatSetPreallocationCOO/MatSetValuesCOO to accept host or device arrays) */
extern "C" int b200PointerIsDevice(const void *ptr, int *is_device)
{
  B200_CHECK(is_device, B200_ERR_ARG_NULL, "null argument");
  *is_device = 0;
  if (!ptr) return 0;
  cudaPointerAttributes at;
  cudaError_t           e = cudaPointerGetAttributes(&at, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0; /* unregistered host memory on old drivers */
  }
  *is_device = (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  return 0;
}
