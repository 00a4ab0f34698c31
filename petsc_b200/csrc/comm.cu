/*
 * comm.cu -- multi-GPU plumbing: NCCL replaces MPI for the two communication steps of the Krylov loop.
 *
 *   C1  halo exchange of MatMult_MPIAIJ (mpiaij.c:1047-1061): VecScatterBegin/End -> PetscSFBcastBegin/End_Basic
 *       (sfbasic.c:352-381) pack kernel + MPI_Isend/Irecv/Waitall (sfmpi.c:6-47)  ==>  one pack kernel + one
 *       ncclGroupStart/Send/Recv/End on a second stream, overlapped with the diagonal-block SpMV.  The receive lands
 *       directly in lvec: garray is sorted (mmaij.c:51), so what one owner sends is one contiguous lvec range and the
 *       "to" side of the scatter is a stride (mmaij.c:108) -- no unpack kernel.
 *   C2  MPIU_Allreduce of nv <= 32 scalars in VecMDot_MPI / VecNorm_MPI (pvecimpl.h:97-172)  ==>  ncclAllReduce on the
 *       device results of the MDot/norm kernel, then one device->host copy.
 *
 * NCCL is resolved with dlopen at first use (libnccl.so.2: the copy already loaded by the process if any), so the
 * library has no link-time NCCL dependency and single-GPU use never touches it.
 */
#include "b200_internal.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
enum { NCCL_FLOAT64 = 8, NCCL_INT32 = 2, NCCL_SUM = 0, NCCL_MAX = 2 };

static struct {
  void *lib;
  int (*GetUniqueId)(nccl_uid_t *);
  int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int);
  int (*CommDestroy)(nccl_comm_t);
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  const char *(*GetErrorString)(int);
} N;

static int nccl_load(void)
{
  if (N.lib) return 0;
  const char *cands[] = {getenv("PETSCB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so", NULL};
  for (int i = 0; i < 4 && !N.lib; i++)
    if (cands[i]) N.lib = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  B200_CHECK(N.lib, B200_ERR_GPU_RESOURCE, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(name) \
  do { \
    *(void **)(&N.name) = dlsym(N.lib, "nccl" #name); \
    B200_CHECK(N.name, B200_ERR_GPU_RESOURCE, "libnccl lacks nccl" #name); \
  } while (0)
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllReduce); SYM(Send); SYM(Recv); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
  return 0;
}

#define B200_NCCL(call) \
  do { \
    int r_ = (call); \
    if (r_ != 0) { \
      b200_set_error(B200_ERR_LIB, "NCCL error %d (%s) at %s:%d", r_, N.GetErrorString ? N.GetErrorString(r_) : "?", __FILE__, __LINE__); \
      return B200_ERR_LIB; \
    } \
  } while (0)

extern "C" int b200CommGetUniqueId(void *id128)
{
  B200_CHECK(id128, B200_ERR_ARG_NULL, "null id");
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid_t id;
  B200_NCCL(N.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int b200CommInitRank(b200Handle h, int nranks, int rank, const void *id128)
{
  B200_CHECK(h && id128, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, B200_ERR_ARG_OUTOFRANGE, "bad rank %d of %d", rank, nranks);
  B200_CHECK(!h->nccl_comm, B200_ERR_ORDER, "communicator already initialised");
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid_t id;
  memcpy(&id, id128, sizeof id);
  B200_CUDA(cudaSetDevice(h->device));
  nccl_comm_t c;
  B200_NCCL(N.CommInitRank(&c, nranks, id, rank));
  h->nccl_comm = c;
  h->rank      = rank;
  h->nranks    = nranks;
  return 0;
}

extern "C" int b200CommDestroy(b200Handle h)
{
  if (h && h->nccl_comm) {
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->halo_stream);
    N.CommDestroy((nccl_comm_t)h->nccl_comm);
    h->nccl_comm = NULL;
    h->rank      = 0;
    h->nranks    = 1;
  }
  return 0;
}

extern "C" int b200CommRank(b200Handle h, int *rank, int *nranks)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (rank) *rank = h->rank;
  if (nranks) *nranks = h->nranks;
  return 0;
}

static int allreduce(b200Handle h, double *buf, int count, int op)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (h->nranks == 1 || count == 0) return 0;
  B200_CHECK(h->nccl_comm, B200_ERR_ORDER, "communicator not initialised");
  B200_NCCL(N.AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, op, (nccl_comm_t)h->nccl_comm, h->stream));
  return 0;
}
extern "C" int b200CommAllreduceSum(b200Handle h, double *d_buf, int count) { return allreduce(h, d_buf, count, NCCL_SUM); }
extern "C" int b200CommAllreduceMax(b200Handle h, double *d_buf, int count) { return allreduce(h, d_buf, count, NCCL_MAX); }
extern "C" int b200CommBarrier(b200Handle h)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (h->nranks == 1) return b200Synchronize(h);
  B200_CUDA(cudaMemsetAsync(h->d_result + B200_MAX_NV, 0, sizeof(double), h->stream));
  int rc = allreduce(h, h->d_result + B200_MAX_NV, 1, NCCL_SUM);
  if (rc) return rc;
  return b200Synchronize(h);
}

/* ------------------------------------------------------------------ halo */
struct b200Halo_s {
  int     npeers;
  int    *peers, *send_counts, *send_offsets, *recv_counts, *recv_offsets;
  int     nsend;
  int    *d_send_idx;
  double *d_send_buf;
};

__global__ void halo_pack_kernel(int n, const int *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf)
{
  /* d_Pack of sfcupm_impl.hpp:57: buf[i] = x[idx[i]] */
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = x[idx[i]];
}

extern "C" int b200HaloCreate(b200Handle h, int npeers, const int *peers, const int *send_counts, const int *h_send_idx, const int *recv_counts, const int *recv_offsets, b200Halo *halo)
{
  B200_CHECK(h && halo, B200_ERR_ARG_NULL, "null argument");
  B200_CHECK(npeers >= 0, B200_ERR_ARG_OUTOFRANGE, "negative peer count");
  b200Halo p = (b200Halo)calloc(1, sizeof(*p));
  B200_CHECK(p, B200_ERR_MEM, "out of host memory");
  p->npeers       = npeers;
  p->peers        = (int *)malloc(sizeof(int) * (size_t)(npeers + 1));
  p->send_counts  = (int *)malloc(sizeof(int) * (size_t)(npeers + 1));
  p->send_offsets = (int *)malloc(sizeof(int) * (size_t)(npeers + 1));
  p->recv_counts  = (int *)malloc(sizeof(int) * (size_t)(npeers + 1));
  p->recv_offsets = (int *)malloc(sizeof(int) * (size_t)(npeers + 1));
  int ns          = 0;
  for (int i = 0; i < npeers; i++) {
    B200_CHECK(peers[i] >= 0 && peers[i] < h->nranks && peers[i] != h->rank, B200_ERR_ARG_OUTOFRANGE, "bad peer rank %d", peers[i]);
    p->peers[i]        = peers[i];
    p->send_counts[i]  = send_counts[i];
    p->send_offsets[i] = ns;
    p->recv_counts[i]  = recv_counts[i];
    p->recv_offsets[i] = recv_offsets[i];
    ns += send_counts[i];
  }
  p->nsend = ns;
  if (ns) {
    B200_CUDA(cudaMalloc(&p->d_send_idx, sizeof(int) * (size_t)ns));
    B200_CUDA(cudaMalloc(&p->d_send_buf, sizeof(double) * (size_t)ns));
    B200_CUDA(cudaMemcpyAsync(p->d_send_idx, h_send_idx, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  *halo = p;
  return 0;
}

extern "C" int b200HaloDestroy(b200Halo p)
{
  if (!p) return 0;
  cudaFree(p->d_send_idx); cudaFree(p->d_send_buf);
  free(p->peers); free(p->send_counts); free(p->send_offsets); free(p->recv_counts); free(p->recv_offsets);
  free(p);
  return 0;
}

extern "C" int b200HaloBegin(b200Handle h, b200Halo p, const double *d_x, double *d_lvec)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->npeers) return 0;
  B200_CHECK(h->nccl_comm, B200_ERR_ORDER, "communicator not initialised");
  /* the exchange reads x as produced by work already queued on the main stream */
  B200_CUDA(cudaEventRecord(h->ev_main, h->stream));
  B200_CUDA(cudaStreamWaitEvent(h->halo_stream, h->ev_main, 0));
  if (p->nsend) {
    int g = (p->nsend + 255) / 256;
    if (g > h->num_sms * 4) g = h->num_sms * 4;
    halo_pack_kernel<<<g, 256, 0, h->halo_stream>>>(p->nsend, p->d_send_idx, d_x, p->d_send_buf);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
  }
  B200_NCCL(N.GroupStart());
  for (int i = 0; i < p->npeers; i++) {
    if (p->send_counts[i]) B200_NCCL(N.Send(p->d_send_buf + p->send_offsets[i], (size_t)p->send_counts[i], NCCL_FLOAT64, p->peers[i], (nccl_comm_t)h->nccl_comm, h->halo_stream));
    if (p->recv_counts[i]) B200_NCCL(N.Recv(d_lvec + p->recv_offsets[i], (size_t)p->recv_counts[i], NCCL_FLOAT64, p->peers[i], (nccl_comm_t)h->nccl_comm, h->halo_stream));
  }
  B200_NCCL(N.GroupEnd());
  B200_CUDA(cudaEventRecord(h->ev_halo, h->halo_stream));
  return 0;
}

extern "C" int b200HaloEnd(b200Handle h, b200Halo p)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->npeers) return 0;
  B200_CUDA(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
  return 0;
}

/* setup-time all-to-all of 32-bit index lists (the request lists that PetscSFSetUp exchanges with MPI) */
extern "C" int b200CommAlltoallvInt(b200Handle h, const int *sendcounts, const int *d_send, const int *recvcounts, int *d_recv)
{
  B200_CHECK(h, B200_ERR_ARG_NULL, "null handle");
  if (h->nranks == 1) return 0;
  B200_CHECK(h->nccl_comm, B200_ERR_ORDER, "communicator not initialised");
  size_t so = 0, ro = 0;
  B200_NCCL(N.GroupStart());
  for (int p = 0; p < h->nranks; p++) {
    if (p != h->rank) {
      if (sendcounts[p]) B200_NCCL(N.Send(d_send + so, (size_t)sendcounts[p], NCCL_INT32, p, (nccl_comm_t)h->nccl_comm, h->stream));
      if (recvcounts[p]) B200_NCCL(N.Recv(d_recv + ro, (size_t)recvcounts[p], NCCL_INT32, p, (nccl_comm_t)h->nccl_comm, h->stream));
    }
    so += (size_t)sendcounts[p];
    ro += (size_t)recvcounts[p];
  }
  B200_NCCL(N.GroupEnd());
  return 0;
}

/* ------------------------------------------------------------------ reverse scatter: PetscSFReduceBegin/End with MPI_SUM
   (VecScatterBegin/End(..., ADD_VALUES, SCATTER_REVERSE) in MatMultTranspose_MPIAIJ, mpiaij.c:1086-1097): every rank returns
   its lvec segments to their owners, which add them into y[send_idx].  Same plan, roles swapped: what Bcast received
   contiguously is now sent contiguously (no pack kernel); what Bcast packed is now unpacked with an add.  The adds of one
   peer touch distinct entries (garray is duplicate-free); peers are applied one after the other in plan order, so the
   result does not depend on timing. */
__global__ void halo_unpack_add_kernel(int n, const int *__restrict__ idx, const double *__restrict__ buf, double *y)
{
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[idx[i]] = __dadd_rn(y[idx[i]], buf[i]);
}

extern "C" int b200HaloReduceBegin(b200Handle h, b200Halo p, const double *d_lvec)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->npeers) return 0;
  B200_CHECK(h->nccl_comm, B200_ERR_ORDER, "communicator not initialised");
  B200_CUDA(cudaEventRecord(h->ev_main, h->stream));
  B200_CUDA(cudaStreamWaitEvent(h->halo_stream, h->ev_main, 0));
  B200_NCCL(N.GroupStart());
  for (int i = 0; i < p->npeers; i++) {
    if (p->recv_counts[i]) B200_NCCL(N.Send(d_lvec + p->recv_offsets[i], (size_t)p->recv_counts[i], NCCL_FLOAT64, p->peers[i], (nccl_comm_t)h->nccl_comm, h->halo_stream));
    if (p->send_counts[i]) B200_NCCL(N.Recv(p->d_send_buf + p->send_offsets[i], (size_t)p->send_counts[i], NCCL_FLOAT64, p->peers[i], (nccl_comm_t)h->nccl_comm, h->halo_stream));
  }
  B200_NCCL(N.GroupEnd());
  B200_CUDA(cudaEventRecord(h->ev_halo, h->halo_stream));
  return 0;
}

extern "C" int b200HaloReduceEnd(b200Handle h, b200Halo p, double *d_y)
{
  B200_CHECK(h && p, B200_ERR_ARG_NULL, "null argument");
  if (!p->npeers) return 0;
  B200_CUDA(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
  for (int i = 0; i < p->npeers; i++) {
    const int n = p->send_counts[i];
    if (!n) continue;
    int g = (n + 255) / 256;
    if (g > h->num_sms * 4) g = h->num_sms * 4;
    halo_unpack_add_kernel<<<g, 256, 0, h->stream>>>(n, p->d_send_idx + p->send_offsets[i], p->d_send_buf + p->send_offsets[i], d_y);
    B200_LAUNCHED(1);
    B200_KERNEL_CHECK();
  }
  return 0;
}

/* ------------------------------------------------------------------ MatSetUpMultiply_MPIAIJ, host side (mmaij.c:8-126)
   Shared by the PETSc plugin (mpiaijb200) and the test harness, so the index work that must be bit-exact lives once. */
static int cmp_int_(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* garray = sorted distinct off-process columns (mmaij.c:25-51); h_bj is renumbered in place into positions in garray
   (mmaij.c:55-61).  *h_garray is malloc'ed (free with b200HostFree). */
extern "C" int b200MpiaijBuildGarray(int64_t nzB, int *h_bj, int **h_garray, int *ec)
{
  B200_CHECK(h_garray && ec && (h_bj || !nzB), B200_ERR_ARG_NULL, "null argument");
  int *g = (int *)malloc(sizeof(int) * ((size_t)nzB + 1));
  B200_CHECK(g, B200_ERR_MEM, "out of host memory");
  int n = 0;
  if (nzB) {
    memcpy(g, h_bj, sizeof(int) * (size_t)nzB);
    qsort(g, (size_t)nzB, sizeof(int), cmp_int_);
    n = 1;
    for (int64_t k = 1; k < nzB; k++)
      if (g[k] != g[n - 1]) g[n++] = g[k];
    for (int64_t k = 0; k < nzB; k++) {
      int c = h_bj[k], lo = 0, hi = n - 1;
      while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (g[mid] < c) lo = mid + 1;
        else hi = mid;
      }
      h_bj[k] = lo;
    }
  }
  *h_garray = g;
  *ec       = n;
  return 0;
}

/* host column split of a row block (local rows, GLOBAL columns) into the diagonal block (columns [cstart,cend) renumbered
   to local) and the off-diagonal block (global columns kept): two passes, call with Aj == NULL to get the counts first */
extern "C" int b200MpiaijSplitHost(int m, int cstart, int cend, const int *ai, const int *aj, const double *aa, int64_t *nzA, int64_t *nzB, int *Ai, int *Aj, double *Aa, int *Bi, int *Bj, double *Ba)
{
  B200_CHECK(ai && (aj || !ai[m]) && nzA && nzB, B200_ERR_ARG_NULL, "null argument");
  int64_t ka = 0, kb = 0;
  if (!Aj) {
    for (int r = 0; r < m; r++)
      for (int k = ai[r]; k < ai[r + 1]; k++) {
        if (aj[k] >= cstart && aj[k] < cend) ka++;
        else kb++;
      }
    *nzA = ka;
    *nzB = kb;
    return 0;
  }
  Ai[0] = Bi[0] = 0;
  for (int r = 0; r < m; r++) {
    for (int k = ai[r]; k < ai[r + 1]; k++) {
      const int c = aj[k];
      if (c >= cstart && c < cend) {
        Aj[ka]   = c - cstart;
        Aa[ka++] = aa[k];
      } else {
        Bj[kb]   = c;
        Ba[kb++] = aa[k];
      }
    }
    Ai[r + 1] = (int)ka;
    Bi[r + 1] = (int)kb;
  }
  *nzA = ka;
  *nzB = kb;
  return 0;
}

extern "C" int b200HostFree(void *p)
{
  free(p);
  return 0;
}

/* host values through the device: setup-time collectives on <= a few thousand doubles */
static int allreduce_host(b200Handle h, double *v, int n, int op)
{
  if (h->nranks == 1 || n == 0) return 0;
  double *d = NULL;
  B200_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)n));
  B200_CUDA(cudaMemcpyAsync(d, v, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  int rc = allreduce(h, d, n, op);
  if (!rc) {
    B200_CUDA(cudaMemcpyAsync(v, d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  cudaFree(d);
  return rc;
}

/* The Mvctx scatter of MatSetUpMultiply_MPIAIJ (mmaij.c:103-117) from garray alone: all-gathers the local sizes into the
   ownership ranges (ranges[nranks+1], caller storage), finds the owner of every garray entry (one contiguous lvec range
   per owner), tells each owner which of its entries are wanted (the request exchange PetscSFSetUp does over MPI) and
   creates the halo plan.  Collective over the handle's communicator. */
extern "C" int b200HaloCreateFromGarray(b200Handle h, int m_local, int ec, const int *h_garray, int64_t *ranges, b200Halo *halo)
{
  B200_CHECK(h && halo && ranges && (h_garray || !ec), B200_ERR_ARG_NULL, "null argument");
  const int size = h->nranks, rank = h->rank;
  double   *v    = (double *)calloc((size_t)size * (size + 1), sizeof(double));
  B200_CHECK(v, B200_ERR_MEM, "out of host memory");
  v[rank] = (double)m_local; /* exact below 2^53 */
  int rc  = allreduce_host(h, v, size, NCCL_SUM);
  if (rc) { free(v); return rc; }
  ranges[0] = 0;
  for (int p = 0; p < size; p++) ranges[p + 1] = ranges[p] + (int64_t)v[p];
  int *rcnt = (int *)calloc((size_t)size, sizeof(int)), *roff = (int *)calloc((size_t)size, sizeof(int)), *scnt = (int *)calloc((size_t)size, sizeof(int));
  int  k = 0;
  for (int p = 0; p < size; p++) {
    roff[p] = k;
    while (k < ec && h_garray[k] < ranges[p + 1]) k++;
    rcnt[p] = k - roff[p];
  }
  if (k != ec || rcnt[rank]) {
    free(v); free(rcnt); free(roff); free(scnt);
    B200_CHECK(0, B200_ERR_ARG_OUTOFRANGE, "garray holds columns that are locally owned or outside the global range");
  }
  /* count matrix: row = requester, column = owner */
  double *mat = v;
  memset(mat, 0, sizeof(double) * (size_t)size * size);
  for (int p = 0; p < size; p++) mat[(size_t)rank * size + p] = rcnt[p];
  rc = allreduce_host(h, mat, size * size, NCCL_SUM);
  if (rc) { free(v); free(rcnt); free(roff); free(scnt); return rc; }
  size_t ns = 0;
  for (int p = 0; p < size; p++) {
    scnt[p] = (int)mat[(size_t)p * size + rank];
    ns += (size_t)scnt[p];
  }
  free(v);
  int *req = (int *)malloc(sizeof(int) * ((size_t)ec + 1)), *need = (int *)malloc(sizeof(int) * (ns + 1));
  for (int p = 0, q = 0; q < ec; q++) {
    while (h_garray[q] >= ranges[p + 1]) p++;
    req[q] = (int)(h_garray[q] - ranges[p]); /* local index on the owner */
  }
  if (size > 1) {
    int *d_s = NULL, *d_r = NULL;
    B200_CUDA(cudaMalloc(&d_s, sizeof(int) * ((size_t)ec + 1)));
    B200_CUDA(cudaMalloc(&d_r, sizeof(int) * (ns + 1)));
    if (ec) B200_CUDA(cudaMemcpyAsync(d_s, req, sizeof(int) * (size_t)ec, cudaMemcpyHostToDevice, h->stream));
    rc = b200CommAlltoallvInt(h, rcnt, d_s, scnt, d_r);
    if (!rc && ns) B200_CUDA(cudaMemcpyAsync(need, d_r, sizeof(int) * ns, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_s); cudaFree(d_r);
  }
  int np = 0, *peers = (int *)malloc(sizeof(int) * (size_t)(size + 1)), *psc = (int *)malloc(sizeof(int) * (size_t)(size + 1)), *prc = (int *)malloc(sizeof(int) * (size_t)(size + 1)), *pro = (int *)malloc(sizeof(int) * (size_t)(size + 1));
  int *sidx = (int *)malloc(sizeof(int) * (ns + 1)), nsi = 0, off = 0, bad = 0;
  for (int p = 0; p < size && !rc; p++) {
    if (p != rank && (scnt[p] || rcnt[p])) {
      peers[np] = p; psc[np] = scnt[p]; prc[np] = rcnt[p]; pro[np] = roff[p];
      for (int q = 0; q < scnt[p]; q++) {
        if (need[off + q] < 0 || need[off + q] >= m_local) bad = 1;
        sidx[nsi++] = need[off + q];
      }
      np++;
    }
    off += scnt[p];
  }
  if (!rc && bad) { b200_set_error(B200_ERR_ARG_OUTOFRANGE, "a peer requested an entry outside my row range"); rc = B200_ERR_ARG_OUTOFRANGE; }
  if (!rc) rc = b200HaloCreate(h, np, peers, psc, sidx, prc, pro, halo);
  free(rcnt); free(roff); free(scnt); free(req); free(need); free(peers); free(psc); free(prc); free(pro); free(sidx);
  return rc;
}
