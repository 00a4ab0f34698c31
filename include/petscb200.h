/*
 * petscb200.h -- C ABI of libpetscb200.so: hand-written sm_100a kernels for PETSc's Krylov inner loop.
 *
 * This is the drop-in boundary.  The reference's device back-ends reach their kernels through a C
 * library handle obtained from PETSc's device layer -- PetscCUBLASGetHandle() / PetscGetCurrentCUDAStream()
 * (include/petscdevice_cuda.h:180-183) -- and then call cusparseSpMV / cublasXdot / ... with raw device pointers.
 * Every entry point below replaces one such call site (cited per function); arguments are plain pointers and sizes,
 * no PETSc, torch or C++ types.  Host code above this ABI stays in C: either the PETSc type plugin
 * (petsc_plugin/, -mat_type aijb200 -vec_type b200) or the stand-alone host mirror (include/petscb200_host.h).
 *
 * Conventions
 *  - PetscScalar = double, PetscInt = int32 (the reference's default configuration).
 *  - Every function returns 0 (PETSC_SUCCESS) or a PetscErrorCode value (B200_ERR_*), never throws or exits;
 *    b200GetLastErrorString() gives the message PetscCallCUDA would have printed.
 *  - Pointers named d_* are device pointers; they must come from b200Malloc (256-byte aligned, padded by 256 bytes:
 *    the SpMV kernel issues 16-byte-granular TMA bulk copies that may over-read a CSR array by < 16 bytes).
 *  - All work is enqueued on the handle's stream (default: a non-blocking stream created with the handle; a PETSc
 *    plugin passes PetscDefaultCudaStream via b200SetStream).  Functions returning scalars to host memory synchronise
 *    that stream, exactly as cublasDdot with CUBLAS_POINTER_MODE_HOST does.
 */
#ifndef PETSCB200_H
#define PETSCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PetscErrorCode values (include/petscsystypes.h:45-91) */
enum {
  B200_SUCCESS              = 0,
  B200_ERR_MEM              = 55,
  B200_ERR_SUP              = 56,
  B200_ERR_ORDER            = 58,
  B200_ERR_ARG_SIZ          = 60,
  B200_ERR_ARG_WRONG        = 62,
  B200_ERR_ARG_OUTOFRANGE   = 63,
  B200_ERR_MAT_LU_ZRPVT     = 71,
  B200_ERR_ARG_WRONGSTATE   = 73,
  B200_ERR_LIB              = 76,
  B200_ERR_ARG_NULL         = 85,
  B200_ERR_GPU_RESOURCE     = 96,
  B200_ERR_GPU              = 97
};

typedef struct b200Handle_s *b200Handle;

/* ---- handle / stream (replaces PetscCUBLASGetHandle + PetscGetCurrentCUDAStream, petscdevice_cuda.h:180-183) ---- */
int         b200Create(b200Handle *h, int device);          /* device < 0: current device */
int         b200Destroy(b200Handle h);
int         b200SetStream(b200Handle h, void *cudaStream);  /* cudaStream_t; NULL restores the handle's own stream */
int         b200GetStream(b200Handle h, void **cudaStream);
int         b200Synchronize(b200Handle h);
int         b200DeviceSynchronize(void);                    /* every stream of the current device (cudaDeviceSynchronize): for callers that mix handles */
int         b200GetDevice(b200Handle h, int *device);
int         b200DeviceCount(int *n);
const char *b200GetLastErrorString(void);
const char *b200Version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
long long   b200KernelLaunchCount(void);
/* bytes moved host->device / device->host through b200Memcpy* since load (the PetscLogCpuToGpu / PetscLogGpuToCpu counters
   of aijcusparse.cu:1563, veccupmimpl.h:430; results of reductions that arrive through mapped pinned memory are not copies) */
int         b200TransferCounters(long long *h2d_bytes, long long *d2h_bytes);
/* PetscGetMemType analogue: 1 if ptr is device (or managed) memory, 0 for host memory */
int         b200PointerIsDevice(const void *ptr, int *is_device);

/* timing on the handle's stream (cudaEvent pair; bench.py measures kernels with these, not with host clocks) */
typedef struct b200Event_s *b200Event;
int b200EventCreate(b200Event *ev);
int b200EventDestroy(b200Event ev);
int b200EventRecord(b200Handle h, b200Event ev);
int b200EventSynchronize(b200Event ev);               /* host waits for the work recorded BEFORE ev only (later kernels keep running) */
int b200EventElapsedMs(b200Event start, b200Event stop, double *ms); /* synchronises on stop */

/* ---- memory (replaces cudaMalloc/cudaMemcpy in aijcusparse.cu:1477-1590, veccupmimpl.h:391-440) ---- */
int b200Malloc(b200Handle h, void **d_ptr, size_t bytes);   /* padded + aligned as the kernels require */
int b200Free(b200Handle h, void *d_ptr);
int b200MallocHost(void **h_ptr, size_t bytes);              /* pinned */
int b200FreeHost(void *h_ptr);
/* pinned host memory mapped into the device address space (*d_ptr aliases *h_ptr): a kernel writes its scalar result there and the
   host reads it after b200Synchronize -- no cudaMemcpy on the latency-critical norm of every Krylov iteration; free with b200FreeHost(h_ptr) */
int b200MallocMapped(void **h_ptr, void **d_ptr, size_t bytes);
int b200MemcpyHtoD(b200Handle h, void *d_dst, const void *h_src, size_t bytes); /* stream-ordered, returns after completion */
int b200MemcpyDtoH(b200Handle h, void *h_dst, const void *d_src, size_t bytes);
int b200MemcpyDtoD(b200Handle h, void *d_dst, const void *d_src, size_t bytes); /* asynchronous */
int b200MemcpyHtoDAsync(b200Handle h, void *d_dst, const void *h_src, size_t bytes);
int b200MemcpyDtoHAsync(b200Handle h, void *h_dst, const void *d_src, size_t bytes);
int b200Memset(b200Handle h, void *d_ptr, int byte, size_t bytes);
int b200MemGetInfo(size_t *free_bytes, size_t *total_bytes);

/* ---- CSR SpMV: MatMult_SeqAIJ (aij.c:1444) / cusparseSpMV CSR_ALG1 call site aijcusparse.cu:2528 ---- */
typedef struct b200CsrPlan_s *b200CsrPlan;
/* analysis (the cusparseSpMV_preprocess analogue, aijcusparse.cu:2517): row-length statistics, row-tile size and
   lanes-per-row selection.  d_rowptr[m+1], d_colidx[nnz] are kept by reference (not copied). */
int b200CsrPlanCreate(b200Handle h, int m, int n, int64_t nnz, const int *d_rowptr, const int *d_colidx, b200CsrPlan *plan);
int b200CsrPlanDestroy(b200CsrPlan plan);
/* tuning / introspection: lanes_per_row in {0=auto,1,2,4,8,16,32}; lanes_per_row == 1 reproduces MatMult_SeqAIJ's
   strict left-to-right, FMA-free row sums bit for bit.  rows_per_tile 0 = auto. */
int b200CsrPlanSetLayout(b200CsrPlan plan, int lanes_per_row, int rows_per_tile, int stages, int ctas_per_sm);
/* row-sum association when lanes_per_row > 1 (one lane per row is always the reference's order):
   1 (default) = FMA + shuffle tree: fastest, equal to MatMult_SeqAIJ to rounding (<= 1e-12 relative);
   0 = the reference's strict left-to-right, FMA-free order (the lanes load and multiply a chunk of the row together, the
       products are then added in column order): y is bit-identical to MatMult_SeqAIJ for every layout, at 0.6-1.0x the
       speed of the tree variant (5x faster than forcing one lane per row on 27-entry rows) */
int b200CsrPlanSetSummation(b200CsrPlan plan, int tree);
/* column-blocked layout for matrices whose gathered vector does not stay in the L2 (random CSR with n*8 bytes >> L2/2): the
   columns are cut into nblocks ranges and y = A x runs as nblocks passes y += A_b x in column order, each touching only
   8*n/nblocks bytes of x; every pass continues the row sum of the previous one, so an exact summation stays exact.
   nblocks <= 1 removes the blocked copy.  The blocked SpMV multiplies with a PACKED copy of the values:
   b200CsrPlanPackValues must be called after b200CsrPlanSetColumnBlocks and after every change of the values; the d_val
   argument of b200CsrSpMV* is then ignored.  Costs 16 B/nnz of extra device memory. */
int b200CsrPlanSetColumnBlocks(b200Handle h, b200CsrPlan plan, int nblocks);
int b200CsrPlanPackValues(b200Handle h, b200CsrPlan plan, const double *d_val);
/* chooses the number of column blocks from n, the mean row length and the measured mean column span of a row (0 = none:
   stencil-like or L2-resident gathers); callers that use it must keep the packed values coherent (b200CsrPlanPackValues after
   every value change) -- the PETSc plugin does so through the Mat's object state */
int b200CsrPlanAutoColumnBlocks(b200Handle h, b200CsrPlan plan, int *nblocks_chosen);
/* L2 hints: bit0 = stream val/col/rowptr as evict_first, bit1 = keep x as evict_last, bit2 = persisting access-policy
   window on x for the launch (default: 2 for one lane per row, else 3, plus bit2 when x fits the L2 set-aside) */
int b200CsrPlanSetCacheHints(b200CsrPlan plan, int hints);
int b200CsrPlanGetLayout(b200CsrPlan plan, int *lanes_per_row, int *rows_per_tile, int *stages, int *grid, int *smem_bytes, int *max_row_nnz);
/* y = A x                                             (MatMult_SeqAIJ, aij.c:1444-1499) */
int b200CsrSpMV(b200Handle h, b200CsrPlan plan, const double *d_val, const double *d_x, double *d_y);
/* z = y + A x  (z may alias y)                        (MatMultAdd_SeqAIJ, aij.c:1606-1655) */
int b200CsrSpMVAdd(b200Handle h, b200CsrPlan plan, const double *d_val, const double *d_x, const double *d_y, double *d_z);
/* w = dinv .* (A x), and y = A x too when d_y != NULL  (MatMult + PCApply_Jacobi fused: precon.c:853-854 + jacobi.c:354) */
int b200CsrSpMVJacobi(b200Handle h, b200CsrPlan plan, const double *d_val, const double *d_x, const double *d_dinv, double *d_w, double *d_y);
/* compressed-row z = y + A x for a block whose rows are mostly empty (off-diagonal block B of MATMPIAIJ):
   only the nrows_c rows listed in d_rindex are read/written   (MatMultAdd_SeqAIJ compressed branch, aij.c:1626-1640) */
int b200CsrSpMVAddCompressed(b200Handle h, int nrows_c, const int *d_cr_i, const int *d_rindex, const int *d_colidx, const double *d_val, const double *d_x, const double *d_y, double *d_z);
/* MatMult_MPIAIJ + PCApply_Jacobi fused on the rows that own off-diagonal entries: after b200CsrSpMVJacobi wrote
   w = dinv .* (A_d x) on the diagonal block, this overwrites w[r] = dinv[r] * (A_d(r,:) x  +  B(r,:) lvec) for the nrows_c
   rows of d_rindex, the row sum taken left to right through A_d's entries and then B's (mpiaij.c:1057-1060: the off-diagonal
   multadd continues the diagonal block's y[r]; jacobi.c:354) -- the unfused result bit for bit, without a y vector */
int b200CsrSpMVAddCompressedJacobi(b200Handle h, int nrows_c, const int *d_cr_i, const int *d_rindex, const int *d_bj, const double *d_ba, const double *d_lvec,
                                   const int *d_ai, const int *d_aj, const double *d_aa, const double *d_x, const double *d_dinv, double *d_w);
/* MatAssemblyEnd_SeqAIJ invariants checked on the device: *bad_row = -1 if valid, else an offending row and
   kind 1 = column out of range, 2 = columns not strictly increasing, 3 = decreasing row pointer */
int b200CsrValidate(b200Handle h, int m, int n, const int *d_rowptr, const int *d_colidx, int *bad_row, int *kind);
int b200CsrCountNonemptyRows(b200Handle h, int m, const int *d_rowptr, int *count_host);   /* MatCheckCompressedRow input */
/* d[r] = a[diag(r)] or 0 ; d_diagpos (nullable) gets the position or -1   (MatGetDiagonal_SeqAIJ aij.c:1347, GetDiagonal_CSR aijcupm.hpp:115) */
int b200CsrGetDiagonal(b200Handle h, int m, const int *d_rowptr, const int *d_colidx, const double *d_val, double *d_diag, int *d_diagpos);
/* dinv[r] = 1/d[r], zero diagonal -> 1.0             (PCSetUp_Jacobi, jacobi.c:172-270) ; *nzero_host counts the zeros */
int b200JacobiInvertDiagonal(b200Handle h, int64_t n, const double *d_diag, double *d_dinv, int *nzero_host);

/* setup-time column split of a row block into the MATMPIAIJ diagonal block (local columns [cstart,cend), renumbered)
   and off-diagonal block (global columns kept)   (MatSetUpMultiply_MPIAIJ first step, mmaij.c:25-61).  Outputs are
   b200Malloc'ed and owned by the caller. */
int b200CsrSplitColumns(b200Handle h, int m, const int *d_i, const int *d_j, const double *d_a, int cstart, int cend,
                        int **d_Ai, int **d_Aj, double **d_Aa, int64_t *nzA, int **d_Bi, int **d_Bj, double **d_Ba, int64_t *nzB);

/* ---- BLAS-1: VECSEQ ops / cuBLAS + MDot_kernel/MAXPY_kernel call sites in vecseqcupm_impl.hpp ---- */
int b200VecSet(b200Handle h, int64_t n, double alpha, double *d_x);                                   /* VecSet */
int b200VecCopy(b200Handle h, int64_t n, const double *d_x, double *d_y);                             /* VecCopy */
int b200VecScale(b200Handle h, int64_t n, double alpha, double *d_x);                                 /* bvec1.c:51 / :1473 cublasXscal */
int b200VecAXPY(b200Handle h, int64_t n, double alpha, const double *d_x, double *d_y);               /* bvec1.c:70 / :519 cublasXaxpy */
int b200VecAYPX(b200Handle h, int64_t n, double alpha, const double *d_x, double *d_y);               /* dvec2.c:753  y = x + a y */
int b200VecAXPBY(b200Handle h, int64_t n, double alpha, double beta, const double *d_x, double *d_y); /* bvec1.c:91   y = a x + b y */
int b200VecWAXPY(b200Handle h, int64_t n, double alpha, const double *d_x, const double *d_y, double *d_w); /* dvec2.c:791 w = a x + y */
int b200VecPointwiseMult(b200Handle h, int64_t n, const double *d_x, const double *d_y, double *d_w); /* bvec2.c:72 */
int b200VecPointwiseDivide(b200Handle h, int64_t n, const double *d_x, const double *d_y, double *d_w);
int b200VecReciprocal(b200Handle h, int64_t n, double *d_x);
int b200VecShift(b200Handle h, int64_t n, double shift, double *d_x);
/* reductions: result in host memory, stream synchronised on return */
int b200VecDot(b200Handle h, int64_t n, const double *d_x, const double *d_y, double *result);        /* bvec1.c:33 / :1125 cublasXdot */
int b200VecNorm2(b200Handle h, int64_t n, const double *d_x, double *result);                         /* bvec2.c:201 / :1749 cublasXnrm2 */
int b200VecNorm(b200Handle h, int64_t n, const double *d_x, int type /*0:1-norm 1:2-norm 3:inf*/, double *result);
int b200VecSum(b200Handle h, int64_t n, const double *d_x, double *result);
int b200VecMax(b200Handle h, int64_t n, const double *d_x, int64_t *idx, double *result);
int b200VecMin(b200Handle h, int64_t n, const double *d_x, int64_t *idx, double *result);
/* z[j] = x . y_j, j < nv : ONE kernel, x read once       (VecMDot_Seq dvec2.c:83 / MDot_kernel vecseqcupm_impl.hpp:1148-1360)
   y = host array of nv device pointers */
int b200VecMDot(b200Handle h, int64_t n, int nv, const double *d_x, const double *const *y, double *result);
/* x += sum_j alpha_j y_j : ONE pass, association identical to VecMAXPY_Seq (dvec2.c:658, petscaxpy.h:125): the
   nv&3 remainder group first, then groups of 4, FMA-free -> bit-identical to the CPU reference.
   norm2_out (nullable): 2-norm of the UPDATED x from the same pass (fused MAXPY+norm; host value, synchronises) */
int b200VecMAXPY(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *d_x, double *norm2_out);
/* y += alpha x and dot = y_new . z in one pass (fused AXPY+dot; z may be y for the squared norm) */
int b200VecAXPYDot(b200Handle h, int64_t n, double alpha, const double *d_x, double *d_y, const double *d_z, double *result);
/* the eight vector recurrences of one KSPPIPECG iteration in one pass (cg/pipecg/pipecg.c:124-141):
     first != 0:  z = n, q = m, p = u, s = w                       (VecCopy x4)
     else      :  z = n + beta z, q = m + beta q, p = u + beta p, s = w + beta s   (VecAYPX x4)
     then      :  x += alpha p, u -= alpha q, w -= alpha z, r -= alpha s             (VecAXPY x4)
   bit-identical to the eight separate calls; all ten vectors must be distinct */
int b200VecPipeCGUpdate(b200Handle h, int64_t n, double alpha, double beta, int first, const double *d_n, const double *d_m, double *d_u, double *d_w,
                        double *d_z, double *d_q, double *d_p, double *d_s, double *d_x, double *d_r);
/* device-result variants (no host sync): results land in d_result[nv]; used by the MPI vector type, which all-reduces
   them before the single device->host copy (pvecimpl.h:97-172) */
int b200VecMDotAsync(b200Handle h, int64_t n, int nv, const double *d_x, const double *const *y, double *d_result);
int b200VecMAXPYAsync(b200Handle h, int64_t n, int nv, const double *alpha, const double *const *y, double *d_x, double *d_sumsq /*nullable*/);

/* ---- ILU(0): MatILUFactorSymbolic_SeqAIJ_ilu0 / MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ_NaturalOrdering
        (aijfact.c:1471, 216, 2413) ; cusparseXcsrilu02 + cusparseSpSV call sites aijcusparse.cu:766-827, 643-693 ---- */
typedef struct b200IluPlan_s *b200IluPlan;
/* symbolic: factor layout bi/bj/bdiag exactly as aijfact.c:1454-1469, plus dependency analysis for the device solves.
   Host CSR pattern in, everything else on device. */
int b200Ilu0Symbolic(b200Handle h, int n, const int *h_ai, const int *h_aj, b200IluPlan *plan);
int b200Ilu0Destroy(b200IluPlan plan);
/* numeric factorisation on device from A's device values (same row/element operation order as aijfact.c:216-389,
   FMA-free => factor bit-identical to the CPU reference); shifttype NONZERO semantics (matimpl.h:795-811) */
int b200Ilu0Numeric(b200Handle h, b200IluPlan plan, const double *d_aval, double zeropivot, double shiftamount, int *nshift);
/* x = U^{-1} L^{-1} b (aijfact.c:2413-2457), rows summed left to right without FMA: bit-identical */
int b200Ilu0Solve(b200Handle h, b200IluPlan plan, const double *d_b, double *d_x);
int b200Ilu0GetFactor(b200Handle h, b200IluPlan plan, int *h_bi, int *h_bj, int *h_bdiag, double *h_ba); /* tests: copies out */
int b200Ilu0GetInfo(b200IluPlan plan, int *nlevels_lower, int *nlevels_upper, int64_t *nnz);
/* schedule of the segment-marching sweeps (ilu.cu): lanes per row, padded segment slots and segment dependency levels */
int b200Ilu0GetSegmentInfo(b200IluPlan plan, int *lanes, int *nslot_lower, int *nlev_lower, int *nslot_upper, int *nlev_upper);

/* ---- ICC(0): MatICCFactorSymbolic_SeqAIJ (levels 0, natural ordering; aijfact.c:2049-2094),
        MatCholeskyFactorNumeric_SeqAIJ (aijfact.c:1701-1866), MatSolve_SeqSBAIJ_1_NaturalOrdering (sbaijfact2.c:2030-2065);
        the reference's GPU path is cusparseXcsric02 + cusparseSpSV (aijcusparse.cu) ---- */
typedef struct b200IccPlan_s *b200IccPlan;
/* symbolic: the reference's factor layout (row i = strictly upper entries of A's row i, diagonal LAST), the order in which
   the reference's linked lists merge finished rows into a row (it fixes the rounding of every entry), dependency levels, the
   column view for the gather form of the forward sweep, segment schedules.  Host CSR pattern in. */
int b200Icc0Symbolic(b200Handle h, int n, const int *h_ai, const int *h_aj, b200IccPlan *plan);
int b200Icc0Destroy(b200IccPlan plan);
/* numeric factorisation on the device from A's device values, contributors merged in the reference's order, FMA-free => factor
   bit-identical to the CPU reference.  *zero_pivot_row = 0, or 1 + a row whose pivot failed dk > zeropivot*rowsum
   (MatPivotCheck_pd, matimpl.h:813-833, would shift and refactor; here the factorisation is reported as failed) */
int b200Icc0Numeric(b200Handle h, b200IccPlan plan, const double *d_aval, double zeropivot, int *zero_pivot_row);
/* x = U^-1 D^-1 U^-T b, operation order of MatSolve_SeqSBAIJ_1_NaturalOrdering: bit-identical */
int b200Icc0Solve(b200Handle h, b200IccPlan plan, const double *d_b, double *d_x);
int b200Icc0GetFactor(b200Handle h, b200IccPlan plan, int *h_ui, int *h_uj, int *h_udiag, double *h_ua); /* tests: copies out */
int b200Icc0GetInfo(b200IccPlan plan, int64_t *nz_factor, int *nlevels_numeric, int *nlev_forward, int *nlev_backward);

/* ---- multi-GPU: NCCL replaces MPI in VecScatter/PetscSF (sfbasic.c:352-381, sfmpi.c:6-47) and MPIU_Allreduce
        (pvecimpl.h:101-171) ---- */
#define B200_UNIQUE_ID_BYTES 128
int b200CommGetUniqueId(void *id128);                                    /* rank 0, then broadcast out-of-band */
int b200CommInitRank(b200Handle h, int nranks, int rank, const void *id128);
int b200CommDestroy(b200Handle h);
int b200CommRank(b200Handle h, int *rank, int *nranks);                  /* 0,1 when no communicator */
int b200CommAllreduceSum(b200Handle h, double *d_buf, int count);        /* in place, on the handle's stream */
int b200CommAllreduceMax(b200Handle h, double *d_buf, int count);
int b200CommBarrier(b200Handle h);
/* setup-time exchange of 32-bit index lists: segment p of d_send (sendcounts[p] ints, rank order) goes to rank p */
int b200CommAlltoallvInt(b200Handle h, const int *sendcounts, const int *d_send, const int *recvcounts, int *d_recv);
/* Halo plan = the Mvctx scatter of MatSetUpMultiply_MPIAIJ (mmaij.c:108-117): for each peer, which of MY entries it
   needs (send lists, local indices) and how many consecutive lvec slots I receive from it (garray is sorted, so the
   entries owned by one rank are contiguous in lvec: no unpack kernel). */
typedef struct b200Halo_s *b200Halo;
int b200HaloCreate(b200Handle h, int npeers, const int *peers, const int *send_counts, const int *h_send_idx /*concatenated*/,
                   const int *recv_counts, const int *recv_offsets, b200Halo *halo);
int b200HaloDestroy(b200Halo halo);
/* pack x[send_idx] and ncclSend/ncclRecv on the handle's halo stream (ordered after work already on the main stream) */
int b200HaloBegin(b200Handle h, b200Halo halo, const double *d_x, double *d_lvec);
/* make the main stream wait for the exchange */
int b200HaloEnd(b200Handle h, b200Halo halo);

/* reverse scatter = PetscSFReduceBegin/End(MPI_SUM) of VecScatter(ADD_VALUES, SCATTER_REVERSE) in MatMultTranspose_MPIAIJ
   (mpiaij.c:1086-1097): every rank returns its lvec segments to their owners (halo stream), which add them into
   y[send_idx] on the main stream, peer after peer in plan order (deterministic) */
int b200HaloReduceBegin(b200Handle h, b200Halo halo, const double *d_lvec);
int b200HaloReduceEnd(b200Handle h, b200Halo halo, double *d_y);
/* MatSetUpMultiply_MPIAIJ host side, shared by the PETSc plugin and the test harness (index work, bit-exact):
   b200MpiaijSplitHost: column split of a row block with GLOBAL columns into the diagonal block (local columns) and the
     off-diagonal block (global columns); call with Aj == NULL first for the counts        (mpiaij.c MatSetValues_MPIAIJ routing)
   b200MpiaijBuildGarray: garray = sorted distinct off-process columns, h_bj renumbered in place (mmaij.c:25-61);
     *h_garray is malloc'ed: release with b200HostFree
   b200HaloCreateFromGarray: ownership ranges (all-gather of m_local; ranges[nranks+1] is caller storage), owners of the
     garray entries, the request exchange and the halo plan (mmaij.c:103-117 + PetscSFSetUp); collective */
int b200MpiaijSplitHost(int m, int cstart, int cend, const int *ai, const int *aj, const double *aa, int64_t *nzA, int64_t *nzB, int *Ai, int *Aj, double *Aa, int *Bi, int *Bj, double *Ba);
int b200MpiaijBuildGarray(int64_t nzB, int *h_bj, int **h_garray, int *ec);
int b200HaloCreateFromGarray(b200Handle h, int m_local, int ec, const int *h_garray, int64_t *ranges, b200Halo *halo);
int b200HostFree(void *p);

/* ---- COO assembly ----------------------------------------------------------------------------------
 * replaces MatSetPreallocationCOO_SeqAIJ / MatSetValuesCOO_SeqAIJ (src/mat/impls/aij/seq/aij.c:4524-4732) and the value
 * kernel of MatSetValuesCOO_SeqAIJCUSPARSE (aijcusparse.cu).  Index arrays are DEVICE pointers; entries with a negative
 * row or column are ignored; a row >= M or column >= N is B200_ERR_ARG_OUTOFRANGE (aij.c:4561,4631).  The plan owns the
 * CSR pattern it builds (sorted, unique columns per row).  Repeated (i,j) pairs are added in the order of the user's
 * array; b200CooPlanCreateFromMaps instead adopts the reference's own jmap[nnz+1] / perm[atot] (MatCOOStruct_SeqAIJ,
 * aij.h:170-176, HOST arrays of PetscCount) and then reproduces the reference's values bit for bit. */
typedef struct b200CooPlan_s *b200CooPlan;
int b200CooPlanCreate(b200Handle h, int M, int N, int64_t coo_n, const int *d_coo_i, const int *d_coo_j, b200CooPlan *plan);
int b200CooPlanCreateFromMaps(b200Handle h, int64_t nnz, int64_t atot, const int64_t *h_jmap, const int64_t *h_perm, b200CooPlan *plan);
int b200CooPlanDestroy(b200CooPlan plan);
/* borrowed device pointers to rowptr[M+1] / colidx[nnz] (NULL for plans made from maps); atot = number of entries kept */
int b200CooPlanGetCsr(b200CooPlan plan, int64_t *nnz, int64_t *atot, const int **d_rowptr, const int **d_colidx);
/* copies jmap[nnz+1] and perm[atot] to HOST arrays (tests) */
int b200CooPlanGetMaps(b200Handle h, b200CooPlan plan, int *h_jmap, int *h_perm);
/* a[q] = (insert ? 0 : a[q]) + (0 + v[perm[jmap[q]]] + ... )   (aij.c:4724-4728); d_v has coo_n entries */
int b200CooSetValues(b200Handle h, b200CooPlan plan, const double *d_v, int insert, double *d_a);

/* ---- transposed product ---------------------------------------------------------------------------
 * replaces MatMultTranspose_SeqAIJ / MatMultTransposeAdd_SeqAIJ (aij.c:1383-1440; reference device path: cusparseSpMV with
 * CUSPARSE_OPERATION_TRANSPOSE or an explicit transpose, aijcusparse.cu MatSeqAIJCUSPARSEFormExplicitTranspose).  The
 * transposed pattern is built once on the device; b200CsrTransposeSetValues re-gathers the values after A changed;
 * the product adds x[i]*a(i,c) into y[c] in increasing row order starting from 0 (or z[c]) exactly like the reference
 * (with one lane per row; see b200CsrTransposeGetPlan). */
typedef struct b200CsrTranspose_s *b200CsrTranspose;
int b200CsrTransposeCreate(b200Handle h, int m, int n, int64_t nnz, const int *d_rowptr, const int *d_colidx, b200CsrTranspose *T);
int b200CsrTransposeDestroy(b200CsrTranspose T);
int b200CsrTransposeSetValues(b200Handle h, b200CsrTranspose T, const double *d_a);
int b200CsrTransposeSpMV(b200Handle h, b200CsrTranspose T, const double *d_x, const double *d_z, double *d_y);
/* the SpMV plan on the transposed pattern, for b200CsrPlanSetLayout: lanes_per_row = 1 is the bit-exact parity mode (as for
   MatMult); the automatic layout uses several lanes per long row and is then exact only to rounding */
int b200CsrTransposeGetPlan(b200CsrTranspose T, b200CsrPlan *plan);
/* copies the transposed pattern (tptr[n+1], trow[nnz]) and the value permutation tperm[nnz] to HOST arrays (tests) */
int b200CsrTransposeGet(b200Handle h, b200CsrTranspose T, int *h_tptr, int *h_trow, int *h_tperm);

/* ---- local star-forest broadcast / reduction on device data -------------------------------------------
 * replaces PetscSFLinkScatterLocal (sfpack.c:1082) with the host ScatterAnd<Op> loops of sfpack.c:190-222 (reference device
 * path: the d_ScatterAnd<Op> kernels of src/vec/is/sf/impls/basic/cupm): for i = 0 .. n-1 IN THIS ORDER
 *     dst[didx[i]*bs + c] = dst[didx[i]*bs + c] <op> src[sidx[i]*bs + c],   c < bs.
 * sidx / didx are HOST int arrays (NULL = contiguous from s0 / d0); the plan copies what it needs to the device.  Entries that
 * share a destination are applied in entry order (stable grouping, no atomics), so sums equal the reference's bit for bit.
 * dtype: B200_SF_F64 (PetscScalar / PetscReal) or B200_SF_I32 (PetscInt); op: the MPI_Op a VecScatter can ask for
 * (vscat.c:62-66) plus MPI_PROD.  d_src == d_dst (an in-place scatter has sequential semantics) is B200_ERR_SUP. */
enum { B200_SF_F64 = 0, B200_SF_I32 = 1 };
enum { B200_SF_REPLACE = 0, B200_SF_SUM = 1, B200_SF_PROD = 2, B200_SF_MAX = 3, B200_SF_MIN = 4 };
typedef struct b200IndexedPlan_s *b200IndexedPlan;
int b200IndexedPlanCreate(b200Handle h, int64_t n, const int *h_sidx, int s0, const int *h_didx, int d0, b200IndexedPlan *plan);
int b200IndexedPlanDestroy(b200Handle h, b200IndexedPlan plan);
/* the host half of the plan (no device): contiguity, extents and the stable grouping by destination; gdst[ngroups], goff[ngroups+1],
   gsrc[n] are malloc'ed (b200HostFree) when *grouped, else NULL */
int b200IndexedGroupHost(int64_t n, const int *h_sidx, int s0, const int *h_didx, int d0, int *src_contig, int *dst_contig, int64_t *src_extent, int64_t *dst_extent, int *grouped, int64_t *ngroups, int **gdst, int **goff, int **gsrc);
int b200IndexedPlanGetInfo(b200IndexedPlan plan, int64_t *n, int64_t *ngroups, int *grouped, int *src_contig, int *dst_contig, int64_t *src_extent, int64_t *dst_extent);
int b200IndexedOp(b200Handle h, b200IndexedPlan plan, int dtype, int bs, int op, const void *d_src, void *d_dst);

/* ---- device-side generators of the benchmark operators (bench/test utility; SURVEY 8d inputs) ---- */
/* rows [r0,r1) of the 7-point nx*ny*nz Laplacian, local row pointer, GLOBAL 32-bit columns */
int b200GenLaplace7(b200Handle h, int nx, int ny, int nz, int64_t r0, int64_t r1, int *d_rowptr, int *d_colidx, double *d_val);
int b200GenLaplace7Nnz(int nx, int ny, int nz, int64_t r0, int64_t r1, int64_t *nnz);
/* the 27-point n^3 operator of bench_kspsolve.c:115-303; d_rowptr[n^3+1] */
int b200GenLaplace27(b200Handle h, int n, int *d_rowptr, int *d_colidx, double *d_val);
int b200GenLaplace27Nnz(int n, int64_t *nnz);
/* random CSR, fixed row length d, sorted distinct stratified columns, values in (-1,1)   (BASELINE config 5) */
int b200GenRandomCsr(b200Handle h, int n, int ncols, int d, uint64_t seed, int *d_rowptr, int *d_colidx, double *d_val);

#ifdef __cplusplus
}
#endif
#endif
