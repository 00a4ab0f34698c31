/* mat.c -- Mat interface (src/mat/interface/{matrix.c,matreg.c}) and the device matrix types
   seqaijb200 (Mat_SeqAIJ CSR, aij.h:47-92, ops of aij.c) and mpiaijb200 (Mat_MPIAIJ = diag block A + off-diag block B +
   garray + lvec + scatter, mpiaij.h:41-76, mpiaij.c:1047-1061, mmaij.c:8-126). */
#include "hostimpl.h"

#define H (PetscB200.h)
static PetscFunctionList MatList = NULL;
static int               MatRegisterAllCalled = 0;
static struct {
  char r[32], s[32], m[32];
} rootnames[8];
static int nroot = 0;

PetscErrorCode MatRegister(const char sname[], PetscErrorCode (*function)(Mat)) { return PetscFunctionListAdd(&MatList, sname, (void *)function); }
PetscErrorCode MatRegisterRootName(const char rname[], const char sname[], const char mname[])
{
  PetscCheck(nroot < 8, 0, PETSC_ERR_MEM, "root-name table full");
  strncpy(rootnames[nroot].r, rname, 31);
  strncpy(rootnames[nroot].s, sname, 31);
  strncpy(rootnames[nroot].m, mname, 31);
  nroot++;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatRegisterAll(void)
{
  if (MatRegisterAllCalled) return PETSC_SUCCESS;
  MatRegisterAllCalled = 1;
  PetscCall(MatRegisterRootName(MATAIJB200, MATSEQAIJB200, MATMPIAIJB200));
  PetscCall(MatRegisterRootName(MATAIJ, MATSEQAIJB200, MATMPIAIJB200));
  PetscCall(MatRegister(MATSEQAIJB200, MatCreate_SeqAIJB200));
  PetscCall(MatRegister(MATMPIAIJB200, MatCreate_MPIAIJB200));
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ interface */
PetscErrorCode MatCreate(MPI_Comm comm, Mat *A)
{
  PetscValidPointer(A, 2);
  PetscCall(PetscB200EnsureInit());
  Mat B = (Mat)calloc(1, sizeof(*B));
  PetscCheck(B, comm, PETSC_ERR_MEM, "out of memory");
  B->hdr.comm  = comm;
  B->hdr.refct = 1;
  B->m = B->n = B->M = B->N = -1;
  strcpy(B->defaultvectype, VECB200);
  *A = B;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N)
{
  PetscValidHeader(A, 1);
  PetscCheck(!(M > 0 && m > M), A->hdr.comm, PETSC_ERR_ARG_INCOMP, "Local row size %d cannot be larger than global row size %d", m, M);
  PetscCheck(!(N > 0 && n > N), A->hdr.comm, PETSC_ERR_ARG_INCOMP, "Local column size %d cannot be larger than global column size %d", n, N);
  A->m = m; A->n = n; A->M = M; A->N = N;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatSetUpLayout(Mat A)
{
  if (A->sizes_set) return PETSC_SUCCESS;
  PetscCheck((A->m >= 0 || A->M >= 0) && (A->n >= 0 || A->N >= 0), A->hdr.comm, PETSC_ERR_ORDER, "Must call MatSetSizes() first");
  PetscCall(PetscSplitOwnership(A->hdr.comm, &A->m, &A->M));
  PetscCall(PetscSplitOwnership(A->hdr.comm, &A->n, &A->N));
  int      size = PetscB200CommSize(A->hdr.comm), rank = PetscB200CommRank(A->hdr.comm);
  int64_t *all = (int64_t *)malloc(sizeof(int64_t) * (size_t)size), s = 0;
  PetscCall(PetscB200AllgatherInt64(A->hdr.comm, A->m, all));
  for (int r = 0; r < rank; r++) s += all[r];
  A->rstart = (PetscInt)s; A->rend = (PetscInt)(s + A->m);
  PetscCall(PetscB200AllgatherInt64(A->hdr.comm, A->n, all));
  s = 0;
  for (int r = 0; r < rank; r++) s += all[r];
  A->cstart = (PetscInt)s; A->cend = (PetscInt)(s + A->n);
  free(all);
  A->sizes_set = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetType(Mat mat, MatType matype)
{
  PetscErrorCode (*create)(Mat) = NULL;
  const char *name = matype;
  PetscValidHeader(mat, 1);
  PetscCall(MatRegisterAll());
  for (int i = 0; i < nroot; i++) /* matreg.c:128-138: root name -> seq/mpi flavour by communicator size */
    if (!strcmp(rootnames[i].r, matype)) name = PetscB200CommSize(mat->hdr.comm) == 1 ? rootnames[i].s : rootnames[i].m;
  if (!strcmp(mat->hdr.type_name, name)) return PETSC_SUCCESS;
  PetscCall(PetscFunctionListFind(MatList, name, (void **)&create));
  PetscCheck(create, mat->hdr.comm, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unknown Mat type given: %s", matype);
  PetscCheck(!mat->data, mat->hdr.comm, PETSC_ERR_SUP, "Cannot convert an existing %s matrix to %s", mat->hdr.type_name, name);
  PetscCall((*create)(mat));
  mat->type_set = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetFromOptions(Mat B)
{
  char type[64] = MATAIJB200;
  PetscCall(PetscOptionsGetString(NULL, B->hdr.prefix, "-mat_type", type, sizeof type, NULL));
  if (!strcmp(type, "seqaij") || !strcmp(type, "mpiaij") || !strcmp(type, "aijcusparse")) strcpy(type, MATAIJB200);
  PetscCall(MatSetType(B, type));
  PetscInt v[4] = {B->spmv_layout[0], B->spmv_layout[1], B->spmv_layout[2], B->spmv_layout[3]};
  PetscCall(PetscOptionsGetInt(NULL, B->hdr.prefix, "-mat_b200_spmv_lanes", &v[0], NULL));
  PetscCall(PetscOptionsGetInt(NULL, B->hdr.prefix, "-mat_b200_spmv_rows_per_tile", &v[1], NULL));
  PetscCall(PetscOptionsGetInt(NULL, B->hdr.prefix, "-mat_b200_spmv_stages", &v[2], NULL));
  PetscCall(PetscOptionsGetInt(NULL, B->hdr.prefix, "-mat_b200_spmv_ctas_per_sm", &v[3], NULL));
  for (int i = 0; i < 4; i++) B->spmv_layout[i] = v[i];
  {
    PetscBool ordered = (PetscBool)B->spmv_ordered;
    PetscCall(PetscOptionsGetBool(NULL, B->hdr.prefix, "-mat_b200_spmv_ordered", &ordered, NULL));
    B->spmv_ordered = ordered ? 1 : 0;
  }
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetType(Mat mat, MatType *type)
{
  *type = mat->hdr.type_name;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatPrep(Mat A)
{
  PetscValidHeader(A, 1);
  if (!A->type_set) PetscCall(MatSetType(A, MATAIJB200));
  PetscCall(MatSetUpLayout(A));
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetUp(Mat A) { return MatPrep(A); }
PetscErrorCode MatSeqAIJSetPreallocation(Mat B, PetscInt nz, const PetscInt nnz[])
{
  (void)nz; (void)nnz;
  return MatPrep(B); /* values are staged and compacted at assembly: nothing to preallocate */
}
PetscErrorCode MatMPIAIJSetPreallocation(Mat B, PetscInt d_nz, const PetscInt d_nnz[], PetscInt o_nz, const PetscInt o_nnz[])
{
  (void)d_nz; (void)d_nnz; (void)o_nz; (void)o_nnz;
  return MatPrep(B);
}
PetscErrorCode MatSetOption(Mat mat, MatOption op, PetscBool flg)
{
  (void)mat; (void)op; (void)flg;
  return PETSC_SUCCESS;
}

PetscErrorCode MatSetValues(Mat mat, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode addv)
{
  PetscCall(MatPrep(mat));
  PetscCheck(addv == INSERT_VALUES || addv == ADD_VALUES, mat->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "InsertMode must be INSERT_VALUES or ADD_VALUES");
  /* MatSetValues (matrix.c:1480): one mode between two assemblies */
  PetscCheck(!mat->stash_mode || mat->stash_mode == (int)addv, mat->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Cannot mix add values and insert values");
  /* this harness keeps no host copy of an assembled device matrix to merge new entries into (the reference merges into the
     existing rows, aij.c:420-560): refuse instead of silently replacing the matrix */
  PetscCheck(!mat->ever_assembled, mat->hdr.comm, PETSC_ERR_SUP, "MatSetValues on an already assembled matrix is not supported by the test harness (use MatSetValuesCOO or recreate the matrix)");
  mat->stash_mode = (int)addv;
  for (PetscInt i = 0; i < m; i++) {
    if (idxm[i] < 0) continue;
    PetscCheck(idxm[i] >= mat->rstart && idxm[i] < mat->rend, mat->hdr.comm, PETSC_ERR_SUP, "MatSetValues of off-process row %d (owned [%d,%d)) is not supported by the b200 matrix types", idxm[i], mat->rstart, mat->rend);
    for (PetscInt j = 0; j < n; j++) {
      if (idxn[j] < 0) continue;
      PetscCheck(idxn[j] < mat->N, mat->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Column too large: col %d max %d", idxn[j], mat->N - 1);
      if (mat->ncoo == mat->coocap) {
        mat->coocap = mat->coocap ? mat->coocap * 2 : 1024;
        mat->coo    = (COOEntry *)realloc(mat->coo, mat->coocap * sizeof(COOEntry));
        PetscCheck(mat->coo, mat->hdr.comm, PETSC_ERR_MEM, "out of memory");
      }
      mat->coo[mat->ncoo].row = idxm[i];
      mat->coo[mat->ncoo].col = idxn[j];
      mat->coo[mat->ncoo].v   = v[(size_t)i * n + j];
      mat->coo[mat->ncoo].seq = mat->ncoo;
      mat->ncoo++;
    }
  }
  mat->assembled = 0;
  return PETSC_SUCCESS;
}
static int coo_cmp(const void *a, const void *b)
{
  const COOEntry *x = (const COOEntry *)a, *y = (const COOEntry *)b;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0); /* arrival order: stable */
}
PetscErrorCode MatAssemblyBegin(Mat mat, MatAssemblyType type)
{
  (void)mat; (void)type;
  return PETSC_SUCCESS;
}
PetscErrorCode MatAssemblyEnd(Mat mat, MatAssemblyType type)
{
  PetscCall(MatPrep(mat));
  if (type == MAT_FLUSH_ASSEMBLY) return PETSC_SUCCESS;
  if (mat->ncoo || !mat->assembled) {
    {
      /* MatAssemblyEnd_SeqAIJ (aij.c:1085): compact the staged entries into sorted CSR rows, duplicates summed */
      size_t    k = 0, nent = mat->ncoo;
      PetscInt *ai = (PetscInt *)calloc((size_t)mat->m + 1, sizeof(PetscInt));
      /* stable within equal (row,col) through the arrival index: ADD_VALUES sums repeats in arrival order, INSERT_VALUES keeps
         the last one (MatSetValues_SeqAIJ overwrites, aij.c:470) */
      const int insert = mat->stash_mode == (int)INSERT_VALUES;
      qsort(mat->coo, nent, sizeof(COOEntry), coo_cmp);
      PetscInt *aj = (PetscInt *)malloc(sizeof(PetscInt) * (nent + 1));
      double   *aa = (double *)malloc(sizeof(double) * (nent + 1));
      for (size_t e = 0; e < nent; e++) {
        if (k && mat->coo[e].row == mat->coo[e - 1].row && mat->coo[e].col == mat->coo[e - 1].col) aa[k - 1] = insert ? mat->coo[e].v : aa[k - 1] + mat->coo[e].v;
        else {
          aj[k] = mat->coo[e].col;
          aa[k] = mat->coo[e].v;
          ai[mat->coo[e].row - mat->rstart + 1]++;
          k++;
        }
      }
      for (PetscInt r = 0; r < mat->m; r++) ai[r + 1] += ai[r];
      if (mat->ops.setcsr && (nent || !mat->assembled)) PetscCall((*mat->ops.setcsr)(mat, ai, aj, aa, 0));
      free(ai); free(aj); free(aa);
      free(mat->coo);
      mat->coo  = NULL;
      mat->ncoo = mat->coocap = 0;
      if (nent) mat->ever_assembled = 1;
      mat->stash_mode = 0;
    }
  }
  if (mat->ops.assemblyend) PetscCall((*mat->ops.assemblyend)(mat, type));
  mat->assembled = 1;
  mat->hdr.state++;
  return PETSC_SUCCESS;
}

PetscErrorCode MatSeqAIJSetPreallocationCSR(Mat B, const PetscInt i[], const PetscInt j[], const PetscScalar v[])
{
  PetscCall(MatPrep(B));
  PetscCheck(i[0] == 0, B->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "i[0] must be 0 it is %d", i[0]);
  PetscCall((*B->ops.setcsr)(B, i, j, v, 0));
  B->assembled = 1;
  B->hdr.state++;
  return PETSC_SUCCESS;
}
PetscErrorCode MatMPIAIJSetPreallocationCSR(Mat B, const PetscInt i[], const PetscInt j[], const PetscScalar v[]) { return MatSeqAIJSetPreallocationCSR(B, i, j, v); }
PetscErrorCode MatB200SetCSRDevice(Mat B, const PetscInt *d_i, const PetscInt *d_j, const PetscScalar *d_a)
{
  PetscCall(MatPrep(B));
  PetscCall((*B->ops.setcsr)(B, d_i, d_j, d_a, 1));
  B->assembled = 1;
  B->hdr.state++;
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreateSeqAIJWithArrays(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt i[], PetscInt j[], PetscScalar a[], Mat *mat)
{
  PetscCall(MatCreate(comm == PETSC_COMM_WORLD && PetscB200.size > 1 ? PETSC_COMM_SELF : comm, mat));
  PetscCall(MatSetSizes(*mat, m, n, m, n));
  PetscCall(MatSetType(*mat, MATSEQAIJB200));
  PetscCall(MatSeqAIJSetPreallocationCSR(*mat, i, j, a));
  return PETSC_SUCCESS;
}

PetscErrorCode MatGetSize(Mat mat, PetscInt *m, PetscInt *n)
{
  PetscCall(MatPrep(mat));
  if (m) *m = mat->M;
  if (n) *n = mat->N;
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetLocalSize(Mat mat, PetscInt *m, PetscInt *n)
{
  PetscCall(MatPrep(mat));
  if (m) *m = mat->m;
  if (n) *n = mat->n;
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetOwnershipRange(Mat mat, PetscInt *m, PetscInt *n)
{
  PetscCall(MatPrep(mat));
  if (m) *m = mat->rstart;
  if (n) *n = mat->rend;
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreateVecs(Mat mat, Vec *right, Vec *left)
{
  PetscCall(MatPrep(mat));
  if (right) {
    PetscCall(VecCreate(mat->hdr.comm, right));
    PetscCall(VecSetSizes(*right, mat->n, mat->N));
    PetscCall(VecSetType(*right, mat->defaultvectype)); /* matrix.c:10069 */
  }
  if (left) {
    PetscCall(VecCreate(mat->hdr.comm, left));
    PetscCall(VecSetSizes(*left, mat->m, mat->M));
    PetscCall(VecSetType(*left, mat->defaultvectype));
  }
  return PETSC_SUCCESS;
}
#define MatCheckAssembled(A) PetscCheck((A)->assembled, (A)->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Not for unassembled matrix")
PetscErrorCode MatMult(Mat mat, Vec x, Vec y)
{
  PetscValidHeader(mat, 1);
  PetscValidHeader(x, 2);
  PetscValidHeader(y, 3);
  MatCheckAssembled(mat);
  PetscCheck(x != y, mat->hdr.comm, PETSC_ERR_ARG_IDN, "x and y must be different vectors");
  PetscInt nx, ny;
  PetscCall(VecGetLocalSize(x, &nx));
  PetscCall(VecGetLocalSize(y, &ny));
  PetscCheck(mat->n == nx, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat mat,Vec x: local dim %d %d", mat->n, nx); /* matrix.c:2708-2712 */
  PetscCheck(mat->m == ny, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat mat,Vec y: local dim %d %d", mat->m, ny);
  PetscCall((*mat->ops.mult)(mat, x, y));
  return PETSC_SUCCESS;
}
PetscErrorCode MatMultAdd(Mat mat, Vec v1, Vec v2, Vec v3)
{
  PetscValidHeader(mat, 1);
  MatCheckAssembled(mat);
  PetscCheck(v1 != v3, mat->hdr.comm, PETSC_ERR_ARG_IDN, "v1 and v3 must be different vectors");
  PetscInt n1, n2, n3;
  PetscCall(VecGetLocalSize(v1, &n1));
  PetscCall(VecGetLocalSize(v2, &n2));
  PetscCall(VecGetLocalSize(v3, &n3));
  PetscCheck(mat->n == n1 && mat->m == n2 && mat->m == n3, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat/Vec local dimensions do not conform");
  PetscCall((*mat->ops.multadd)(mat, v1, v2, v3));
  return PETSC_SUCCESS;
}
PetscErrorCode MatMultTranspose(Mat mat, Vec x, Vec y)
{
  PetscValidHeader(mat, 1);
  PetscValidHeader(x, 2);
  PetscValidHeader(y, 3);
  MatCheckAssembled(mat);
  PetscCheck(x != y, mat->hdr.comm, PETSC_ERR_ARG_IDN, "x and y must be different vectors");
  PetscCheck(mat->ops.multtranspose, mat->hdr.comm, PETSC_ERR_SUP, "No method multtranspose for Mat of type %s", mat->hdr.type_name);
  PetscInt nx, ny;
  PetscCall(VecGetLocalSize(x, &nx));
  PetscCall(VecGetLocalSize(y, &ny));
  PetscCheck(mat->m == nx, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat mat,Vec x: local dim %d %d", mat->m, nx);
  PetscCheck(mat->n == ny, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat mat,Vec y: local dim %d %d", mat->n, ny);
  PetscCall((*mat->ops.multtranspose)(mat, x, y));
  return PETSC_SUCCESS;
}
PetscErrorCode MatMultTransposeAdd(Mat mat, Vec v1, Vec v2, Vec v3)
{
  PetscValidHeader(mat, 1);
  MatCheckAssembled(mat);
  PetscCheck(v1 != v3, mat->hdr.comm, PETSC_ERR_ARG_IDN, "v1 and v3 must be different vectors");
  PetscCheck(mat->ops.multtransposeadd, mat->hdr.comm, PETSC_ERR_SUP, "No method multtransposeadd for Mat of type %s", mat->hdr.type_name);
  PetscInt n1, n2, n3;
  PetscCall(VecGetLocalSize(v1, &n1));
  PetscCall(VecGetLocalSize(v2, &n2));
  PetscCall(VecGetLocalSize(v3, &n3));
  PetscCheck(mat->m == n1 && mat->n == n2 && mat->n == n3, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "Mat/Vec local dimensions do not conform");
  PetscCall((*mat->ops.multtransposeadd)(mat, v1, v2, v3));
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetPreallocationCOO(Mat A, PetscCount ncoo, PetscInt coo_i[], PetscInt coo_j[])
{
  PetscValidHeader(A, 1);
  PetscCheck(ncoo >= 0, A->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "ncoo %lld overflowed or negative", (long long)ncoo);
  PetscCall(MatPrep(A));
  PetscCheck(A->ops.setpreallocationcoo, A->hdr.comm, PETSC_ERR_SUP, "No method setpreallocationcoo for Mat of type %s", A->hdr.type_name);
  PetscCall((*A->ops.setpreallocationcoo)(A, ncoo, coo_i, coo_j));
  A->coo_n     = ncoo;
  A->assembled = 1; /* matrix.c MatSetPreallocationCOO: the pattern is final, values are zero */
  A->hdr.state++;
  return PETSC_SUCCESS;
}
PetscErrorCode MatSetValuesCOO(Mat A, const PetscScalar coo_v[], InsertMode imode)
{
  PetscValidHeader(A, 1);
  PetscCheck(imode == INSERT_VALUES || imode == ADD_VALUES, A->hdr.comm, PETSC_ERR_ARG_WRONG, "InsertMode must be INSERT_VALUES or ADD_VALUES");
  PetscCheck(A->ops.setvaluescoo, A->hdr.comm, PETSC_ERR_SUP, "No method setvaluescoo for Mat of type %s", A->hdr.type_name);
  PetscCall((*A->ops.setvaluescoo)(A, coo_v, imode));
  A->assembled = 1;
  A->hdr.state++;
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetDiagonal(Mat mat, Vec v)
{
  PetscValidHeader(mat, 1);
  MatCheckAssembled(mat);
  PetscCall((*mat->ops.getdiagonal)(mat, v));
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetDiagonalBlock(Mat A, Mat *a)
{
  PetscValidHeader(A, 1);
  MatCheckAssembled(A);
  PetscCheck(A->ops.getdiagonalblock, A->hdr.comm, PETSC_ERR_SUP, "No support for this operation for this matrix type");
  return (*A->ops.getdiagonalblock)(A, a);
}
PetscErrorCode MatDestroy(Mat *A)
{
  if (!A || !*A) return PETSC_SUCCESS;
  if (--(*A)->hdr.refct > 0) {
    *A = NULL;
    return PETSC_SUCCESS;
  }
  if ((*A)->ops.destroy) PetscCall((*(*A)->ops.destroy)(*A));
  free((*A)->coo);
  free(*A);
  *A = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode MatGetInfo(Mat mat, MatInfoType flag, MatInfo *info)
{
  PetscValidHeader(mat, 1);
  PetscValidPointer(info, 3);
  memset(info, 0, sizeof(*info));
  info->block_size = 1.0;
  double nz        = 0.0;
  if (!strcmp(mat->hdr.type_name, MATSEQAIJB200) && mat->data) nz = (double)((Mat_SeqAIJB200 *)mat->data)->nz;
  else if (!strcmp(mat->hdr.type_name, MATMPIAIJB200) && mat->data) {
    Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data; /* MatGetInfo_MPIAIJ: diag + off-diag blocks */
    if (a->A && a->A->data) nz += (double)((Mat_SeqAIJB200 *)a->A->data)->nz;
    if (a->B && a->B->data) nz += (double)((Mat_SeqAIJB200 *)a->B->data)->nz;
  }
  if (flag != MAT_LOCAL) PetscCall(PetscB200AllreduceHost(mat->hdr.comm, &nz, 1, flag == MAT_GLOBAL_MAX ? 1 : 0));
  info->nz_used = info->nz_allocated = nz;
  info->memory                       = nz * 12.0;
  info->assemblies                   = (double)mat->assembled;
  return PETSC_SUCCESS;
}

/* -mat_b200_spmv_ordered: MatMult / MatMultAdd / MatMultTranspose row sums in the reference's left-to-right FMA-free order for
   every lane count (bit-identical to MatMult_SeqAIJ); default off = fastest association (equal to rounding) */
PetscErrorCode MatB200SetSpMVOrdered(Mat A, PetscBool ordered)
{
  A->spmv_ordered = ordered ? 1 : 0;
  if (!strcmp(A->hdr.type_name, MATSEQAIJB200) && A->data) {
    Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
    if (a->plan) PetscCallB200(b200CsrPlanSetSummation(a->plan, ordered ? 0 : 1));
    if (a->T) {
      b200CsrPlan tp;
      PetscCallB200(b200CsrTransposeGetPlan(a->T, &tp));
      PetscCallB200(b200CsrPlanSetSummation(tp, ordered ? 0 : 1));
    }
  } else if (!strcmp(A->hdr.type_name, MATMPIAIJB200) && A->data) {
    Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)A->data;
    if (a->A) PetscCall(MatB200SetSpMVOrdered(a->A, ordered));
  }
  return PETSC_SUCCESS;
}
PetscErrorCode MatB200SetSpMVLayout(Mat A, PetscInt lanes, PetscInt rows, PetscInt stages, PetscInt ctas)
{
  A->spmv_layout[0] = lanes; A->spmv_layout[1] = rows; A->spmv_layout[2] = stages; A->spmv_layout[3] = ctas;
  if (!strcmp(A->hdr.type_name, MATSEQAIJB200) && A->data) {
    Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
    if (a->plan) PetscCallB200(b200CsrPlanSetLayout(a->plan, lanes, rows, stages, ctas));
    if (a->T) {
      b200CsrPlan tp;
      PetscCallB200(b200CsrTransposeGetPlan(a->T, &tp));
      PetscCallB200(b200CsrPlanSetLayout(tp, lanes, rows, stages, ctas));
    }
  } else if (!strcmp(A->hdr.type_name, MATMPIAIJB200) && A->data) {
    Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)A->data;
    if (a->A) PetscCall(MatB200SetSpMVLayout(a->A, lanes, rows, stages, ctas));
  }
  return PETSC_SUCCESS;
}

/* ================================================================== seqaijb200 */
static PetscErrorCode MatSeqAIJB200_Free(Mat_SeqAIJB200 *a)
{
  if (a->plan) b200CsrPlanDestroy(a->plan);
  a->plan = NULL;
  if (a->T) b200CsrTransposeDestroy(a->T);
  a->T = NULL;
  if (a->coo) b200CooPlanDestroy(a->coo);
  a->coo = NULL;
  PetscCallB200(b200Free(H, a->d_i)); PetscCallB200(b200Free(H, a->d_j)); PetscCallB200(b200Free(H, a->d_a));
  PetscCallB200(b200Free(H, a->d_cr_i)); PetscCallB200(b200Free(H, a->d_cr_rindex));
  a->d_i = a->d_j = a->d_cr_i = a->d_cr_rindex = NULL;
  a->d_a = NULL;
  free(a->h_i); free(a->h_j); free(a->h_a);
  a->h_i = a->h_j = NULL;
  a->h_a = NULL;
  return PETSC_SUCCESS;
}

/* CSR with LOCAL column indices; host or device source */
static PetscErrorCode MatSetCSR_SeqAIJB200(Mat A, const PetscInt *ai, const PetscInt *aj, const PetscScalar *aa, int on_device)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  PetscInt        m = A->m;
  int64_t         nz;
  PetscCall(MatSeqAIJB200_Free(a));
  a->m = m; a->n = A->n;
  if (on_device) {
    int last = 0;
    if (m) PetscCallB200(b200MemcpyDtoH(H, &last, ai + m, sizeof(int)));
    nz = last;
  } else {
    nz = m ? ai[m] : 0;
  }
  a->nz = nz;
  if (on_device == 2) { /* adopt: the arrays were allocated with b200Malloc by the caller and now belong to the matrix */
    a->d_i = (int *)ai; a->d_j = (int *)aj; a->d_a = (double *)aa;
  } else {
    PetscCallB200(b200Malloc(H, (void **)&a->d_i, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200Malloc(H, (void **)&a->d_j, sizeof(int) * (size_t)(nz + 1)));
    PetscCallB200(b200Malloc(H, (void **)&a->d_a, sizeof(double) * (size_t)(nz + 1)));
  }
  if (on_device == 2) {
  } else if (on_device) {
    PetscCallB200(b200MemcpyDtoD(H, a->d_i, ai, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyDtoD(H, a->d_j, aj, sizeof(int) * (size_t)nz));
    PetscCallB200(b200MemcpyDtoD(H, a->d_a, aa, sizeof(double) * (size_t)nz));
  } else {
    if (m == 0) {
      int zero = 0;
      PetscCallB200(b200MemcpyHtoD(H, a->d_i, &zero, sizeof(int)));
    } else PetscCallB200(b200MemcpyHtoD(H, a->d_i, ai, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyHtoD(H, a->d_j, aj, sizeof(int) * (size_t)nz));
    PetscCallB200(b200MemcpyHtoD(H, a->d_a, aa, sizeof(double) * (size_t)nz));
  }
  { /* MatAssemblyEnd_SeqAIJ invariants (sorted, in-range columns), checked where the data now lives */
    int bad = -1, kind = 0;
    PetscCallB200(b200CsrValidate(H, m, A->n, a->d_i, a->d_j, &bad, &kind));
    PetscCheck(kind != 1, A->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Column out of range [0,%d) in row %d", A->n, bad);
    PetscCheck(kind != 2, A->hdr.comm, PETSC_ERR_ARG_WRONG, "Row %d: column indices must be strictly increasing", bad);
    PetscCheck(kind != 3, A->hdr.comm, PETSC_ERR_ARG_WRONG, "Row %d: row pointer is decreasing", bad);
  }
  PetscCallB200(b200CsrPlanCreate(H, m, A->n, nz, a->d_i, a->d_j, &a->plan));
  if (A->spmv_layout[0] || A->spmv_layout[1] || A->spmv_layout[2] || A->spmv_layout[3]) PetscCallB200(b200CsrPlanSetLayout(a->plan, A->spmv_layout[0], A->spmv_layout[1], A->spmv_layout[2], A->spmv_layout[3]));
  if (A->spmv_ordered) PetscCallB200(b200CsrPlanSetSummation(a->plan, 0));
  {
    int cnt = 0;
    PetscCallB200(b200CsrCountNonemptyRows(H, m, a->d_i, &cnt));
    a->nonzerorowcnt = cnt;
  }
  /* MatCheckCompressedRow (src/mat/utils/compressedrow.c): use the compressed view when > 60% of the rows are empty */
  a->cr_use = 0;
  if (m && (double)(m - a->nonzerorowcnt) > 0.6 * m) {
    int *hi = (int *)malloc(sizeof(int) * ((size_t)m + 1));
    if (on_device) PetscCallB200(b200MemcpyDtoH(H, hi, a->d_i, sizeof(int) * ((size_t)m + 1)));
    else memcpy(hi, ai, sizeof(int) * ((size_t)m + 1));
    int  nr = a->nonzerorowcnt, k = 0;
    int *cri = (int *)malloc(sizeof(int) * ((size_t)nr + 1)), *rid = (int *)malloc(sizeof(int) * ((size_t)nr + 1));
    cri[0] = 0;
    for (PetscInt r = 0; r < m; r++)
      if (hi[r + 1] > hi[r]) {
        rid[k]     = r;
        cri[k + 1] = hi[r + 1];
        k++;
      }
    a->cr_nrows = nr;
    PetscCallB200(b200Malloc(H, (void **)&a->d_cr_i, sizeof(int) * ((size_t)nr + 1)));
    PetscCallB200(b200Malloc(H, (void **)&a->d_cr_rindex, sizeof(int) * ((size_t)nr + 1)));
    PetscCallB200(b200MemcpyHtoD(H, a->d_cr_i, cri, sizeof(int) * ((size_t)nr + 1)));
    PetscCallB200(b200MemcpyHtoD(H, a->d_cr_rindex, rid, sizeof(int) * (size_t)nr));
    free(hi); free(cri); free(rid);
    a->cr_use = 1;
  }
  return PETSC_SUCCESS;
}

static PetscErrorCode MatMult_SeqAIJB200(Mat A, Vec x, Vec y)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  const double   *dx;
  double         *dy;
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(y, &dy));
  PetscCallB200(b200CsrSpMV(H, a->plan, a->d_a, dx, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMultAdd_SeqAIJB200(Mat A, Vec x, Vec y, Vec z)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  const double   *dx, *dy;
  double         *dz;
  PetscCall(VecB200GetArrayRead(x, &dx));
  if (a->cr_use) {
    /* compressed-row branch (aij.c:1626-1640): z = y on the empty rows, then update only the rows that own entries */
    if (z != y) PetscCall(VecCopy(y, z));
    PetscCall(VecB200GetArray(z, &dz));
    PetscCallB200(b200CsrSpMVAddCompressed(H, a->cr_nrows, a->d_cr_i, a->d_cr_rindex, a->d_j, a->d_a, dx, dz, dz));
    return PETSC_SUCCESS;
  }
  PetscCall(VecB200GetArrayRead(y, &dy));
  if (z == y) PetscCall(VecB200GetArray(z, &dz));
  else PetscCall(VecB200GetArrayWrite(z, &dz));
  PetscCallB200(b200CsrSpMVAdd(H, a->plan, a->d_a, dx, dy, dz));
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMultJacobi_SeqAIJB200(Mat A, Vec x, Vec dinv, Vec w)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  const double   *dx, *dd;
  double         *dw;
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayRead(dinv, &dd));
  PetscCall(VecB200GetArrayWrite(w, &dw));
  PetscCallB200(b200CsrSpMVJacobi(H, a->plan, a->d_a, dx, dd, dw, NULL));
  return PETSC_SUCCESS;
}
/* MatMultTranspose[Add]_SeqAIJ (aij.c:1383-1440) through the explicit transposed pattern */
static PetscErrorCode MatTransposeSync_SeqAIJB200(Mat A)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  if (!a->T) {
    PetscCallB200(b200CsrTransposeCreate(H, a->m, a->n, a->nz, a->d_i, a->d_j, &a->T));
    a->T_state = -1;
    if (A->spmv_ordered || A->spmv_layout[0] || A->spmv_layout[1] || A->spmv_layout[2] || A->spmv_layout[3]) { /* -mat_b200_spmv_* apply to A^T too */
      b200CsrPlan tp;
      PetscCallB200(b200CsrTransposeGetPlan(a->T, &tp));
      PetscCallB200(b200CsrPlanSetLayout(tp, A->spmv_layout[0], A->spmv_layout[1], A->spmv_layout[2], A->spmv_layout[3]));
      if (A->spmv_ordered) PetscCallB200(b200CsrPlanSetSummation(tp, 0));
    }
  }
  if (a->T_state != A->hdr.state) {
    PetscCallB200(b200CsrTransposeSetValues(H, a->T, a->d_a));
    a->T_state = A->hdr.state;
  }
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMultTranspose_SeqAIJB200(Mat A, Vec x, Vec y)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  const double   *dx;
  double         *dy;
  PetscCall(MatTransposeSync_SeqAIJB200(A));
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(y, &dy));
  PetscCallB200(b200CsrTransposeSpMV(H, a->T, dx, NULL, dy));
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMultTransposeAdd_SeqAIJB200(Mat A, Vec x, Vec y, Vec z)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  const double   *dx, *dy;
  double         *dz;
  PetscCall(MatTransposeSync_SeqAIJB200(A));
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayRead(y, &dy));
  if (z == y) PetscCall(VecB200GetArray(z, &dz));
  else PetscCall(VecB200GetArrayWrite(z, &dz));
  PetscCallB200(b200CsrTransposeSpMV(H, a->T, dx, dy, dz));
  return PETSC_SUCCESS;
}
/* MatSetPreallocationCOO_SeqAIJ (aij.c:4524): the sort / unique / CSR construction run on the device */
static PetscErrorCode MatSetPreallocationCOO_SeqAIJB200(Mat A, PetscCount ncoo, const PetscInt coo_i[], const PetscInt coo_j[])
{
  Mat_SeqAIJB200 *a;
  int             dev_i = 0, dev_j = 0;
  int            *d_ci = NULL, *d_cj = NULL;
  double         *d_zero = NULL;
  b200CooPlan     plan = NULL;
  const int      *d_rp, *d_cx;
  int64_t         nnz, atot;
  PetscCallB200(b200PointerIsDevice(coo_i, &dev_i));
  PetscCallB200(b200PointerIsDevice(coo_j, &dev_j));
  if (!dev_i && ncoo) {
    PetscCallB200(b200Malloc(H, (void **)&d_ci, sizeof(int) * (size_t)ncoo));
    PetscCallB200(b200MemcpyHtoD(H, d_ci, coo_i, sizeof(int) * (size_t)ncoo));
  }
  if (!dev_j && ncoo) {
    PetscCallB200(b200Malloc(H, (void **)&d_cj, sizeof(int) * (size_t)ncoo));
    PetscCallB200(b200MemcpyHtoD(H, d_cj, coo_j, sizeof(int) * (size_t)ncoo));
  }
  {
    int rc = b200CooPlanCreate(H, A->m, A->n, ncoo, dev_i ? coo_i : d_ci, dev_j ? coo_j : d_cj, &plan);
    PetscCallB200(b200Free(H, d_ci));
    PetscCallB200(b200Free(H, d_cj));
    PetscCheck(rc != B200_ERR_ARG_OUTOFRANGE, A->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "%s", b200GetLastErrorString());
    PetscCallB200(rc);
  }
  PetscCallB200(b200CooPlanGetCsr(plan, &nnz, &atot, &d_rp, &d_cx));
  PetscCallB200(b200Malloc(H, (void **)&d_zero, sizeof(double) * (size_t)(nnz + 1)));
  PetscCallB200(b200VecSet(H, nnz, 0.0, d_zero));
  PetscCall(MatSetCSR_SeqAIJB200(A, d_rp, d_cx, d_zero, 1)); /* copies the pattern; frees a previous COO plan */
  PetscCallB200(b200Free(H, d_zero));
  a      = (Mat_SeqAIJB200 *)A->data;
  a->coo = plan;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatSetValuesCOO_SeqAIJB200(Mat A, const PetscScalar v[], InsertMode imode)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  int             dev = 0;
  double         *d_v = NULL;
  int64_t         nnz, atot;
  PetscCheck(a->coo, A->hdr.comm, PETSC_ERR_PLIB, "Not found MatCOOStruct on this matrix"); /* aij.c:4721 */
  PetscCallB200(b200CooPlanGetCsr(a->coo, &nnz, &atot, NULL, NULL));
  PetscCallB200(b200PointerIsDevice(v, &dev));
  if (!dev && v) {
    int64_t n = A->coo_n;
    PetscCallB200(b200Malloc(H, (void **)&d_v, sizeof(double) * (size_t)(n + 1)));
    PetscCallB200(b200MemcpyHtoD(H, d_v, v, sizeof(double) * (size_t)n));
  }
  PetscCallB200(b200CooSetValues(H, a->coo, dev ? v : d_v, imode == INSERT_VALUES, a->d_a));
  PetscCallB200(b200Free(H, d_v));
  free(a->h_i); free(a->h_j); free(a->h_a); /* host copies handed out by MatSeqAIJGetCSRHost are stale now */
  a->h_i = a->h_j = NULL;
  a->h_a = NULL;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatGetDiagonal_SeqAIJB200(Mat A, Vec v)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  double         *dv;
  PetscInt        n;
  PetscCall(VecGetLocalSize(v, &n));
  PetscCheck(n == A->m, A->hdr.comm, PETSC_ERR_ARG_SIZ, "Nonconforming matrix and vector");
  PetscCall(VecB200GetArrayWrite(v, &dv));
  PetscCallB200(b200CsrGetDiagonal(H, A->m, a->d_i, a->d_j, a->d_a, dv, NULL));
  return PETSC_SUCCESS;
}
/* seq matrices: the diagonal block is the matrix itself */
static PetscErrorCode MatGetDiagonalBlock_SeqAIJB200(Mat A, Mat *blk)
{
  *blk = A;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatDestroy_SeqAIJB200(Mat A)
{
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  if (a) {
    PetscCall(MatSeqAIJB200_Free(a));
    free(a);
  }
  A->data = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreate_SeqAIJB200(Mat A)
{
  PetscCheck(PetscB200CommSize(A->hdr.comm) == 1, A->hdr.comm, PETSC_ERR_ARG_WRONG, "Comm must be of size 1");
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)calloc(1, sizeof(*a));
  PetscCheck(a, A->hdr.comm, PETSC_ERR_MEM, "out of memory");
  A->data            = a;
  A->ops.mult        = MatMult_SeqAIJB200;
  A->ops.multadd     = MatMultAdd_SeqAIJB200;
  A->ops.getdiagonal = MatGetDiagonal_SeqAIJB200;
  A->ops.destroy     = MatDestroy_SeqAIJB200;
  A->ops.getdiagonalblock = MatGetDiagonalBlock_SeqAIJB200;
  A->ops.setcsr      = MatSetCSR_SeqAIJB200;
  A->ops.multjacobi  = MatMultJacobi_SeqAIJB200;
  A->ops.multtranspose       = MatMultTranspose_SeqAIJB200;
  A->ops.multtransposeadd    = MatMultTransposeAdd_SeqAIJB200;
  A->ops.setpreallocationcoo = MatSetPreallocationCOO_SeqAIJB200;
  A->ops.setvaluescoo        = MatSetValuesCOO_SeqAIJB200;
  strcpy(A->hdr.type_name, MATSEQAIJB200);
  strcpy(A->defaultvectype, VECB200); /* aijcusparse.cu:2820 analogue */
  return PETSC_SUCCESS;
}
PetscErrorCode MatSeqAIJGetCSRHost(Mat A, PetscInt *m, const PetscInt **i, const PetscInt **j, const PetscScalar **aa)
{
  PetscCheck(!strcmp(A->hdr.type_name, MATSEQAIJB200), A->hdr.comm, PETSC_ERR_ARG_WRONG, "Not a seqaijb200 matrix");
  Mat_SeqAIJB200 *a = (Mat_SeqAIJB200 *)A->data;
  if (!a->h_i) {
    a->h_i = (int *)malloc(sizeof(int) * ((size_t)a->m + 1));
    a->h_j = (int *)malloc(sizeof(int) * (size_t)(a->nz + 1));
    a->h_a = (double *)malloc(sizeof(double) * (size_t)(a->nz + 1));
    PetscCallB200(b200MemcpyDtoH(H, a->h_i, a->d_i, sizeof(int) * ((size_t)a->m + 1)));
    PetscCallB200(b200MemcpyDtoH(H, a->h_j, a->d_j, sizeof(int) * (size_t)a->nz));
    PetscCallB200(b200MemcpyDtoH(H, a->h_a, a->d_a, sizeof(double) * (size_t)a->nz));
  }
  if (m) *m = a->m;
  if (i) *i = a->h_i;
  if (j) *j = a->h_j;
  if (aa) *aa = a->h_a;
  return PETSC_SUCCESS;
}

/* ================================================================== mpiaijb200 */
static int cmp_int(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* Pure host restatement of the column split + MatSetUpMultiply_MPIAIJ numbering (mmaij.c:25-61): exported so the
   world_size-2 CPU tests can check garray / renumbering without a GPU.  Outputs are malloc'ed. */
PetscErrorCode PetscB200MPIAIJSplit(PetscInt m, PetscInt cstart, PetscInt cend, const PetscInt *ai, const PetscInt *aj, const PetscScalar *aa, PetscInt **Ai, PetscInt **Aj, PetscScalar **Aa, PetscInt **Bi, PetscInt **Bj, PetscScalar **Ba, PetscInt **garray, PetscInt *ec)
{
  size_t nzA = 0, nzB = 0;
  for (PetscInt r = 0; r < m; r++)
    for (PetscInt k = ai[r]; k < ai[r + 1]; k++) {
      if (aj[k] >= cstart && aj[k] < cend) nzA++;
      else nzB++;
    }
  *Ai = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)m + 1)); *Bi = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)m + 1));
  *Aj = (PetscInt *)malloc(sizeof(PetscInt) * (nzA + 1)); *Bj = (PetscInt *)malloc(sizeof(PetscInt) * (nzB + 1));
  *Aa = (PetscScalar *)malloc(sizeof(PetscScalar) * (nzA + 1)); *Ba = (PetscScalar *)malloc(sizeof(PetscScalar) * (nzB + 1));
  PetscInt *g = (PetscInt *)malloc(sizeof(PetscInt) * (nzB + 1));
  size_t    kb = 0;
  for (PetscInt r = 0; r < m; r++)
    for (PetscInt k = ai[r]; k < ai[r + 1]; k++)
      if (aj[k] < cstart || aj[k] >= cend) g[kb++] = aj[k];
  PetscInt n = 0;
  if (nzB) { /* mmaij.c:51 PetscSortInt + uniq */
    qsort(g, nzB, sizeof(PetscInt), cmp_int);
    n = 1;
    for (size_t k = 1; k < nzB; k++)
      if (g[k] != g[n - 1]) g[n++] = g[k];
  }
  size_t ka = 0;
  kb        = 0;
  (*Ai)[0] = (*Bi)[0] = 0;
  for (PetscInt r = 0; r < m; r++) {
    for (PetscInt k = ai[r]; k < ai[r + 1]; k++) {
      PetscInt c = aj[k];
      if (c >= cstart && c < cend) {
        (*Aj)[ka] = c - cstart;
        (*Aa)[ka++] = aa[k];
      } else { /* mmaij.c:55-61: position in garray */
        PetscInt lo = 0, hi = n - 1;
        while (lo < hi) {
          PetscInt mid = (lo + hi) / 2;
          if (g[mid] < c) lo = mid + 1;
          else hi = mid;
        }
        (*Bj)[kb] = lo;
        (*Ba)[kb++] = aa[k];
      }
    }
    (*Ai)[r + 1] = (PetscInt)ka;
    (*Bi)[r + 1] = (PetscInt)kb;
  }
  *garray = g;
  *ec     = n;
  return PETSC_SUCCESS;
}

/* Receive side of the scatter (mmaij.c:108-117 + PetscSFSetGraph): garray is sorted, so the entries owned by rank p
   are one contiguous range of lvec.  ranges[size+1] = column ownership.  Outputs sized [size]. */
PetscErrorCode PetscB200HaloPlanRecv(PetscInt ec, const PetscInt *garray, int size, const int64_t *ranges, int *recv_counts, int *recv_offsets)
{
  PetscInt k = 0;
  for (int p = 0; p < size; p++) {
    recv_offsets[p] = k;
    while (k < ec && garray[k] < ranges[p + 1]) k++;
    recv_counts[p] = k - recv_offsets[p];
  }
  PetscCheck(k == ec, 0, PETSC_ERR_PLIB, "garray holds columns outside the global range");
  return PETSC_SUCCESS;
}

static PetscErrorCode MatSetUpMultiply_MPIAIJB200(Mat mat)
{
  Mat_MPIAIJB200 *a    = (Mat_MPIAIJB200 *)mat->data;
  int             size = PetscB200CommSize(mat->hdr.comm), rank = PetscB200CommRank(mat->hdr.comm);
  int            *rc = (int *)calloc((size_t)size, sizeof(int)), *ro = (int *)calloc((size_t)size, sizeof(int));
  int            *sc = (int *)calloc((size_t)size, sizeof(int));
  int64_t        *n_all = (int64_t *)malloc(sizeof(int64_t) * (size_t)size);
  free(a->ranges);
  a->ranges = (int64_t *)malloc(sizeof(int64_t) * ((size_t)size + 1));
  PetscCall(PetscB200AllgatherInt64(mat->hdr.comm, mat->n, n_all));
  a->ranges[0] = 0;
  for (int p = 0; p < size; p++) a->ranges[p + 1] = a->ranges[p] + n_all[p];
  free(n_all);
  PetscCall(PetscB200HaloPlanRecv(a->ec, a->garray, size, a->ranges, rc, ro));
  PetscCheck(rc[rank] == 0, mat->hdr.comm, PETSC_ERR_PLIB, "off-diagonal block references locally owned columns");
  /* tell every owner which of its entries I need (local index on the owner) */
  int *req = (int *)malloc(sizeof(int) * ((size_t)a->ec + 1)), *need = NULL;
  for (PetscInt k = 0; k < a->ec; k++) {
    int p = 0;
    while (a->garray[k] >= a->ranges[p + 1]) p++;
    req[k] = (int)(a->garray[k] - a->ranges[p]);
  }
  PetscCall(PetscB200AlltoallvInt(mat->hdr.comm, rc, req, sc, &need));
  /* peers = ranks I exchange with in either direction */
  int  np = 0, *peers = (int *)malloc(sizeof(int) * (size_t)size), *psc = (int *)malloc(sizeof(int) * (size_t)size), *prc = (int *)malloc(sizeof(int) * (size_t)size), *pro = (int *)malloc(sizeof(int) * (size_t)size);
  int *sidx, ns = 0, off = 0;
  {
    size_t tot = 0;
    for (int p = 0; p < size; p++) tot += (size_t)sc[p];
    sidx = (int *)malloc(sizeof(int) * (tot + 1));
  }
  for (int p = 0; p < size; p++) {
    if (p != rank && (sc[p] || rc[p])) {
      peers[np] = p; psc[np] = sc[p]; prc[np] = rc[p]; pro[np] = ro[p];
      for (int k = 0; k < sc[p]; k++) {
        PetscCheck(need[off + k] >= 0 && need[off + k] < mat->n, mat->hdr.comm, PETSC_ERR_PLIB, "peer %d requested entry %d outside my range", p, need[off + k]);
        sidx[ns++] = need[off + k];
      }
      np++;
    }
    off += sc[p];
  }
  if (a->Mvctx) b200HaloDestroy(a->Mvctx);
  PetscCallB200(b200HaloCreate(H, np, peers, psc, sidx, prc, pro, &a->Mvctx));
  free(rc); free(ro); free(sc); free(req); free(need); free(peers); free(psc); free(prc); free(pro); free(sidx);
  /* lvec (mmaij.c:103) */
  PetscCall(VecDestroy(&a->lvec));
  PetscCall(VecCreate(PETSC_COMM_SELF, &a->lvec));
  PetscCall(VecSetSizes(a->lvec, a->ec, a->ec));
  PetscCall(VecSetType(a->lvec, VECSEQB200));
  return PETSC_SUCCESS;
}

/* local rows, GLOBAL columns */
static PetscErrorCode MatSetCSR_MPIAIJB200(Mat mat, const PetscInt *ai, const PetscInt *aj, const PetscScalar *aa, int on_device)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  PetscInt        m  = mat->m;
  if (on_device) {
    /* device-resident input: the column split runs on the device (b200CsrSplitColumns); only the small off-diagonal
       block visits the host, where garray is built exactly as mmaij.c:25-61 does */
    int    *dAi, *dAj, *dBi, *dBj;
    double *dAa, *dBa;
    int64_t nzA, nzB;
    PetscCallB200(b200CsrSplitColumns(H, m, ai, aj, aa, mat->cstart, mat->cend, &dAi, &dAj, &dAa, &nzA, &dBi, &dBj, &dBa, &nzB));
    PetscInt    *hBi = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)m + 1)), *hBj = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)nzB + 1));
    PetscScalar *hBa = (PetscScalar *)malloc(sizeof(PetscScalar) * ((size_t)nzB + 1));
    PetscCallB200(b200MemcpyDtoH(H, hBi, dBi, sizeof(PetscInt) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyDtoH(H, hBj, dBj, sizeof(PetscInt) * (size_t)nzB));
    PetscCallB200(b200MemcpyDtoH(H, hBa, dBa, sizeof(PetscScalar) * (size_t)nzB));
    PetscCallB200(b200Free(H, dBi)); PetscCallB200(b200Free(H, dBj)); PetscCallB200(b200Free(H, dBa));
    PetscInt *g = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)nzB + 1)), ec = 0;
    memcpy(g, hBj, sizeof(PetscInt) * (size_t)nzB);
    if (nzB) {
      qsort(g, (size_t)nzB, sizeof(PetscInt), cmp_int);
      ec = 1;
      for (int64_t k = 1; k < nzB; k++)
        if (g[k] != g[ec - 1]) g[ec++] = g[k];
    }
    for (int64_t k = 0; k < nzB; k++) { /* mmaij.c:55-61 */
      PetscInt c = hBj[k], lo = 0, hi2 = ec - 1;
      PetscCheck(c >= 0 && c < mat->N, mat->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Column %d out of range", c);
      while (lo < hi2) {
        PetscInt mid = (lo + hi2) / 2;
        if (g[mid] < c) lo = mid + 1;
        else hi2 = mid;
      }
      hBj[k] = lo;
    }
    free(a->garray);
    a->garray = g;
    a->ec     = ec;
    PetscCall(MatDestroy(&a->A));
    PetscCall(MatDestroy(&a->B));
    PetscCall(MatCreate(PETSC_COMM_SELF, &a->A));
    PetscCall(MatSetSizes(a->A, m, mat->n, m, mat->n));
    PetscCall(MatSetType(a->A, MATSEQAIJB200));
    for (int i = 0; i < 4; i++) a->A->spmv_layout[i] = mat->spmv_layout[i];
  a->A->spmv_ordered = mat->spmv_ordered;
    a->A->spmv_ordered = mat->spmv_ordered;
    PetscCall(MatSetUp(a->A));
    PetscCall((*a->A->ops.setcsr)(a->A, dAi, dAj, dAa, 2)); /* adopts the device arrays */
    a->A->assembled = 1;
    PetscCall(MatCreate(PETSC_COMM_SELF, &a->B));
    PetscCall(MatSetSizes(a->B, m, ec, m, ec));
    PetscCall(MatSetType(a->B, MATSEQAIJB200));
    PetscCall(MatSeqAIJSetPreallocationCSR(a->B, hBi, hBj, hBa));
    free(hBi); free(hBj); free(hBa);
    PetscCall(MatSetUpMultiply_MPIAIJB200(mat));
    return PETSC_SUCCESS;
  }
  if (m && ai[m] > (1 << 16)) {
    /* large host CSR: one host->device copy of the three arrays, then the device split above -- the host loop below walks
       every entry twice and costs seconds per 10^9 nonzeros, the copy runs at PCIe speed */
    int           *d_i = NULL, *d_j = NULL;
    double        *d_a = NULL;
    const size_t   nz  = (size_t)ai[m];
    PetscErrorCode ierr;
    PetscCallB200(b200Malloc(H, (void **)&d_i, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200Malloc(H, (void **)&d_j, sizeof(int) * nz));
    PetscCallB200(b200Malloc(H, (void **)&d_a, sizeof(double) * nz));
    PetscCallB200(b200MemcpyHtoD(H, d_i, ai, sizeof(int) * ((size_t)m + 1)));
    PetscCallB200(b200MemcpyHtoD(H, d_j, aj, sizeof(int) * nz));
    PetscCallB200(b200MemcpyHtoD(H, d_a, aa, sizeof(double) * nz));
    ierr = MatSetCSR_MPIAIJB200(mat, d_i, d_j, d_a, 1);
    PetscCallB200(b200Free(H, d_i)); PetscCallB200(b200Free(H, d_j)); PetscCallB200(b200Free(H, d_a));
    return ierr;
  }
  PetscInt    *Ai, *Aj, *Bi, *Bj, *g, ec;
  PetscScalar *Aa, *Ba;
  for (PetscInt r = 0; r < m; r++)
    for (PetscInt k = ai[r]; k < ai[r + 1]; k++) PetscCheck(aj[k] >= 0 && aj[k] < mat->N, mat->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Column %d out of range in local row %d", aj[k], r);
  PetscCall(PetscB200MPIAIJSplit(m, mat->cstart, mat->cend, ai, aj, aa, &Ai, &Aj, &Aa, &Bi, &Bj, &Ba, &g, &ec));
  free(a->garray);
  a->garray = g;
  a->ec     = ec;
  PetscCall(MatDestroy(&a->A));
  PetscCall(MatDestroy(&a->B));
  PetscCall(MatCreate(PETSC_COMM_SELF, &a->A));
  PetscCall(MatSetSizes(a->A, m, mat->n, m, mat->n));
  PetscCall(MatSetType(a->A, MATSEQAIJB200)); /* the MatMPIAIJSetPreallocation_C hook of mpiaijcupm.hpp:326-377 */
  for (int i = 0; i < 4; i++) a->A->spmv_layout[i] = mat->spmv_layout[i];
  a->A->spmv_ordered = mat->spmv_ordered;
  PetscCall(MatSeqAIJSetPreallocationCSR(a->A, Ai, Aj, Aa));
  PetscCall(MatCreate(PETSC_COMM_SELF, &a->B));
  PetscCall(MatSetSizes(a->B, m, ec, m, ec));
  PetscCall(MatSetType(a->B, MATSEQAIJB200));
  PetscCall(MatSeqAIJSetPreallocationCSR(a->B, Bi, Bj, Ba));
  free(Ai); free(Aj); free(Aa); free(Bi); free(Bj); free(Ba);
  PetscCall(MatSetUpMultiply_MPIAIJB200(mat));
  return PETSC_SUCCESS;
}

static PetscErrorCode MatMult_MPIAIJB200(Mat mat, Vec x, Vec y)
{
  /* MatMult_MPIAIJ (mpiaij.c:1047-1061): scatter begin / diag block / scatter end / off-diag multadd.
     The exchange runs on the halo stream (pack kernel + grouped ncclSend/Recv) while the diagonal-block SpMV runs on
     the main stream. */
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  const double   *dx;
  double         *dl;
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(a->lvec, &dl));
  PetscCallB200(b200HaloBegin(H, a->Mvctx, dx, dl));
  PetscCall((*a->A->ops.mult)(a->A, x, y));
  PetscCallB200(b200HaloEnd(H, a->Mvctx));
  PetscCall((*a->B->ops.multadd)(a->B, a->lvec, y, y));
  return PETSC_SUCCESS;
}
/* PCApplyBAorAB with PCJACOBI on the row-partitioned matrix, fused: w = dinv .* (A_d x + B_o lvec) without the intermediate y
   (mpiaij.c:1047-1061 + jacobi.c:354).  The diagonal block runs its fused kernel while the halo travels; the (few) rows that
   own off-diagonal entries are then redone with both blocks in the reference's order. */
static PetscErrorCode MatMultJacobi_MPIAIJB200(Mat mat, Vec x, Vec dinv, Vec w)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  Mat_SeqAIJB200 *A = (Mat_SeqAIJB200 *)a->A->data, *B = (Mat_SeqAIJB200 *)a->B->data;
  const double   *dx, *dd, *dlr;
  double         *dl, *dw;
  if (!a->A->ops.multjacobi || (B->nz && !B->cr_use)) { /* off-diagonal block not in compressed-row form: unfused */
    Vec work;
    PetscCall(VecDuplicate(w, &work));
    PetscCall(MatMult(mat, x, work));
    PetscCall(VecPointwiseMult(w, work, dinv));
    PetscCall(VecDestroy(&work));
    return PETSC_SUCCESS;
  }
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(a->lvec, &dl));
  PetscCallB200(b200HaloBegin(H, a->Mvctx, dx, dl));
  PetscCall((*a->A->ops.multjacobi)(a->A, x, dinv, w));
  PetscCallB200(b200HaloEnd(H, a->Mvctx));
  if (B->nz) {
    PetscCall(VecB200GetArrayRead(a->lvec, &dlr));
    PetscCall(VecB200GetArrayRead(dinv, &dd));
    PetscCall(VecB200GetArray(w, &dw));
    PetscCallB200(b200CsrSpMVAddCompressedJacobi(H, B->cr_nrows, B->d_cr_i, B->d_cr_rindex, B->d_j, B->d_a, dlr, A->d_i, A->d_j, A->d_a, dx, dd, dw));
  }
  return PETSC_SUCCESS;
}
static PetscErrorCode MatMultAdd_MPIAIJB200(Mat mat, Vec x, Vec y, Vec z)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data; /* mpiaij.c:1072-1084 */
  const double   *dx;
  double         *dl;
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(a->lvec, &dl));
  PetscCallB200(b200HaloBegin(H, a->Mvctx, dx, dl));
  PetscCall((*a->A->ops.multadd)(a->A, x, y, z));
  PetscCallB200(b200HaloEnd(H, a->Mvctx));
  PetscCall((*a->B->ops.multadd)(a->B, a->lvec, z, z));
  return PETSC_SUCCESS;
}
static PetscErrorCode MatGetDiagonal_MPIAIJB200(Mat mat, Vec v)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data; /* mpiaij.c:1158: diagonal of the diagonal block */
  PetscCheck(mat->rstart == mat->cstart && mat->rend == mat->cend, mat->hdr.comm, PETSC_ERR_ARG_SIZ, "row partition must equal col partition");
  return (*a->A->ops.getdiagonal)(a->A, v);
}
static PetscErrorCode MatGetDiagonalBlock_MPIAIJB200(Mat mat, Mat *blk)
{
  *blk = ((Mat_MPIAIJB200 *)mat->data)->A;
  return PETSC_SUCCESS;
}
static PetscErrorCode MatDestroy_MPIAIJB200(Mat mat)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)mat->data;
  if (a) {
    PetscCall(MatDestroy(&a->A));
    PetscCall(MatDestroy(&a->B));
    PetscCall(VecDestroy(&a->lvec));
    if (a->Mvctx) b200HaloDestroy(a->Mvctx);
    free(a->garray); free(a->ranges);
    free(a);
  }
  mat->data = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode MatCreate_MPIAIJB200(Mat A)
{
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)calloc(1, sizeof(*a));
  PetscCheck(a, A->hdr.comm, PETSC_ERR_MEM, "out of memory");
  A->data                 = a;
  A->ops.mult             = MatMult_MPIAIJB200;
  A->ops.multadd          = MatMultAdd_MPIAIJB200;
  A->ops.multjacobi       = MatMultJacobi_MPIAIJB200;
  A->ops.getdiagonal      = MatGetDiagonal_MPIAIJB200;
  A->ops.getdiagonalblock = MatGetDiagonalBlock_MPIAIJB200;
  A->ops.destroy          = MatDestroy_MPIAIJB200;
  A->ops.setcsr           = MatSetCSR_MPIAIJB200;
  strcpy(A->hdr.type_name, MATMPIAIJB200);
  strcpy(A->defaultvectype, VECB200);
  return PETSC_SUCCESS;
}
PetscErrorCode MatMPIAIJGetSeqAIJ(Mat A, Mat *Ad, Mat *Ao, const PetscInt *colmap[])
{
  PetscCheck(!strcmp(A->hdr.type_name, MATMPIAIJB200), A->hdr.comm, PETSC_ERR_ARG_WRONG, "This function requires a MATMPIAIJ matrix as input");
  Mat_MPIAIJB200 *a = (Mat_MPIAIJB200 *)A->data;
  if (Ad) *Ad = a->A;
  if (Ao) *Ao = a->B;
  if (colmap) *colmap = a->garray;
  return PETSC_SUCCESS;
}
