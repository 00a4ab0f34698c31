/* pc.c -- PC interface (src/ksp/pc/interface/precon.c) and PCNONE / PCJACOBI / PCBJACOBI / PCILU(0) for the b200 types
   (src/ksp/pc/impls/{jacobi/jacobi.c, bjacobi/bjacobi.c, factor/ilu/ilu.c}). */
#include "hostimpl.h"

#define H (PetscB200.h)
static PetscFunctionList PCList = NULL;
static int               PCRegisterAllCalled = 0;

PetscErrorCode PCRegister(const char sname[], PetscErrorCode (*function)(PC)) { return PetscFunctionListAdd(&PCList, sname, (void *)function); }
static PetscErrorCode PCRegisterAll(void)
{
  if (PCRegisterAllCalled) return PETSC_SUCCESS;
  PCRegisterAllCalled = 1;
  PetscCall(PCRegister(PCNONE, PCCreate_None));
  PetscCall(PCRegister(PCJACOBI, PCCreate_Jacobi));
  PetscCall(PCRegister(PCBJACOBI, PCCreate_BJacobi));
  PetscCall(PCRegister(PCILU, PCCreate_ILU));
  return PETSC_SUCCESS;
}

PetscErrorCode PCCreate(MPI_Comm comm, PC *newpc)
{
  PetscValidPointer(newpc, 2);
  PetscCall(PetscB200EnsureInit());
  PC pc = (PC)calloc(1, sizeof(*pc));
  PetscCheck(pc, comm, PETSC_ERR_MEM, "out of memory");
  pc->hdr.comm  = comm;
  pc->hdr.refct = 1;
  *newpc        = pc;
  return PETSC_SUCCESS;
}
static PetscErrorCode PCReset_Private(PC pc)
{
  if (pc->ops.reset) PetscCall((*pc->ops.reset)(pc));
  pc->setupcalled = 0;
  return PETSC_SUCCESS;
}
PetscErrorCode PCSetType(PC pc, PCType type)
{
  PetscErrorCode (*create)(PC) = NULL;
  PetscValidHeader(pc, 1);
  PetscCall(PCRegisterAll());
  if (!strcmp(pc->hdr.type_name, type)) return PETSC_SUCCESS;
  PetscCall(PetscFunctionListFind(PCList, type, (void **)&create));
  PetscCheck(create, pc->hdr.comm, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unable to find requested PC type %s", type);
  if (pc->type_set) {
    PetscCall(PCReset_Private(pc));
    if (pc->ops.destroy) PetscCall((*pc->ops.destroy)(pc));
    memset(&pc->ops, 0, sizeof pc->ops);
    pc->data = NULL;
  }
  PetscCall((*create)(pc));
  strncpy(pc->hdr.type_name, type, sizeof pc->hdr.type_name - 1);
  pc->type_set = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode PCGetType(PC pc, PCType *type)
{
  *type = pc->hdr.type_name;
  return PETSC_SUCCESS;
}
static PetscErrorCode PCSetDefaultType(PC pc)
{
  /* PCGetDefaultType_Private (precon.c:33-70): one process and a factorable matrix -> ILU; otherwise block Jacobi */
  if (pc->type_set) return PETSC_SUCCESS;
  if (PetscB200CommSize(pc->hdr.comm) == 1) return PCSetType(pc, PCILU);
  return PCSetType(pc, PCBJACOBI);
}
PetscErrorCode PCSetFromOptions(PC pc)
{
  char      type[64];
  PetscBool set = PETSC_FALSE;
  PetscCall(PetscOptionsGetString(NULL, pc->hdr.prefix, "-pc_type", type, sizeof type, &set));
  if (set) PetscCall(PCSetType(pc, type));
  else PetscCall(PCSetDefaultType(pc));
  if (pc->ops.setfromoptions) PetscCall((*pc->ops.setfromoptions)(pc));
  return PETSC_SUCCESS;
}
PetscErrorCode PCSetOperators(PC pc, Mat Amat, Mat Pmat)
{
  PetscValidHeader(pc, 1);
  if (Amat) Amat->hdr.refct++;
  if (Pmat) Pmat->hdr.refct++;
  PetscCall(MatDestroy(&pc->mat));
  PetscCall(MatDestroy(&pc->pmat));
  pc->mat  = Amat;
  pc->pmat = Pmat;
  if (pc->setupcalled) PetscCall(PCReset_Private(pc));
  return PETSC_SUCCESS;
}
PetscErrorCode PCSetUp(PC pc)
{
  PetscValidHeader(pc, 1);
  PetscCheck(pc->mat, pc->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Matrix must be set first");
  /* PCSetUp (precon.c:1080-1110) compares the operator's object state with the one recorded at the last set-up and rebuilds the
     preconditioner when the matrix changed (MatSetValuesCOO, MatB200SetCSRDevice, re-assembly all bump hdr.state): stale
     ILU(0) factors or a stale 1/diag would otherwise be combined with the new A */
  if (pc->setupcalled && pc->matstate == (int64_t)pc->pmat->hdr.state) return PETSC_SUCCESS;
  if (pc->setupcalled) PetscCall(PCReset_Private(pc)); /* full rebuild: releases the old factors / diagonal / sub-block Mat */
  PetscCall(PCSetDefaultType(pc));
  if (pc->ops.setup) PetscCall((*pc->ops.setup)(pc));
  pc->setupcalled = 1;
  pc->matstate    = (int64_t)pc->pmat->hdr.state;
  return PETSC_SUCCESS;
}
PetscErrorCode PCApply(PC pc, Vec x, Vec y)
{
  PetscValidHeader(pc, 1);
  PetscCheck(x != y, pc->hdr.comm, PETSC_ERR_ARG_IDN, "x and y must be different vectors"); /* precon.c:532 */
  PetscCall(PCSetUp(pc));
  PetscCall((*pc->ops.apply)(pc, x, y));
  return PETSC_SUCCESS;
}
PetscErrorCode PCApplyBAorAB(PC pc, int side, Vec x, Vec y, Vec work)
{
  /* precon.c:810-865, PC_LEFT branch: y = B A x, through ops->applyBA when the PC provides the fusion */
  PetscValidHeader(pc, 1);
  PetscCheck(side == 0, pc->hdr.comm, PETSC_ERR_SUP, "only left preconditioning is implemented by the b200 host mirror");
  PetscCheck(x != y, pc->hdr.comm, PETSC_ERR_ARG_IDN, "x and y must be different vectors");
  PetscCall(PCSetUp(pc));
  if (pc->ops.applyBA) return (*pc->ops.applyBA)(pc, side, x, y, work);
  PetscCall(MatMult(pc->mat, x, work));
  PetscCall(PCApply(pc, work, y));
  return PETSC_SUCCESS;
}
PetscErrorCode PCDestroy(PC *pc)
{
  if (!pc || !*pc) return PETSC_SUCCESS;
  if (--(*pc)->hdr.refct > 0) {
    *pc = NULL;
    return PETSC_SUCCESS;
  }
  PetscCall(PCReset_Private(*pc));
  if ((*pc)->ops.destroy) PetscCall((*(*pc)->ops.destroy)(*pc));
  PetscCall(MatDestroy(&(*pc)->mat));
  PetscCall(MatDestroy(&(*pc)->pmat));
  free(*pc);
  *pc = NULL;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ none */
static PetscErrorCode PCApply_None(PC pc, Vec x, Vec y)
{
  (void)pc;
  return VecCopy(x, y);
}
PetscErrorCode PCCreate_None(PC pc)
{
  pc->ops.apply = PCApply_None;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ jacobi (jacobi.c:172-270, 354-362) */
typedef struct {
  Vec diag; /* holds 1/diag(A) */
  int fuse; /* -pc_jacobi_b200_fuse: use the fused SpMV+Jacobi kernel in applyBA */
} PC_Jacobi;
static PetscErrorCode PCSetUp_Jacobi(PC pc)
{
  PC_Jacobi *j = (PC_Jacobi *)pc->data;
  double    *d;
  int        nzero = 0;
  PetscInt   n;
  PetscCall(VecDestroy(&j->diag));
  PetscCall(MatCreateVecs(pc->pmat, &j->diag, NULL));
  PetscCall(MatGetDiagonal(pc->pmat, j->diag));
  PetscCall(VecGetLocalSize(j->diag, &n));
  PetscCall(VecB200GetArray(j->diag, &d));
  /* VecReciprocal + the zero-diagonal fix-up loop of jacobi.c:253-266 in one kernel */
  PetscCallB200(b200JacobiInvertDiagonal(H, n, d, d, &nzero));
  return PETSC_SUCCESS;
}
static PetscErrorCode PCApply_Jacobi(PC pc, Vec x, Vec y)
{
  PC_Jacobi *j = (PC_Jacobi *)pc->data;
  return VecPointwiseMult(y, x, j->diag); /* jacobi.c:359 */
}
static PetscErrorCode PCApplyBA_Jacobi(PC pc, int side, Vec x, Vec y, Vec work)
{
  PC_Jacobi *j = (PC_Jacobi *)pc->data;
  (void)side;
  if (j->fuse && pc->mat->ops.multjacobi && pc->mat == pc->pmat) return (*pc->mat->ops.multjacobi)(pc->mat, x, j->diag, y);
  PetscCall(MatMult(pc->mat, x, work));
  return PCApply_Jacobi(pc, work, y);
}
static PetscErrorCode PCSetFromOptions_Jacobi(PC pc)
{
  PC_Jacobi *j = (PC_Jacobi *)pc->data;
  PetscBool  b = (PetscBool)j->fuse;
  PetscCall(PetscOptionsGetBool(NULL, pc->hdr.prefix, "-pc_jacobi_b200_fuse", &b, NULL));
  j->fuse = b;
  return PETSC_SUCCESS;
}
static PetscErrorCode PCReset_Jacobi(PC pc)
{
  PC_Jacobi *j = (PC_Jacobi *)pc->data;
  return VecDestroy(&j->diag);
}
static PetscErrorCode PCDestroy_Jacobi(PC pc)
{
  free(pc->data);
  pc->data = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PCCreate_Jacobi(PC pc)
{
  PC_Jacobi *j = (PC_Jacobi *)calloc(1, sizeof(*j));
  PetscCheck(j, pc->hdr.comm, PETSC_ERR_MEM, "out of memory");
  j->fuse                = 1;
  pc->data               = j;
  pc->ops.setup          = PCSetUp_Jacobi;
  pc->ops.apply          = PCApply_Jacobi;
  pc->ops.applyBA        = PCApplyBA_Jacobi;
  pc->ops.setfromoptions = PCSetFromOptions_Jacobi;
  pc->ops.reset          = PCReset_Jacobi;
  pc->ops.destroy        = PCDestroy_Jacobi;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ ilu(0) (ilu.c:68-192 + aijfact.c) */
typedef struct {
  b200IluPlan plan;
  double      zeropivot, shiftamount;
  int         nshift;
} PC_ILU;
static PetscErrorCode PCSetUp_ILU(PC pc)
{
  PC_ILU          *ilu = (PC_ILU *)pc->data;
  Mat              P   = pc->pmat;
  PetscInt         m;
  const PetscInt  *hi, *hj;
  PetscCheck(!strcmp(P->hdr.type_name, MATSEQAIJB200), pc->hdr.comm, PETSC_ERR_SUP, "PCILU requires a sequential matrix (use -pc_type bjacobi in parallel); got %s", P->hdr.type_name); /* MatGetFactor: matrix.c:5017 */
  PetscCheck(P->m == P->n, pc->hdr.comm, PETSC_ERR_ARG_WRONG, "matrix must be square");
  PetscCall(MatSeqAIJGetCSRHost(P, &m, &hi, &hj, NULL));
  if (ilu->plan) PetscCallB200(b200Ilu0Destroy(ilu->plan));
  ilu->plan = NULL;
  PetscCallB200(b200Ilu0Symbolic(H, m, hi, hj, &ilu->plan));                                    /* MatILUFactorSymbolic_SeqAIJ_ilu0 */
  PetscCallB200(b200Ilu0Numeric(H, ilu->plan, ((Mat_SeqAIJB200 *)P->data)->d_a, ilu->zeropivot, ilu->shiftamount, &ilu->nshift)); /* MatLUFactorNumeric_SeqAIJ */
  return PETSC_SUCCESS;
}
static PetscErrorCode PCApply_ILU(PC pc, Vec x, Vec y)
{
  PC_ILU       *ilu = (PC_ILU *)pc->data;
  const double *dx;
  double       *dy;
  PetscCall(VecB200GetArrayRead(x, &dx));
  PetscCall(VecB200GetArrayWrite(y, &dy));
  PetscCallB200(b200Ilu0Solve(H, ilu->plan, dx, dy)); /* MatSolve_SeqAIJ_NaturalOrdering */
  return PETSC_SUCCESS;
}
static PetscErrorCode PCSetFromOptions_ILU(PC pc)
{
  PC_ILU  *ilu = (PC_ILU *)pc->data;
  PetscInt levels = 0;
  PetscCall(PetscOptionsGetInt(NULL, pc->hdr.prefix, "-pc_factor_levels", &levels, NULL));
  PetscCheck(levels == 0, pc->hdr.comm, PETSC_ERR_SUP, "only ILU(0) is implemented on the device (got -pc_factor_levels %d)", levels);
  PetscCall(PetscOptionsGetReal(NULL, pc->hdr.prefix, "-pc_factor_zeropivot", &ilu->zeropivot, NULL));
  PetscCall(PetscOptionsGetReal(NULL, pc->hdr.prefix, "-pc_factor_shift_amount", &ilu->shiftamount, NULL));
  return PETSC_SUCCESS;
}
static PetscErrorCode PCReset_ILU(PC pc)
{
  PC_ILU *ilu = (PC_ILU *)pc->data;
  if (ilu->plan) PetscCallB200(b200Ilu0Destroy(ilu->plan));
  ilu->plan = NULL;
  return PETSC_SUCCESS;
}
static PetscErrorCode PCDestroy_ILU(PC pc)
{
  free(pc->data);
  pc->data = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PCCreate_ILU(PC pc)
{
  PC_ILU *ilu = (PC_ILU *)calloc(1, sizeof(*ilu));
  PetscCheck(ilu, pc->hdr.comm, PETSC_ERR_MEM, "out of memory");
  /* PCCreate_ILU (ilu.c): shifttype NONZERO, shiftamount 100 eps, zeropivot 100 eps */
  ilu->zeropivot = ilu->shiftamount = 100.0 * 2.220446049250313e-16;
  pc->data               = ilu;
  pc->ops.setup          = PCSetUp_ILU;
  pc->ops.apply          = PCApply_ILU;
  pc->ops.setfromoptions = PCSetFromOptions_ILU;
  pc->ops.reset          = PCReset_ILU;
  pc->ops.destroy        = PCDestroy_ILU;
  return PETSC_SUCCESS;
}

/* ------------------------------------------------------------------ bjacobi, one block per rank (bjacobi.c:11-130, 579-598, 730-808) */
typedef struct {
  PC  sub;   /* the sub-KSP is preonly: its solve is exactly one PCApply of the sub-PC (bjacobi.c:752-756) */
  Vec x, y;  /* local-array aliases (bjacobi.c:787-789) */
} PC_BJacobi;
static PetscErrorCode PCSetUp_BJacobi(PC pc)
{
  PC_BJacobi *bj = (PC_BJacobi *)pc->data;
  Mat         blk;
  char        type[64] = PCILU; /* default sub-PC of a seqaij block */
  PetscInt    blocks = -1, m;
  PetscCall(PetscOptionsGetInt(NULL, pc->hdr.prefix, "-pc_bjacobi_blocks", &blocks, NULL));
  PetscCheck(blocks < 0 || blocks == PetscB200CommSize(pc->hdr.comm), pc->hdr.comm, PETSC_ERR_SUP, "the b200 block Jacobi supports exactly one block per rank");
  if (!strcmp(pc->pmat->hdr.type_name, MATSEQAIJB200)) blk = pc->pmat;
  else PetscCall(MatGetDiagonalBlock(pc->pmat, &blk)); /* bjacobi.c:117-123 */
  PetscCall(PCDestroy(&bj->sub));
  PetscCall(PCCreate(PETSC_COMM_SELF, &bj->sub));
  snprintf(bj->sub->hdr.prefix, sizeof bj->sub->hdr.prefix, "%ssub_", pc->hdr.prefix);
  PetscCall(PetscOptionsGetString(NULL, bj->sub->hdr.prefix, "-pc_type", type, sizeof type, NULL));
  PetscCall(PCSetType(bj->sub, type));
  PetscCall(PCSetFromOptions(bj->sub));
  PetscCall(PCSetOperators(bj->sub, blk, blk));
  PetscCall(PCSetUp(bj->sub));
  PetscCall(MatGetLocalSize(blk, &m, NULL));
  PetscCall(VecDestroy(&bj->x));
  PetscCall(VecDestroy(&bj->y));
  PetscCall(VecCreate(PETSC_COMM_SELF, &bj->x));
  PetscCall(VecSetSizes(bj->x, m, m));
  PetscCall(VecSetType(bj->x, VECSEQB200));
  PetscCall(VecDuplicate(bj->x, &bj->y));
  return PETSC_SUCCESS;
}
static PetscErrorCode PCApply_BJacobi(PC pc, Vec x, Vec y)
{
  PC_BJacobi *bj = (PC_BJacobi *)pc->data;
  PetscCall(VecGetLocalVectorRead(x, bj->x)); /* bjacobi.c:586-593 */
  PetscCall(VecGetLocalVector(y, bj->y));
  PetscCall(PCApply(bj->sub, bj->x, bj->y));
  PetscCall(VecRestoreLocalVectorRead(x, bj->x));
  PetscCall(VecRestoreLocalVector(y, bj->y));
  return PETSC_SUCCESS;
}
static PetscErrorCode PCReset_BJacobi(PC pc)
{
  PC_BJacobi *bj = (PC_BJacobi *)pc->data;
  PetscCall(PCDestroy(&bj->sub));
  PetscCall(VecDestroy(&bj->x));
  PetscCall(VecDestroy(&bj->y));
  return PETSC_SUCCESS;
}
static PetscErrorCode PCDestroy_BJacobi(PC pc)
{
  free(pc->data);
  pc->data = NULL;
  return PETSC_SUCCESS;
}
PetscErrorCode PCCreate_BJacobi(PC pc)
{
  PC_BJacobi *bj = (PC_BJacobi *)calloc(1, sizeof(*bj));
  PetscCheck(bj, pc->hdr.comm, PETSC_ERR_MEM, "out of memory");
  pc->data        = bj;
  pc->ops.setup   = PCSetUp_BJacobi;
  pc->ops.apply   = PCApply_BJacobi;
  pc->ops.reset   = PCReset_BJacobi;
  pc->ops.destroy = PCDestroy_BJacobi;
  return PETSC_SUCCESS;
}
