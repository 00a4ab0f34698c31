/* ksp.c -- KSP interface (src/ksp/ksp/interface/{itcreate.c,itfunc.c,iterativ.c,itres.c}) and the Krylov methods of
   the hot path: KSPGMRES (impls/gmres/{gmres.c,borthog2.c}), KSPCG (impls/cg/cg.c), KSPPREONLY.  These are the CALLERS
   of the device kernels: the op sequence per iteration is the reference's, nothing is reordered. */
#include "hostimpl.h"

static PetscFunctionList KSPList = NULL;
static int               KSPRegisterAllCalled = 0;

PetscErrorCode KSPRegister(const char sname[], PetscErrorCode (*function)(KSP)) { return PetscFunctionListAdd(&KSPList, sname, (void *)function); }
static PetscErrorCode KSPRegisterAll(void)
{
  if (KSPRegisterAllCalled) return PETSC_SUCCESS;
  KSPRegisterAllCalled = 1;
  PetscCall(KSPRegister(KSPGMRES, KSPCreate_GMRES));
  PetscCall(KSPRegister(KSPCG, KSPCreate_CG));
  PetscCall(KSPRegister(KSPPIPECG, KSPCreate_PIPECG));
  PetscCall(KSPRegister(KSPPREONLY, KSPCreate_PREONLY));
  return PETSC_SUCCESS;
}

PetscErrorCode KSPCreate(MPI_Comm comm, KSP *inksp)
{
  PetscValidPointer(inksp, 2);
  PetscCall(PetscB200EnsureInit());
  KSP ksp = (KSP)calloc(1, sizeof(*ksp));
  PetscCheck(ksp, comm, PETSC_ERR_MEM, "out of memory");
  ksp->hdr.comm  = comm;
  ksp->hdr.refct = 1;
  /* itcreate.c:809-814 */
  ksp->max_it     = 10000;
  ksp->rtol       = 1.e-5;
  ksp->abstol     = 1.e-50;
  ksp->divtol     = 1.e4;
  ksp->guess_zero = 1;
  *inksp          = ksp;
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPReset_Private(KSP ksp)
{
  if (ksp->ops.reset) PetscCall((*ksp->ops.reset)(ksp));
  if (ksp->work) PetscCall(VecDestroyVecs(ksp->nwork, &ksp->work));
  ksp->nwork       = 0;
  ksp->setupcalled = 0;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetType(KSP ksp, KSPType type)
{
  PetscErrorCode (*create)(KSP) = NULL;
  PetscValidHeader(ksp, 1);
  PetscCall(KSPRegisterAll());
  if (!strcmp(ksp->hdr.type_name, type)) return PETSC_SUCCESS;
  PetscCall(PetscFunctionListFind(KSPList, type, (void **)&create));
  PetscCheck(create, ksp->hdr.comm, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unable to find requested KSP type %s", type);
  if (ksp->type_set) {
    PetscCall(KSPReset_Private(ksp));
    if (ksp->ops.destroy) PetscCall((*ksp->ops.destroy)(ksp));
    memset(&ksp->ops, 0, sizeof ksp->ops);
    ksp->data = NULL;
  }
  PetscCall((*create)(ksp));
  strncpy(ksp->hdr.type_name, type, sizeof ksp->hdr.type_name - 1);
  ksp->type_set = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetType(KSP ksp, KSPType *type)
{
  *type = ksp->hdr.type_name;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetPC(KSP ksp, PC *pc)
{
  if (!ksp->pc) PetscCall(PCCreate(ksp->hdr.comm, &ksp->pc));
  *pc = ksp->pc;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetOperators(KSP ksp, Mat Amat, Mat Pmat)
{
  PC pc;
  PetscCall(KSPGetPC(ksp, &pc));
  PetscCall(PCSetOperators(pc, Amat, Pmat));
  if (ksp->setupcalled) PetscCall(KSPReset_Private(ksp));
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
  /* itfunc.c KSPSetTolerances: PETSC_CURRENT / PETSC_DEFAULT(-2) keep the current value */
  if (rtol != (PetscReal)PETSC_CURRENT && rtol != (PetscReal)PETSC_DEFAULT) {
    PetscCheck(rtol >= 0.0 && rtol < 1.0, ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Relative tolerance %g must be non-negative and less than 1.0", rtol);
    ksp->rtol = rtol;
  }
  if (abstol != (PetscReal)PETSC_CURRENT && abstol != (PetscReal)PETSC_DEFAULT) {
    PetscCheck(abstol >= 0.0, ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Absolute tolerance %g must be non-negative", abstol);
    ksp->abstol = abstol;
  }
  if (dtol != (PetscReal)PETSC_CURRENT && dtol != (PetscReal)PETSC_DEFAULT) {
    PetscCheck(dtol >= 0.0, ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Divergence tolerance %g must be larger than 1.0", dtol);
    ksp->divtol = dtol;
  }
  if (maxits != PETSC_CURRENT && maxits != PETSC_DEFAULT) {
    PetscCheck(maxits >= 0, ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Maximum number of iterations %d must be non-negative", maxits);
    ksp->max_it = maxits;
  }
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetInitialGuessNonzero(KSP ksp, PetscBool flg)
{
  ksp->guess_zero = !flg;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetFromOptions(KSP ksp)
{
  char      type[64] = KSPGMRES; /* itcreate.c: default KSPGMRES */
  PetscBool set;
  PC        pc;
  if (ksp->type_set) strcpy(type, ksp->hdr.type_name);
  PetscCall(PetscOptionsGetString(NULL, ksp->hdr.prefix, "-ksp_type", type, sizeof type, NULL));
  PetscCall(KSPSetType(ksp, type));
  PetscCall(PetscOptionsGetReal(NULL, ksp->hdr.prefix, "-ksp_rtol", &ksp->rtol, NULL));
  PetscCall(PetscOptionsGetReal(NULL, ksp->hdr.prefix, "-ksp_atol", &ksp->abstol, NULL));
  PetscCall(PetscOptionsGetReal(NULL, ksp->hdr.prefix, "-ksp_divtol", &ksp->divtol, NULL));
  PetscCall(PetscOptionsGetInt(NULL, ksp->hdr.prefix, "-ksp_max_it", &ksp->max_it, NULL));
  {
    PetscBool nz = (PetscBool)!ksp->guess_zero;
    PetscCall(PetscOptionsGetBool(NULL, ksp->hdr.prefix, "-ksp_initial_guess_nonzero", &nz, NULL));
    ksp->guess_zero = !nz;
  }
  {
    PetscBool mon = PETSC_FALSE;
    PetscCall(PetscOptionsGetBool(NULL, ksp->hdr.prefix, "-ksp_monitor", &mon, &set));
    if (set) ksp->monitor_stdout = mon;
  }
  if (ksp->ops.setfromoptions) PetscCall((*ksp->ops.setfromoptions)(ksp));
  PetscCall(KSPGetPC(ksp, &pc));
  strncpy(pc->hdr.prefix, ksp->hdr.prefix, sizeof pc->hdr.prefix - 1);
  PetscCall(PCSetFromOptions(pc));
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetUp(KSP ksp)
{
  PC pc;
  PetscValidHeader(ksp, 1);
  if (!ksp->type_set) PetscCall(KSPSetType(ksp, KSPGMRES));
  PetscCall(KSPGetPC(ksp, &pc));
  if (ksp->setupcalled) return PCSetUp(pc); /* KSPSetUp always hands over to PCSetUp, which refreshes when the operator changed (itfunc.c:449) */
  PetscCheck(pc->mat, ksp->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "Matrix must be set first: call KSPSetOperators()");
  PetscCheck(ksp->vec_rhs && ksp->vec_sol, ksp->hdr.comm, PETSC_ERR_ARG_WRONGSTATE, "KSPSetUp() needs the vectors of KSPSolve() in this mirror");
  if (ksp->ops.setup) PetscCall((*ksp->ops.setup)(ksp));
  PetscCall(PCSetUp(pc)); /* itfunc.c:449 */
  ksp->setupcalled = 1;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x)
{
  PetscValidHeader(ksp, 1);
  PetscValidHeader(b, 2);
  PetscValidHeader(x, 3);
  PetscCheck(b != x, ksp->hdr.comm, PETSC_ERR_ARG_IDN, "b and x must be different vectors (in-place solve is not mirrored)");
  ksp->vec_rhs = b;
  ksp->vec_sol = x;
  PetscCall(KSPSetUp(ksp));
  if (ksp->guess_zero) PetscCall(VecSet(x, 0.0)); /* itfunc.c:905 */
  ksp->reason = KSP_CONVERGED_ITERATING;
  ksp->its    = 0;
  if (ksp->res_hist_reset) ksp->res_hist_len = 0;
  PetscCall((*ksp->ops.solve)(ksp)); /* itfunc.c:936 */
  PetscCheck(ksp->reason, ksp->hdr.comm, PETSC_ERR_PLIB, "Internal error, solver returned without setting converged reason");
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its)
{
  *its = ksp->its;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm)
{
  *rnorm = ksp->rnorm;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetConvergedReason(KSP ksp, KSPConvergedReason *reason)
{
  *reason = ksp->reason;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPSetResidualHistory(KSP ksp, PetscReal a[], PetscInt na, PetscBool reset)
{
  if (ksp->res_hist_alloc) free(ksp->res_hist);
  ksp->res_hist_alloc = 0;
  if (a) {
    ksp->res_hist     = a;
    ksp->res_hist_max = na;
  } else {
    ksp->res_hist_max = (na > 0) ? na : 10000; /* iterativ.c: default length */
    ksp->res_hist     = (PetscReal *)calloc((size_t)ksp->res_hist_max, sizeof(PetscReal));
    ksp->res_hist_alloc = 1;
  }
  ksp->res_hist_len   = 0;
  ksp->res_hist_reset = reset;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGetResidualHistory(KSP ksp, const PetscReal *a[], PetscInt *na)
{
  if (a) *a = ksp->res_hist;
  if (na) *na = ksp->res_hist_len;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPMonitorSet(KSP ksp, PetscErrorCode (*monitor)(KSP, PetscInt, PetscReal, void *), void *ctx, PetscErrorCode (*monitordestroy)(void **))
{
  (void)monitordestroy;
  ksp->monitor    = monitor;
  ksp->monitorctx = ctx;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPLogResidualHistory(KSP ksp, PetscReal norm)
{
  if (ksp->res_hist && ksp->res_hist_max > ksp->res_hist_len) ksp->res_hist[ksp->res_hist_len++] = norm; /* kspimpl.h KSPLogResidualHistory */
  return PETSC_SUCCESS;
}
PetscErrorCode KSPMonitor(KSP ksp, PetscInt it, PetscReal rnorm)
{
  if (ksp->monitor_stdout && PetscB200CommRank(ksp->hdr.comm) == 0) {
    printf("%3d KSP Residual norm %14.12e\n", it, (double)rnorm); /* iterativ.c:141 */
    fflush(stdout);
  }
  if (ksp->monitor) PetscCall((*ksp->monitor)(ksp, it, rnorm, ksp->monitorctx));
  return PETSC_SUCCESS;
}
PetscErrorCode KSPDestroy(KSP *ksp)
{
  if (!ksp || !*ksp) return PETSC_SUCCESS;
  if (--(*ksp)->hdr.refct > 0) {
    *ksp = NULL;
    return PETSC_SUCCESS;
  }
  PetscCall(KSPReset_Private(*ksp));
  if ((*ksp)->ops.destroy) PetscCall((*(*ksp)->ops.destroy)(*ksp));
  PetscCall(PCDestroy(&(*ksp)->pc));
  if ((*ksp)->res_hist_alloc) free((*ksp)->res_hist);
  free(*ksp);
  *ksp = NULL;
  return PETSC_SUCCESS;
}

/* KSPConvergedDefault (iterativ.c:1490-1581), preconditioned norm, left PC */
PetscErrorCode KSPConvergedDefault(KSP ksp, PetscInt n, PetscReal rnorm, KSPConvergedReason *reason)
{
  *reason = KSP_CONVERGED_ITERATING;
  if (!n) {
    if (!ksp->guess_zero) { /* iterativ.c:1512-1539: nonzero guess -> norm of the preconditioned right-hand side */
      PetscReal snorm = 0.0;
      Vec       z;
      PetscCall(VecDuplicate(ksp->vec_rhs, &z));
      PetscCall(PCApply(ksp->pc, ksp->vec_rhs, z));
      PetscCall(VecNorm(z, NORM_2, &snorm));
      PetscCall(VecDestroy(&z));
      if (!snorm) snorm = rnorm;
      ksp->rnorm0 = snorm;
    } else ksp->rnorm0 = rnorm;
    ksp->ttol = fmax(ksp->rtol * ksp->rnorm0, ksp->abstol);
  }
  if (isnan(rnorm) || isinf(rnorm)) {
    *reason = KSP_DIVERGED_NANORINF;
    return PETSC_SUCCESS;
  }
  if (rnorm <= ksp->ttol) *reason = rnorm < ksp->abstol ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL;
  else if (rnorm >= ksp->divtol * ksp->rnorm0) *reason = KSP_DIVERGED_DTOL;
  return PETSC_SUCCESS;
}

/* KSPInitialResidual (itres.c:35-73), left preconditioning */
PetscErrorCode KSPInitialResidual(KSP ksp, Vec vsoln, Vec vt1, Vec vt2, Vec vres, Vec vb)
{
  if (!ksp->guess_zero) {
    PetscCall(MatMult(ksp->pc->mat, vsoln, vt1));
    PetscCall(VecCopy(vb, vt2));
    PetscCall(VecAXPY(vt2, -1.0, vt1));
    PetscCall(PCApply(ksp->pc, vt2, vres));
  } else {
    PetscCall(VecCopy(vb, vt2));
    PetscCall(PCApply(ksp->pc, vb, vres));
  }
  return PETSC_SUCCESS;
}

/* ================================================================== GMRES (gmres.c, borthog2.c, gmresimpl.h) */
#define VEC_OFFSET     2
#define VEC_TEMP       gmres->vecs[0]
#define VEC_TEMP_MATOP gmres->vecs[1]
#define VEC_VV(i)      gmres->vecs[VEC_OFFSET + (i)]
#define HH(a, b)       (gmres->hh_origin + (b) * (gmres->max_k + 2) + (a))
#define HES(a, b)      (gmres->hes_origin + (b) * (gmres->max_k + 1) + (a))
#define CC(a)          (gmres->cc_origin + (a))
#define SS(a)          (gmres->ss_origin + (a))
#define GRS(a)         (gmres->rs_origin + (a))
typedef struct {
  PetscScalar *hh_origin, *hes_origin, *cc_origin, *ss_origin, *rs_origin, *orthogwork, *nrs;
  PetscInt     max_k, it, delta_allocate, vv_allocated, vecs_allocated, nwork_alloc;
  PetscReal    haptol, breakdowntol, rnorm0;
  KSPGMRESCGSRefinementType cgstype;
  Vec        *vecs, **user_work;
  PetscInt   *mwork_alloc;
  int         q_preallocate;
} KSP_GMRES;

static PetscErrorCode KSPSetUp_GMRES(KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  PetscInt   max_k = gmres->max_k, k;
  /* gmres.c:35-70 */
  gmres->hh_origin  = (PetscScalar *)calloc((size_t)(max_k + 2) * (max_k + 1), sizeof(PetscScalar));
  gmres->hes_origin = (PetscScalar *)calloc((size_t)(max_k + 1) * (max_k + 1), sizeof(PetscScalar));
  gmres->rs_origin  = (PetscScalar *)calloc((size_t)max_k + 2, sizeof(PetscScalar));
  gmres->cc_origin  = (PetscScalar *)calloc((size_t)max_k + 1, sizeof(PetscScalar));
  gmres->ss_origin  = (PetscScalar *)calloc((size_t)max_k + 1, sizeof(PetscScalar));
  gmres->nrs        = (PetscScalar *)calloc((size_t)max_k + 2, sizeof(PetscScalar));
  gmres->vecs_allocated = VEC_OFFSET + 2 + max_k;
  gmres->vecs           = (Vec *)calloc((size_t)gmres->vecs_allocated, sizeof(Vec));
  gmres->user_work      = (Vec **)calloc((size_t)(VEC_OFFSET + 2 + max_k), sizeof(Vec *));
  gmres->mwork_alloc    = (PetscInt *)calloc((size_t)(VEC_OFFSET + 2 + max_k), sizeof(PetscInt));
  if (gmres->q_preallocate) gmres->vv_allocated = VEC_OFFSET + 2 + (max_k < ksp->max_it ? max_k : ksp->max_it);
  else {
    PetscInt five = max_k < 5 ? max_k : 5;
    gmres->vv_allocated = VEC_OFFSET + 2 + (five < ksp->max_it ? five : ksp->max_it);
  }
  PetscCall(VecDuplicateVecs(ksp->vec_rhs, gmres->vv_allocated, &gmres->user_work[0]));
  gmres->mwork_alloc[0] = gmres->vv_allocated;
  gmres->nwork_alloc    = 1;
  for (k = 0; k < gmres->vv_allocated; k++) gmres->vecs[k] = gmres->user_work[0][k];
  return PETSC_SUCCESS;
}

static PetscErrorCode KSPGMRESGetNewVectors(KSP ksp, PetscInt it)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data; /* gmres.c:399-418 */
  PetscInt   nwork = gmres->nwork_alloc, k, nalloc;
  nalloc = ksp->max_it < gmres->delta_allocate ? ksp->max_it : gmres->delta_allocate;
  if (it + VEC_OFFSET + nalloc >= gmres->vecs_allocated) nalloc = gmres->vecs_allocated - it - VEC_OFFSET;
  if (!nalloc) return PETSC_SUCCESS;
  gmres->vv_allocated += nalloc;
  PetscCall(VecDuplicateVecs(ksp->vec_rhs, nalloc, &gmres->user_work[nwork]));
  gmres->mwork_alloc[nwork] = nalloc;
  for (k = 0; k < nalloc; k++) gmres->vecs[it + VEC_OFFSET + k] = gmres->user_work[nwork][k];
  gmres->nwork_alloc++;
  return PETSC_SUCCESS;
}

/* borthog2.c:33-114 */
static PetscErrorCode KSPGMRESClassicalGramSchmidtOrthogonalization(KSP ksp, PetscInt it)
{
  KSP_GMRES   *gmres = (KSP_GMRES *)ksp->data;
  PetscInt     j;
  PetscScalar *hh, *hes, *lhh;
  PetscReal    hnrm, wnrm;
  int          refine = (gmres->cgstype == KSP_GMRES_CGS_REFINE_ALWAYS);
  if (!gmres->orthogwork) gmres->orthogwork = (PetscScalar *)calloc((size_t)gmres->max_k + 2, sizeof(PetscScalar));
  lhh = gmres->orthogwork;
  hh  = HH(0, it);
  hes = HES(0, it);
  for (j = 0; j <= it; j++) hh[j] = hes[j] = 0.0;
  PetscCall(VecMDot(VEC_VV(it + 1), it + 1, &(VEC_VV(0)), lhh)); /* <v,vnew> */
  for (j = 0; j <= it; j++) {
    if (isnan(lhh[j]) || isinf(lhh[j])) { /* KSPCheckDot */
      ksp->reason = KSP_DIVERGED_NANORINF;
      return PETSC_SUCCESS;
    }
    lhh[j] = -lhh[j];
  }
  PetscCall(VecMAXPY(VEC_VV(it + 1), it + 1, lhh, &VEC_VV(0)));
  for (j = 0; j <= it; j++) {
    hh[j] -= lhh[j];
    hes[j] -= lhh[j];
  }
  if (gmres->cgstype == KSP_GMRES_CGS_REFINE_IFNEEDED) {
    hnrm = 0.0;
    for (j = 0; j <= it; j++) hnrm += lhh[j] * lhh[j];
    hnrm = sqrt(hnrm);
    PetscCall(VecNorm(VEC_VV(it + 1), NORM_2, &wnrm));
    if (wnrm < hnrm) refine = 1;
  }
  if (refine) {
    PetscCall(VecMDot(VEC_VV(it + 1), it + 1, &(VEC_VV(0)), lhh));
    for (j = 0; j <= it; j++) lhh[j] = -lhh[j];
    PetscCall(VecMAXPY(VEC_VV(it + 1), it + 1, lhh, &VEC_VV(0)));
    for (j = 0; j <= it; j++) {
      hh[j] -= lhh[j];
      hes[j] -= lhh[j];
    }
  }
  return PETSC_SUCCESS;
}

/* gmres.c:346-397 */
static PetscErrorCode KSPGMRESUpdateHessenberg(KSP ksp, PetscInt it, int hapend, PetscReal *res)
{
  KSP_GMRES   *gmres = (KSP_GMRES *)ksp->data;
  PetscScalar *hh = HH(0, it), *cc = CC(0), *ss = SS(0), tt;
  for (PetscInt j = 1; j <= it; j++) {
    tt  = *hh;
    *hh = *cc * tt + *ss * *(hh + 1);
    hh++;
    *hh = *cc++ * *hh - (*ss++ * tt);
  }
  if (!hapend) {
    tt = sqrt(*hh * *hh + *(hh + 1) * *(hh + 1));
    if (tt == 0.0) {
      ksp->reason = KSP_DIVERGED_NULL;
      return PETSC_SUCCESS;
    }
    *cc          = *hh / tt;
    *ss          = *(hh + 1) / tt;
    *GRS(it + 1) = -(*ss * *GRS(it));
    *GRS(it)     = *cc * *GRS(it);
    *hh          = *cc * *hh + *ss * *(hh + 1);
    *res         = fabs(*GRS(it + 1));
  } else *res = 0.0;
  return PETSC_SUCCESS;
}

/* gmres.c:298-341 */
static PetscErrorCode KSPGMRESBuildSoln(PetscScalar *nrs, Vec vs, Vec vdest, KSP ksp, PetscInt it)
{
  KSP_GMRES  *gmres = (KSP_GMRES *)ksp->data;
  PetscScalar tt;
  PetscInt    ii, k, j;
  if (it < 0) {
    PetscCall(VecCopy(vs, vdest));
    return PETSC_SUCCESS;
  }
  if (*HH(it, it) != 0.0) nrs[it] = *GRS(it) / *HH(it, it);
  else {
    ksp->reason = KSP_DIVERGED_BREAKDOWN;
    return PETSC_SUCCESS;
  }
  for (ii = 1; ii <= it; ii++) {
    k  = it - ii;
    tt = *GRS(k);
    for (j = k + 1; j <= it; j++) tt = tt - *HH(k, j) * nrs[j];
    if (*HH(k, k) == 0.0) {
      ksp->reason = KSP_DIVERGED_BREAKDOWN;
      return PETSC_SUCCESS;
    }
    nrs[k] = tt / *HH(k, k);
  }
  PetscCall(VecMAXPBY(VEC_TEMP, it + 1, nrs, 0, &VEC_VV(0)));
  /* KSPUnwindPreconditioner: nothing to do for left preconditioning */
  if (vdest != vs) PetscCall(VecCopy(vs, vdest));
  PetscCall(VecAXPY(vdest, 1.0, VEC_TEMP));
  return PETSC_SUCCESS;
}

/* gmres.c:88-193 */
static PetscErrorCode KSPGMRESCycle(PetscInt *itcount, KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  PetscReal  res, hapbnd, tt;
  PetscInt   it = 0, max_k = gmres->max_k;
  int        hapend = 0;
  if (itcount) *itcount = 0;
  PetscCall(VecNormalize(VEC_VV(0), &res));
  if (isnan(res) || isinf(res)) { /* KSPCheckNorm */
    ksp->reason = KSP_DIVERGED_NANORINF;
    ksp->rnorm  = res;
    return PETSC_SUCCESS;
  }
  if ((ksp->rnorm > 0.0) && (fabs(res - ksp->rnorm) > gmres->breakdowntol * gmres->rnorm0)) {
    ksp->reason = KSP_DIVERGED_BREAKDOWN;
    return PETSC_SUCCESS;
  }
  *GRS(0) = gmres->rnorm0 = res;
  ksp->rnorm              = res;
  gmres->it               = (it - 1);
  PetscCall(KSPLogResidualHistory(ksp, res));
  PetscCall(KSPMonitor(ksp, ksp->its, res));
  if (!res) {
    ksp->reason = KSP_CONVERGED_ATOL;
    return PETSC_SUCCESS;
  }
  PetscCall(KSPConvergedDefault(ksp, ksp->its, res, &ksp->reason));
  while (!ksp->reason && it < max_k && ksp->its < ksp->max_it) {
    if (it) {
      PetscCall(KSPLogResidualHistory(ksp, res));
      PetscCall(KSPMonitor(ksp, ksp->its, res));
    }
    gmres->it = (it - 1);
    if (gmres->vv_allocated <= it + VEC_OFFSET + 1) PetscCall(KSPGMRESGetNewVectors(ksp, it + 1));
    PetscCall(PCApplyBAorAB(ksp->pc, 0, VEC_VV(it), VEC_VV(1 + it), VEC_TEMP_MATOP)); /* KSP_PCApplyBAorAB, kspimpl.h:469 */
    PetscCall(KSPGMRESClassicalGramSchmidtOrthogonalization(ksp, it));
    if (ksp->reason) break;
    PetscCall(VecNormalize(VEC_VV(it + 1), &tt));
    if (isnan(tt) || isinf(tt)) {
      ksp->reason = KSP_DIVERGED_NANORINF;
      break;
    }
    *HH(it + 1, it)  = tt;
    *HES(it + 1, it) = tt;
    hapbnd = fabs(tt / *GRS(it));
    if (hapbnd > gmres->haptol) hapbnd = gmres->haptol;
    if (tt < hapbnd) hapend = 1;
    PetscCall(KSPGMRESUpdateHessenberg(ksp, it, hapend, &res));
    it++;
    gmres->it = (it - 1);
    ksp->its++;
    ksp->rnorm = res;
    if (ksp->reason) break;
    PetscCall(KSPConvergedDefault(ksp, ksp->its, res, &ksp->reason));
    if (hapend && !ksp->reason) {
      ksp->reason = KSP_DIVERGED_BREAKDOWN;
      break;
    }
  }
  if (itcount) *itcount = it;
  PetscCall(KSPGMRESBuildSoln(GRS(0), ksp->vec_sol, ksp->vec_sol, ksp, it - 1));
  if (ksp->reason == KSP_CONVERGED_ITERATING && ksp->its >= ksp->max_it) ksp->reason = KSP_DIVERGED_ITS;
  if (it && ksp->reason) {
    PetscCall(KSPLogResidualHistory(ksp, res));
    PetscCall(KSPMonitor(ksp, ksp->its, res));
  }
  return PETSC_SUCCESS;
}

/* gmres.c:196-238 */
static PetscErrorCode KSPSolve_GMRES(KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  PetscInt   its, itcount = 0;
  int        guess_zero = ksp->guess_zero;
  ksp->its   = 0;
  ksp->rnorm = -1.0;
  while (!ksp->reason) {
    PetscCall(KSPInitialResidual(ksp, ksp->vec_sol, VEC_TEMP, VEC_TEMP_MATOP, VEC_VV(0), ksp->vec_rhs));
    PetscCall(KSPGMRESCycle(&its, ksp));
    itcount += its;
    if (itcount >= ksp->max_it) {
      if (!ksp->reason) ksp->reason = KSP_DIVERGED_ITS;
      break;
    }
    ksp->guess_zero = 0;
  }
  ksp->guess_zero = guess_zero;
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPReset_GMRES(KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  free(gmres->hh_origin); free(gmres->hes_origin); free(gmres->rs_origin); free(gmres->cc_origin); free(gmres->ss_origin); free(gmres->nrs); free(gmres->orthogwork);
  gmres->hh_origin = gmres->hes_origin = gmres->rs_origin = gmres->cc_origin = gmres->ss_origin = gmres->nrs = gmres->orthogwork = NULL;
  free(gmres->vecs);
  gmres->vecs = NULL;
  for (PetscInt i = 0; i < gmres->nwork_alloc; i++) PetscCall(VecDestroyVecs(gmres->mwork_alloc[i], &gmres->user_work[i]));
  gmres->nwork_alloc = 0;
  free(gmres->user_work); free(gmres->mwork_alloc);
  gmres->user_work   = NULL;
  gmres->mwork_alloc = NULL;
  gmres->vv_allocated = gmres->vecs_allocated = 0;
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPDestroy_GMRES(KSP ksp)
{
  free(ksp->data);
  ksp->data = NULL;
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPSetFromOptions_GMRES(KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  char       s[64];
  PetscBool  set, pre = (PetscBool)gmres->q_preallocate;
  PetscInt   restart = gmres->max_k;
  PetscCall(PetscOptionsGetInt(NULL, ksp->hdr.prefix, "-ksp_gmres_restart", &restart, &set));
  if (set) PetscCall(KSPGMRESSetRestart(ksp, restart));
  PetscCall(PetscOptionsGetReal(NULL, ksp->hdr.prefix, "-ksp_gmres_haptol", &gmres->haptol, NULL));
  PetscCall(PetscOptionsGetReal(NULL, ksp->hdr.prefix, "-ksp_gmres_breakdown_tolerance", &gmres->breakdowntol, NULL));
  PetscCall(PetscOptionsGetBool(NULL, ksp->hdr.prefix, "-ksp_gmres_preallocate", &pre, NULL));
  gmres->q_preallocate = pre;
  PetscCall(PetscOptionsGetString(NULL, ksp->hdr.prefix, "-ksp_gmres_cgs_refinement_type", s, sizeof s, &set));
  if (set) {
    if (!strcmp(s, "refine_never")) gmres->cgstype = KSP_GMRES_CGS_REFINE_NEVER;
    else if (!strcmp(s, "refine_ifneeded")) gmres->cgstype = KSP_GMRES_CGS_REFINE_IFNEEDED;
    else if (!strcmp(s, "refine_always")) gmres->cgstype = KSP_GMRES_CGS_REFINE_ALWAYS;
    else SETERRQ(ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Unknown -ksp_gmres_cgs_refinement_type %s", s);
  }
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGMRESSetRestart(KSP ksp, PetscInt max_k)
{
  PetscCheck(!strcmp(ksp->hdr.type_name, KSPGMRES) || ksp->data, ksp->hdr.comm, PETSC_ERR_ARG_WRONG, "not a GMRES solver");
  KSP_GMRES *gmres = (KSP_GMRES *)ksp->data;
  PetscCheck(max_k >= 1, ksp->hdr.comm, PETSC_ERR_ARG_OUTOFRANGE, "Restart must be positive"); /* gmres.c:590 */
  if (!ksp->setupcalled) gmres->max_k = max_k;
  else if (gmres->max_k != max_k) {
    gmres->max_k = max_k;
    PetscCall(KSPReset_Private(ksp));
  }
  return PETSC_SUCCESS;
}
PetscErrorCode KSPGMRESSetCGSRefinementType(KSP ksp, KSPGMRESCGSRefinementType type)
{
  ((KSP_GMRES *)ksp->data)->cgstype = type;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPCreate_GMRES(KSP ksp)
{
  KSP_GMRES *gmres = (KSP_GMRES *)calloc(1, sizeof(*gmres));
  PetscCheck(gmres, ksp->hdr.comm, PETSC_ERR_MEM, "out of memory");
  /* gmres.c:905-915 */
  gmres->haptol         = 1.0e-30;
  gmres->breakdowntol   = 0.1;
  gmres->delta_allocate = 10;
  gmres->max_k          = 30;
  gmres->cgstype        = KSP_GMRES_CGS_REFINE_NEVER;
  ksp->data               = gmres;
  ksp->ops.setup          = KSPSetUp_GMRES;
  ksp->ops.solve          = KSPSolve_GMRES;
  ksp->ops.reset          = KSPReset_GMRES;
  ksp->ops.destroy        = KSPDestroy_GMRES;
  ksp->ops.setfromoptions = KSPSetFromOptions_GMRES;
  return PETSC_SUCCESS;
}

/* ================================================================== CG (cg.c:119-350), preconditioned norm */
static PetscErrorCode KSPSetUp_CG(KSP ksp)
{
  ksp->nwork = 3; /* cg.c KSPSetUp_CG: KSPSetWorkVecs(ksp, 3) */
  PetscCall(VecDuplicateVecs(ksp->vec_rhs, 3, &ksp->work));
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPSolve_CG(KSP ksp)
{
  PetscInt    i;
  PetscScalar dpi = 0.0, a = 1.0, beta, betaold = 1.0, b = 0, dpiold;
  PetscReal   dp  = 0.0;
  Vec         X = ksp->vec_sol, B = ksp->vec_rhs, R = ksp->work[0], Z = ksp->work[1], P = ksp->work[2], W = Z;
  Mat         Amat = ksp->pc->mat;
  ksp->its = 0;
  if (!ksp->guess_zero) {
    PetscCall(MatMult(Amat, X, R));
    PetscCall(VecAYPX(R, -1.0, B));
  } else PetscCall(VecCopy(B, R));
  PetscCall(PCApply(ksp->pc, R, Z));
  PetscCall(VecNorm(Z, NORM_2, &dp));
  if (isnan(dp) || isinf(dp)) {
    ksp->reason = KSP_DIVERGED_NANORINF;
    return PETSC_SUCCESS;
  }
  PetscCall(KSPLogResidualHistory(ksp, dp));
  PetscCall(KSPMonitor(ksp, ksp->its, dp));
  ksp->rnorm = dp;
  PetscCall(KSPConvergedDefault(ksp, ksp->its, dp, &ksp->reason));
  if (ksp->reason) return PETSC_SUCCESS;
  PetscCall(VecDot(Z, R, &beta));
  i = 0;
  do {
    ksp->its = i + 1;
    if (beta == 0.0) {
      ksp->reason = KSP_CONVERGED_ATOL;
      break;
    } else if ((i > 0) && (beta * betaold < 0.0)) {
      ksp->reason = KSP_DIVERGED_INDEFINITE_PC;
      break;
    }
    if (!i) {
      PetscCall(VecCopy(Z, P));
      b = 0.0;
    } else {
      b = beta / betaold;
      PetscCall(VecAYPX(P, b, Z));
    }
    dpiold = dpi;
    PetscCall(MatMult(Amat, P, W));
    PetscCall(VecDot(P, W, &dpi));
    betaold = beta;
    if ((dpi == 0.0) || ((i > 0) && ((dpi > 0) - (dpi < 0)) * ((dpiold > 0) - (dpiold < 0)) < 0)) {
      ksp->reason = KSP_DIVERGED_INDEFINITE_MAT;
      break;
    }
    a = beta / dpi;
    PetscCall(VecAXPY(X, a, P));
    PetscCall(VecAXPY(R, -a, W));
    PetscCall(PCApply(ksp->pc, R, Z));
    PetscCall(VecNorm(Z, NORM_2, &dp));
    if (isnan(dp) || isinf(dp)) {
      ksp->reason = KSP_DIVERGED_NANORINF;
      break;
    }
    ksp->rnorm = dp;
    PetscCall(KSPLogResidualHistory(ksp, dp));
    PetscCall(KSPMonitor(ksp, i + 1, dp));
    PetscCall(KSPConvergedDefault(ksp, i + 1, dp, &ksp->reason));
    if (ksp->reason) break;
    PetscCall(VecDot(Z, R, &beta));
    i++;
  } while (i < ksp->max_it);
  if (i >= ksp->max_it) ksp->reason = KSP_DIVERGED_ITS;
  (void)b;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPCreate_CG(KSP ksp)
{
  ksp->ops.setup = KSPSetUp_CG;
  ksp->ops.solve = KSPSolve_CG;
  return PETSC_SUCCESS;
}

/* ================================================================== PIPECG (cg/pipecg/pipecg.c:19-160), preconditioned norm
   The single-reduction CG of SURVEY 8f.3: |u|^2, r.u and w.u of an iteration come out of ONE kernel (VecMDot of u against
   {u, r, w}: one pass over the three vectors, one result transfer, one all-reduce of three numbers on several GPUs) instead of
   the three separate reductions -- and three host synchronisations -- of KSPCG.  The PCApply and MatMult that do not depend on
   those numbers are enqueued first, so the device works through them while the host waits for the reduction (the overlap
   the reference gets from VecNormBegin/VecDotBegin + PetscCommSplitReductionBegin). */
static PetscErrorCode KSPSetUp_PIPECG(KSP ksp)
{
  ksp->nwork = 9; /* pipecg.c KSPSetUp_PIPECG: KSPSetWorkVecs(ksp, 9) */
  PetscCall(VecDuplicateVecs(ksp->vec_rhs, 9, &ksp->work));
  return PETSC_SUCCESS;
}
static PetscErrorCode KSPSolve_PIPECG(KSP ksp)
{
  PetscInt    i;
  PetscScalar alpha = 0.0, beta = 0.0, gamma = 0.0, gammaold = 0.0, delta = 0.0, red[3];
  PetscReal   dp = 0.0;
  Vec         X = ksp->vec_sol, B = ksp->vec_rhs, R = ksp->work[0], Z = ksp->work[1], P = ksp->work[2], N = ksp->work[3], W = ksp->work[4];
  Vec         Q = ksp->work[5], U = ksp->work[6], M = ksp->work[7], S = ksp->work[8], ys[3];
  Mat         Amat = ksp->pc->mat;
  PetscBool   fuse = PETSC_TRUE; /* -ksp_pipecg_b200_fuse_update <bool>: the eight vector recurrences as one kernel */
  PetscCall(PetscOptionsGetBool(NULL, ksp->hdr.prefix, "-ksp_pipecg_b200_fuse_update", &fuse, NULL));
  ksp->its = 0;
  if (!ksp->guess_zero) {
    PetscCall(MatMult(Amat, X, R));
    PetscCall(VecAYPX(R, -1.0, B));
  } else PetscCall(VecCopy(B, R));
  PetscCall(PCApply(ksp->pc, R, U)); /* u <- Br */
  PetscCall(MatMult(Amat, U, W));    /* w <- Au */
  PetscCall(VecNorm(U, NORM_2, &dp));
  if (isnan(dp) || isinf(dp)) {
    ksp->reason = KSP_DIVERGED_NANORINF;
    return PETSC_SUCCESS;
  }
  PetscCall(KSPLogResidualHistory(ksp, dp));
  PetscCall(KSPMonitor(ksp, 0, dp));
  ksp->rnorm = dp;
  PetscCall(KSPConvergedDefault(ksp, 0, dp, &ksp->reason));
  if (ksp->reason) return PETSC_SUCCESS;
  ys[0] = U; ys[1] = R; ys[2] = W;
  i = 0;
  do {
    PetscCall(PCApply(ksp->pc, W, M));   /* m <- Bw   (independent of the reduction below) */
    PetscCall(MatMult(Amat, M, N));      /* n <- Am */
    PetscCall(VecMDot(U, 3, ys, red));   /* |u|^2, r.u, w.u: one kernel, one synchronisation */
    gamma = red[1];
    delta = red[2];
    if (i > 0) {
      dp = sqrt(fabs(red[0]));
      if (isnan(dp) || isinf(dp)) {
        ksp->reason = KSP_DIVERGED_NANORINF;
        return PETSC_SUCCESS;
      }
      ksp->rnorm = dp;
      PetscCall(KSPLogResidualHistory(ksp, dp));
      PetscCall(KSPMonitor(ksp, i, dp));
      PetscCall(KSPConvergedDefault(ksp, i, dp, &ksp->reason));
      if (ksp->reason) return PETSC_SUCCESS;
    }
    if (i == 0) alpha = gamma / delta;
    else {
      beta  = gamma / gammaold;
      alpha = gamma / (delta - beta / alpha * gamma);
    }
    if (fuse) {
      /* the four VecCopy/VecAYPX and four VecAXPY below as ONE kernel (b200VecPipeCGUpdate): same arithmetic per entry */
      const double *dn, *dm;
      double       *du, *dw, *dz, *dq, *dp, *ds, *dx, *dr;
      PetscInt      nloc;
      PetscCall(VecGetLocalSize(X, &nloc));
      PetscCall(VecB200GetArrayRead(N, &dn));
      PetscCall(VecB200GetArrayRead(M, &dm));
      PetscCall(VecB200GetArray(U, &du));
      PetscCall(VecB200GetArray(W, &dw));
      if (i == 0) {
        PetscCall(VecB200GetArrayWrite(Z, &dz));
        PetscCall(VecB200GetArrayWrite(Q, &dq));
        PetscCall(VecB200GetArrayWrite(P, &dp));
        PetscCall(VecB200GetArrayWrite(S, &ds));
      } else {
        PetscCall(VecB200GetArray(Z, &dz));
        PetscCall(VecB200GetArray(Q, &dq));
        PetscCall(VecB200GetArray(P, &dp));
        PetscCall(VecB200GetArray(S, &ds));
      }
      PetscCall(VecB200GetArray(X, &dx));
      PetscCall(VecB200GetArray(R, &dr));
      PetscCallB200(b200VecPipeCGUpdate(PetscB200.h, nloc, alpha, beta, i == 0, dn, dm, du, dw, dz, dq, dp, ds, dx, dr));
    } else {
      if (i == 0) {
        PetscCall(VecCopy(N, Z)); /* z <- n */
        PetscCall(VecCopy(M, Q)); /* q <- m */
        PetscCall(VecCopy(U, P)); /* p <- u */
        PetscCall(VecCopy(W, S)); /* s <- w */
      } else {
        PetscCall(VecAYPX(Z, beta, N)); /* z <- n + beta z */
        PetscCall(VecAYPX(Q, beta, M));
        PetscCall(VecAYPX(P, beta, U));
        PetscCall(VecAYPX(S, beta, W));
      }
      PetscCall(VecAXPY(X, alpha, P));  /* x <- x + alpha p */
      PetscCall(VecAXPY(U, -alpha, Q)); /* u <- u - alpha q */
      PetscCall(VecAXPY(W, -alpha, Z)); /* w <- w - alpha z */
      PetscCall(VecAXPY(R, -alpha, S)); /* r <- r - alpha s */
    }
    gammaold = gamma;
    i++;
    ksp->its = i;
  } while (i <= ksp->max_it);
  if (!ksp->reason) ksp->reason = KSP_DIVERGED_ITS;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPCreate_PIPECG(KSP ksp)
{
  ksp->ops.setup = KSPSetUp_PIPECG;
  ksp->ops.solve = KSPSolve_PIPECG;
  return PETSC_SUCCESS;
}

/* ================================================================== PREONLY (impls/preonly/preonly.c) */
static PetscErrorCode KSPSolve_PREONLY(KSP ksp)
{
  PetscCheck(ksp->guess_zero, ksp->hdr.comm, PETSC_ERR_SUP, "Running KSP of preonly doesn't make sense with nonzero initial guess");
  ksp->its = 0;
  PetscCall(PCApply(ksp->pc, ksp->vec_rhs, ksp->vec_sol));
  ksp->its    = 1;
  ksp->reason = KSP_CONVERGED_ITS;
  return PETSC_SUCCESS;
}
PetscErrorCode KSPCreate_PREONLY(KSP ksp)
{
  ksp->ops.solve = KSPSolve_PREONLY;
  return PETSC_SUCCESS;
}
