/* iccsym.c -- host-side SYMBOLIC phase of ICC(0) for the planned device factorisation.
   Index work only (no floating point): the factor layout, the order in which the reference merges finished rows into a row,
   the dependency levels of a row-parallel factorisation and the column view the scatter-free forward sweep needs.  The
   numeric kernels that consume these arrays are not built yet; the arrays are checked on the CPU against the test suite's
   CPU restatement of the reference (tests/test_icc_symbolic_cpu.py: index-exact), so the device path of the next round starts
   from a verified schedule -- the same split b200Ilu0Symbolic (host) / ilu_numeric_kernel (device) uses for ILU(0).
   All output arrays are allocated by the caller. */
#include "hostimpl.h"

/* MatICCFactorSymbolic_SeqAIJ, levels = 0 and identity permutation (aijfact.c:2078-2094): row i of the factor holds the
   strictly-upper entries of A's row i in column order and the diagonal LAST.
   ui[n+1], udiag[n], uj[at least (nnz + n) / 2 + n entries; ai[n] always suffices] */
PetscErrorCode PetscB200ICC0Symbolic(PetscInt n, const PetscInt *ai, const PetscInt *aj, PetscInt *ui, PetscInt *uj, PetscInt *udiag)
{
  PetscCheck(n >= 0 && ai && (aj || !n) && ui && uj && udiag, 0, PETSC_ERR_ARG_NULL, "null argument");
  ui[0] = 0;
  for (PetscInt i = 0; i < n; i++) {
    PetscInt d = -1, q = ui[i];
    for (PetscInt k = ai[i]; k < ai[i + 1]; k++)
      if (aj[k] == i) {
        d = k;
        break;
      }
    PetscCheck(d >= 0, 0, PETSC_ERR_ARG_WRONGSTATE, "Matrix is missing diagonal entries"); /* aijfact.c:2071 */
    for (PetscInt k = d + 1; k < ai[i + 1]; k++) uj[q++] = aj[k];
    uj[q]     = i;
    udiag[i]  = q;
    ui[i + 1] = q + 1;
  }
  return PETSC_SUCCESS;
}

/* The merge order of MatCholeskyFactorNumeric_SeqAIJ (aijfact.c:1750-1800).  The reference keeps, per column, a linked list
   (c2r) of the finished rows whose next unused entry lies in that column; rows are pushed at the head when they are factored
   and again each time they have been merged, so the list order (LIFO) -- and with it the floating-point association of every
   entry of row k -- is a function of the pattern alone.  This walks the same lists on indices:
     mptr[n+1]                 contributors of row k are mrow[mptr[k] .. mptr[k+1])
     mrow[], mpos[]            contributing row i and the position of U(i,k) inside the factor arrays   (ui[n] - n entries)
     level[n], *nlevels        level[k] = 1 + max level of its contributors (0 without any): rows of one level are independent */
PetscErrorCode PetscB200ICC0MergeSchedule(PetscInt n, const PetscInt *ui, const PetscInt *uj, PetscInt *mptr, PetscInt *mrow, PetscInt *mpos, PetscInt *level, PetscInt *nlevels)
{
  PetscCheck(n >= 0 && ui && (uj || !n) && mptr && (mrow || !n) && (mpos || !n) && (level || !n), 0, PETSC_ERR_ARG_NULL, "null argument");
  PetscInt *c2r = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)n + 1)), *il = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)n + 1));
  PetscInt  q = 0, maxlev = -1;
  PetscCheck(c2r && il, 0, PETSC_ERR_MEM, "out of memory");
  for (PetscInt i = 0; i <= n; i++) c2r[i] = n;
  if (n) il[0] = 0;
  mptr[0] = 0;
  for (PetscInt k = 0; k < n; k++) {
    PetscInt i = c2r[k], lev = 0;
    while (i < k) {
      const PetscInt nexti = c2r[i], ili = il[i], jmin = ili + 1, jmax = ui[i + 1];
      mrow[q] = i;
      mpos[q] = ili;
      q++;
      if (level[i] + 1 > lev) lev = level[i] + 1;
      if (jmin < jmax) { /* advance row i to its next entry and push it on that column's list */
        const PetscInt j = uj[jmin];
        il[i]  = jmin;
        c2r[i] = c2r[j];
        c2r[j] = i;
      }
      i = nexti;
    }
    if (ui[k] < ui[k + 1] - 1) { /* row k has off-diagonal entries: it joins the list of its first one */
      const PetscInt c = uj[ui[k]];
      il[k]  = ui[k];
      c2r[k] = c2r[c];
      c2r[c] = k;
    }
    level[k]    = lev;
    mptr[k + 1] = q;
    if (lev > maxlev) maxlev = lev;
  }
  if (nlevels) *nlevels = maxlev + 1;
  free(c2r);
  free(il);
  return PETSC_SUCCESS;
}

/* Column view of the strictly-upper pattern: for column c the entries (i, c) in ascending i -- the order in which the
   forward sweep of MatSolve_SeqSBAIJ_1_NaturalOrdering (sbaijfact2.c:2045-2052) adds v(i,c) * x_i into x[c], which a
   level-scheduled device sweep performs as a gather.  tptr[n+1], trow[ui[n]-n], tpos[ui[n]-n] (position in the factor arrays) */
PetscErrorCode PetscB200ICC0ColumnView(PetscInt n, const PetscInt *ui, const PetscInt *uj, PetscInt *tptr, PetscInt *trow, PetscInt *tpos)
{
  PetscCheck(n >= 0 && ui && (uj || !n) && tptr && (trow || !n) && (tpos || !n), 0, PETSC_ERR_ARG_NULL, "null argument");
  for (PetscInt c = 0; c <= n; c++) tptr[c] = 0;
  for (PetscInt i = 0; i < n; i++)
    for (PetscInt t = ui[i]; t < ui[i + 1] - 1; t++) tptr[uj[t] + 1]++;
  for (PetscInt c = 0; c < n; c++) tptr[c + 1] += tptr[c];
  PetscInt *fill = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)n + 1));
  PetscCheck(fill, 0, PETSC_ERR_MEM, "out of memory");
  for (PetscInt c = 0; c < n; c++) fill[c] = tptr[c];
  for (PetscInt i = 0; i < n; i++) /* ascending i: every column's list comes out in ascending row order */
    for (PetscInt t = ui[i]; t < ui[i + 1] - 1; t++) {
      const PetscInt c = uj[t];
      trow[fill[c]] = i;
      tpos[fill[c]] = t;
      fill[c]++;
    }
  free(fill);
  return PETSC_SUCCESS;
}
