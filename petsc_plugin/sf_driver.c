/*
 * sf_driver.c -- a PETSc program (public API only + the C ABI for raw device buffers) that checks VecScatter / PetscSF on the
 * plugin's device vectors against the same operation on the reference's host vectors (VECSEQ, PETSCSFBASIC host loops) in the
 * same process, bit for bit: general / block / stride index sets, forward and reverse, INSERT / ADD / MAX / MIN, repeated
 * sources and repeated destinations (order-dependent results), in-place and mixed host/device scatters (staged), a raw PetscSF
 * on PetscInt data.  With -vec_type standard every check degenerates to host-vs-host through the sub-classed "basic" type: that is
 * the CPU test that the plugin leaves host scatters alone.  Prints "ok <name>" per check, "all ok" at the end, non-zero exit on
 * the first mismatch.  Test infrastructure (tests/test_petsc_plugin_{cpu,gpu}.py); built by oracle/build_ref_demo.sh.
 */
#include <petscvec.h>
#include <petscsf.h>
#include "petscb200.h"

#define CHECK(cond, name) \
  do { \
    if (!(cond)) { \
      PetscCall(PetscPrintf(PETSC_COMM_SELF, "FAILED %s\n", name)); \
      PetscCall(PetscFinalize()); \
      return 1; \
    } \
    PetscCall(PetscPrintf(PETSC_COMM_SELF, "ok %s\n", name)); \
  } while (0)

static unsigned rng_state = 2463534242u;
static unsigned rnd(void)
{
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}
static double rndval(void) { return ((double)rnd() / 8388608.0) - 1.0; }

/* device-typed vector (whatever -vec_type says) and host reference vector with the same random contents */
static PetscErrorCode MakePair(PetscInt n, PetscInt bs, Vec *v, Vec *r)
{
  PetscScalar *a;
  PetscFunctionBegin;
  PetscCall(VecCreate(PETSC_COMM_SELF, v));
  PetscCall(VecSetSizes(*v, n, n));
  PetscCall(VecSetBlockSize(*v, bs));
  PetscCall(VecSetFromOptions(*v));
  PetscCall(VecCreateSeq(PETSC_COMM_SELF, n, r));
  PetscCall(VecSetBlockSize(*r, bs));
  PetscCall(VecGetArrayWrite(*r, &a));
  for (PetscInt i = 0; i < n; i++) a[i] = rndval();
  PetscCall(VecRestoreArrayWrite(*r, &a));
  PetscCall(VecCopy(*r, *v));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode Same(Vec v, Vec r, PetscBool *same)
{
  const PetscScalar *a, *b;
  PetscInt           n;
  PetscFunctionBegin;
  PetscCall(VecGetLocalSize(r, &n));
  PetscCall(VecGetArrayRead(v, &a));
  PetscCall(VecGetArrayRead(r, &b));
  *same = (PetscBool)(memcmp(a, b, sizeof(PetscScalar) * (size_t)n) == 0);
  PetscCall(VecRestoreArrayRead(v, &a));
  PetscCall(VecRestoreArrayRead(r, &b));
  PetscFunctionReturn(PETSC_SUCCESS);
}
/* the same scatter on (x,y) and on the host pair (xr,yr) */
static PetscErrorCode Both(VecScatter sc, VecScatter scr, Vec x, Vec y, Vec xr, Vec yr, InsertMode im, ScatterMode sm, PetscBool *same)
{
  PetscBool s1, s2;
  PetscFunctionBegin;
  PetscCall(VecScatterBegin(sc, x, y, im, sm));
  PetscCall(VecScatterEnd(sc, x, y, im, sm));
  PetscCall(VecScatterBegin(scr, xr, yr, im, sm));
  PetscCall(VecScatterEnd(scr, xr, yr, im, sm));
  PetscCall(Same(y, yr, &s1));
  PetscCall(Same(x, xr, &s2)); /* the source is left alone */
  *same = (PetscBool)(s1 && s2);
  PetscFunctionReturn(PETSC_SUCCESS);
}

int main(int argc, char **argv)
{
  const PetscInt nx = 5000, ny = 3600, nk = 3000;
  PetscInt      *from, *to, *perm, native = 0, staged = 0, nat0, stg0;
  IS             isf, ist;
  Vec            x, y, xr, yr;
  VecScatter     sc, scr;
  PetscBool      same, isb200;
  PetscErrorCode (*counts)(PetscSF, PetscInt *, PetscInt *) = NULL;

  PetscCall(PetscInitialize(&argc, &argv, NULL, NULL));
  PetscCall(MakePair(nx, 1, &x, &xr));
  PetscCall(MakePair(ny, 1, &y, &yr));
  PetscCall(PetscObjectTypeCompare((PetscObject)x, "seqb200", &isb200));
  {
    VecType vt;
    PetscCall(VecGetType(x, &vt));
    PetscCall(PetscPrintf(PETSC_COMM_SELF, "vec type %s\n", vt));
  }

  /* general index sets: sources repeat (about 1 in 4), destinations distinct */
  PetscCall(PetscMalloc3(nk, &from, nk, &to, ny, &perm));
  for (PetscInt i = 0; i < ny; i++) perm[i] = i;
  for (PetscInt i = ny - 1; i > 0; i--) {
    PetscInt j = (PetscInt)(rnd() % (unsigned)(i + 1)), t = perm[i];
    perm[i] = perm[j];
    perm[j] = t;
  }
  for (PetscInt k = 0; k < nk; k++) {
    from[k] = (PetscInt)(rnd() % (unsigned)nx);
    if (k % 4 == 3) from[k] = from[k - 1];
    to[k] = perm[k];
  }
  PetscCall(ISCreateGeneral(PETSC_COMM_SELF, nk, from, PETSC_COPY_VALUES, &isf));
  PetscCall(ISCreateGeneral(PETSC_COMM_SELF, nk, to, PETSC_COPY_VALUES, &ist));
  PetscCall(VecScatterCreate(x, isf, y, ist, &sc));
  PetscCall(VecScatterCreate(xr, isf, yr, ist, &scr));
  PetscCall(PetscObjectQueryFunction((PetscObject)sc, "PetscSFB200GetCounts_C", &counts));
  CHECK(counts != NULL, "vecscatter_sf_is_the_b200_subclass");
  PetscCall(Both(sc, scr, x, y, xr, yr, INSERT_VALUES, SCATTER_FORWARD, &same));
  CHECK(same, "general_forward_insert");
  PetscCall(Both(sc, scr, x, y, xr, yr, ADD_VALUES, SCATTER_FORWARD, &same));
  CHECK(same, "general_forward_add");
  PetscCall(Both(sc, scr, x, y, xr, yr, MAX_VALUES, SCATTER_FORWARD, &same));
  CHECK(same, "general_forward_max");
  /* reverse: several leaves per root -> the order of application matters for ADD (rounding) and INSERT (last one wins) */
  PetscCall(Both(sc, scr, y, x, yr, xr, ADD_VALUES, SCATTER_REVERSE, &same));
  CHECK(same, "general_reverse_add_repeated_roots");
  PetscCall(Both(sc, scr, y, x, yr, xr, INSERT_VALUES, SCATTER_REVERSE, &same));
  CHECK(same, "general_reverse_insert_repeated_roots");
  PetscCall(Both(sc, scr, y, x, yr, xr, MIN_VALUES, SCATTER_REVERSE, &same));
  CHECK(same, "general_reverse_min_repeated_roots");
  PetscCall((*counts)(sc, &native, &staged));
  PetscCall(PetscPrintf(PETSC_COMM_SELF, "general scatter: %" PetscInt_FMT " operations on the device, %" PetscInt_FMT " staged through the host\n", native, staged));
  CHECK(isb200 ? (native == 6 && staged == 0) : (native == 0 && staged == 0), "general_scatter_ran_where_expected");

  /* mixed: device source, host destination and the other way round (staged) */
  {
    Vec y2, y2r;
    PetscCall(VecDuplicate(yr, &y2r));
    PetscCall(VecCopy(yr, y2r));
    PetscCall(VecDuplicate(yr, &y2));
    PetscCall(VecCopy(yr, y2));
    PetscCall(Both(sc, scr, x, y2, xr, y2r, ADD_VALUES, SCATTER_FORWARD, &same)); /* device -> host */
    CHECK(same, "mixed_device_to_host_add");
    PetscCall(Both(sc, scr, y2, x, y2r, xr, ADD_VALUES, SCATTER_REVERSE, &same)); /* host leaves -> device roots */
    CHECK(same, "mixed_host_to_device_reverse_add");
    PetscCall(VecDestroy(&y2));
    PetscCall(VecDestroy(&y2r));
  }
  PetscCall((*counts)(sc, &nat0, &stg0));
  CHECK(isb200 ? (nat0 == native && stg0 == 2) : (nat0 == 0 && stg0 == 0), "mixed_scatters_were_staged");
  PetscCall(VecScatterDestroy(&sc));
  PetscCall(VecScatterDestroy(&scr));
  PetscCall(ISDestroy(&isf));
  PetscCall(ISDestroy(&ist));

  /* stride -> stride (contiguous on one side) */
  PetscCall(ISCreateStride(PETSC_COMM_SELF, 1200, 7, 3, &isf));
  PetscCall(ISCreateStride(PETSC_COMM_SELF, 1200, 100, 1, &ist));
  PetscCall(VecScatterCreate(x, isf, y, ist, &sc));
  PetscCall(VecScatterCreate(xr, isf, yr, ist, &scr));
  PetscCall(Both(sc, scr, x, y, xr, yr, INSERT_VALUES, SCATTER_FORWARD, &same));
  CHECK(same, "stride_forward_insert");
  PetscCall(Both(sc, scr, y, x, yr, xr, ADD_VALUES, SCATTER_REVERSE, &same));
  CHECK(same, "stride_reverse_add");
  PetscCall(VecScatterDestroy(&sc));
  PetscCall(VecScatterDestroy(&scr));
  PetscCall(ISDestroy(&isf));
  PetscCall(ISDestroy(&ist));

  /* whole-vector copy (stride index set in, default out) */
  {
    Vec x2, x2r;
    PetscCall(MakePair(nx, 1, &x2, &x2r));
    PetscCall(ISCreateStride(PETSC_COMM_SELF, nx, 0, 1, &isf));
    PetscCall(VecScatterCreate(x, isf, x2, NULL, &sc));
    PetscCall(VecScatterCreate(xr, isf, x2r, NULL, &scr));
    PetscCall(ISDestroy(&isf));
    PetscCall(Both(sc, scr, x, x2, xr, x2r, INSERT_VALUES, SCATTER_FORWARD, &same));
    CHECK(same, "identity_forward_insert");
    PetscCall(Both(sc, scr, x, x2, xr, x2r, ADD_VALUES, SCATTER_FORWARD, &same));
    CHECK(same, "identity_forward_add");
    PetscCall(VecScatterDestroy(&sc));
    PetscCall(VecScatterDestroy(&scr));
    PetscCall(VecDestroy(&x2));
    PetscCall(VecDestroy(&x2r));
  }

  /* in place: x -> x, disjoint ranges (source == destination buffer: staged through the parent) */
  PetscCall(ISCreateStride(PETSC_COMM_SELF, 1000, 0, 1, &isf));
  PetscCall(ISCreateStride(PETSC_COMM_SELF, 1000, 2000, 1, &ist));
  PetscCall(VecScatterCreate(x, isf, x, ist, &sc));
  PetscCall(VecScatterCreate(xr, isf, xr, ist, &scr));
  PetscCall(VecScatterBegin(sc, x, x, INSERT_VALUES, SCATTER_FORWARD));
  PetscCall(VecScatterEnd(sc, x, x, INSERT_VALUES, SCATTER_FORWARD));
  PetscCall(VecScatterBegin(scr, xr, xr, INSERT_VALUES, SCATTER_FORWARD));
  PetscCall(VecScatterEnd(scr, xr, xr, INSERT_VALUES, SCATTER_FORWARD));
  PetscCall(Same(x, xr, &same));
  CHECK(same, "in_place_insert");
  PetscCall(VecScatterDestroy(&sc));
  PetscCall(VecScatterDestroy(&scr));
  PetscCall(ISDestroy(&isf));
  PetscCall(ISDestroy(&ist));
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&xr));
  PetscCall(VecDestroy(&y));
  PetscCall(VecDestroy(&yr));

  /* block index sets: unit = 3 scalars */
  {
    const PetscInt bs = 3, nbx = 800, nby = 500, nb = 400;
    PetscInt      *bf, *bt;
    PetscCall(MakePair(nbx * bs, bs, &x, &xr));
    PetscCall(MakePair(nby * bs, bs, &y, &yr));
    PetscCall(PetscMalloc2(nb, &bf, nb, &bt));
    for (PetscInt k = 0; k < nb; k++) {
      bf[k] = (PetscInt)(rnd() % (unsigned)nbx);
      if (k % 3 == 2) bf[k] = bf[k - 2];
      bt[k] = perm[k] % nby; /* may repeat: still compared with the host's sequential result */
    }
    PetscCall(ISCreateBlock(PETSC_COMM_SELF, bs, nb, bf, PETSC_COPY_VALUES, &isf));
    PetscCall(ISCreateBlock(PETSC_COMM_SELF, bs, nb, bt, PETSC_COPY_VALUES, &ist));
    PetscCall(VecScatterCreate(x, isf, y, ist, &sc));
    PetscCall(VecScatterCreate(xr, isf, yr, ist, &scr));
    PetscCall(Both(sc, scr, x, y, xr, yr, ADD_VALUES, SCATTER_FORWARD, &same));
    CHECK(same, "block_forward_add_repeated_destinations");
    PetscCall(Both(sc, scr, x, y, xr, yr, INSERT_VALUES, SCATTER_FORWARD, &same));
    CHECK(same, "block_forward_insert_repeated_destinations");
    PetscCall(Both(sc, scr, y, x, yr, xr, ADD_VALUES, SCATTER_REVERSE, &same));
    CHECK(same, "block_reverse_add");
    PetscCall(VecScatterDestroy(&sc));
    PetscCall(VecScatterDestroy(&scr));
    PetscCall(ISDestroy(&isf));
    PetscCall(ISDestroy(&ist));
    PetscCall(PetscFree2(bf, bt));
  }

  /* VecScatterCreateToAll: the gathered vector has the source's type */
  {
    Vec all, allr;
    PetscCall(VecScatterCreateToAll(x, &sc, &all));
    PetscCall(VecScatterCreateToAll(xr, &scr, &allr));
    PetscCall(Both(sc, scr, x, all, xr, allr, INSERT_VALUES, SCATTER_FORWARD, &same));
    CHECK(same, "scatter_to_all");
    PetscCall(VecScatterDestroy(&sc));
    PetscCall(VecScatterDestroy(&scr));
    PetscCall(VecDestroy(&all));
    PetscCall(VecDestroy(&allr));
  }
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&xr));
  PetscCall(VecDestroy(&y));
  PetscCall(VecDestroy(&yr));

  /* a raw PetscSF on PetscInt and PetscScalar buffers that live on the device (b200Malloc), against host buffers */
  {
    const PetscInt nroots = 700, nleaves = 1500, leafspan = 2000;
    PetscSF        sf;
    PetscSFNode   *remote;
    PetscInt      *local, *hr, *hl, *hr0, *hl0;
    PetscScalar   *sr, *sl, *sr0, *sl0;
    void          *d_r = NULL, *d_l = NULL, *d_sr = NULL, *d_sl = NULL;
    b200Handle     h  = NULL;
    PetscMemType   mt = PETSC_MEMTYPE_HOST;
    PetscCall(PetscMalloc2(nleaves, &local, nleaves, &remote));
    PetscCall(PetscMalloc4(nroots, &hr, leafspan, &hl, nroots, &hr0, leafspan, &hl0));
    PetscCall(PetscMalloc4(nroots, &sr, leafspan, &sl, nroots, &sr0, leafspan, &sl0));
    for (PetscInt k = 0; k < nleaves; k++) {
      local[k]        = (k * 4) % leafspan + (k * 4) / leafspan; /* distinct locations in [0, leafspan) */
      remote[k].rank  = 0;
      remote[k].index = (PetscInt)(rnd() % (unsigned)nroots);
    }
    for (PetscInt i = 0; i < nroots; i++) {
      hr0[i] = (PetscInt)(rnd() % 7u) - 3; /* small: products of a root's leaves stay far from overflow */
      sr0[i] = rndval();
    }
    for (PetscInt i = 0; i < leafspan; i++) {
      hl0[i] = (PetscInt)(rnd() % 7u) - 3;
      sl0[i] = rndval();
    }
    PetscCall(PetscSFCreate(PETSC_COMM_SELF, &sf));
    PetscCall(PetscSFSetFromOptions(sf));
    PetscCall(PetscSFSetGraph(sf, nroots, nleaves, local, PETSC_COPY_VALUES, remote, PETSC_COPY_VALUES));
    PetscCall(PetscSFSetUp(sf));
    if (isb200) {
      mt = PETSC_MEMTYPE_CUDA;
      PetscCheck(!b200Create(&h, -1), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200Create");
      PetscCheck(!b200Malloc(h, &d_r, sizeof(PetscInt) * nroots) && !b200Malloc(h, &d_l, sizeof(PetscInt) * leafspan) && !b200Malloc(h, &d_sr, sizeof(PetscScalar) * nroots) && !b200Malloc(h, &d_sl, sizeof(PetscScalar) * leafspan), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200Malloc");
    }
#define UP(d, hsrc, bytes)   PetscCheck(!b200MemcpyHtoD(h, d, hsrc, bytes), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200MemcpyHtoD")
/* the SF kernels run on the plugin's stream, this handle copies on its own: wait for the device first */
#define DOWN(hdst, d, bytes) PetscCheck(!b200DeviceSynchronize() && !b200MemcpyDtoH(h, hdst, d, bytes), PETSC_COMM_SELF, PETSC_ERR_GPU, "b200MemcpyDtoH")
    struct {
      MPI_Op      op;
      const char *name;
    } ops[] = {{MPI_REPLACE, "replace"}, {MPI_SUM, "sum"}, {MPI_MAX, "max"}, {MPI_MIN, "min"}, {MPI_PROD, "prod"}};
    for (int o = 0; o < 5; o++) {
      for (int dir = 0; dir < 2; dir++) {
        char      name[96];
        PetscInt *gr, *gl;
        PetscScalar *gsr, *gsl;
        PetscCall(PetscMalloc4(nroots, &gr, leafspan, &gl, nroots, &gsr, leafspan, &gsl));
        /* host run: plain host pointers -> the parent's loops */
        PetscCall(PetscArraycpy(hr, hr0, nroots));
        PetscCall(PetscArraycpy(hl, hl0, leafspan));
        PetscCall(PetscArraycpy(sr, sr0, nroots));
        PetscCall(PetscArraycpy(sl, sl0, leafspan));
        if (dir == 0) {
          PetscCall(PetscSFBcastBegin(sf, MPIU_INT, hr, hl, ops[o].op));
          PetscCall(PetscSFBcastEnd(sf, MPIU_INT, hr, hl, ops[o].op));
          PetscCall(PetscSFBcastBegin(sf, MPIU_SCALAR, sr, sl, ops[o].op));
          PetscCall(PetscSFBcastEnd(sf, MPIU_SCALAR, sr, sl, ops[o].op));
        } else {
          PetscCall(PetscSFReduceBegin(sf, MPIU_INT, hl, hr, ops[o].op));
          PetscCall(PetscSFReduceEnd(sf, MPIU_INT, hl, hr, ops[o].op));
          PetscCall(PetscSFReduceBegin(sf, MPIU_SCALAR, sl, sr, ops[o].op));
          PetscCall(PetscSFReduceEnd(sf, MPIU_SCALAR, sl, sr, ops[o].op));
        }
        /* device run (or a second host run with -vec_type standard) */
        if (isb200) {
          UP(d_r, hr0, sizeof(PetscInt) * nroots);
          UP(d_l, hl0, sizeof(PetscInt) * leafspan);
          UP(d_sr, sr0, sizeof(PetscScalar) * nroots);
          UP(d_sl, sl0, sizeof(PetscScalar) * leafspan);
        } else {
          PetscCall(PetscArraycpy(gr, hr0, nroots));
          PetscCall(PetscArraycpy(gl, hl0, leafspan));
          PetscCall(PetscArraycpy(gsr, sr0, nroots));
          PetscCall(PetscArraycpy(gsl, sl0, leafspan));
        }
        {
          void *pr = isb200 ? d_r : (void *)gr, *pl = isb200 ? d_l : (void *)gl, *psr = isb200 ? d_sr : (void *)gsr, *psl = isb200 ? d_sl : (void *)gsl;
          if (dir == 0) {
            PetscCall(PetscSFBcastWithMemTypeBegin(sf, MPIU_INT, mt, pr, mt, pl, ops[o].op));
            PetscCall(PetscSFBcastEnd(sf, MPIU_INT, pr, pl, ops[o].op));
            PetscCall(PetscSFBcastWithMemTypeBegin(sf, MPIU_SCALAR, mt, psr, mt, psl, ops[o].op));
            PetscCall(PetscSFBcastEnd(sf, MPIU_SCALAR, psr, psl, ops[o].op));
          } else {
            PetscCall(PetscSFReduceWithMemTypeBegin(sf, MPIU_INT, mt, pl, mt, pr, ops[o].op));
            PetscCall(PetscSFReduceEnd(sf, MPIU_INT, pl, pr, ops[o].op));
            PetscCall(PetscSFReduceWithMemTypeBegin(sf, MPIU_SCALAR, mt, psl, mt, psr, ops[o].op));
            PetscCall(PetscSFReduceEnd(sf, MPIU_SCALAR, psl, psr, ops[o].op));
          }
        }
        if (isb200) {
          DOWN(gr, d_r, sizeof(PetscInt) * nroots);
          DOWN(gl, d_l, sizeof(PetscInt) * leafspan);
          DOWN(gsr, d_sr, sizeof(PetscScalar) * nroots);
          DOWN(gsl, d_sl, sizeof(PetscScalar) * leafspan);
        }
        same = (PetscBool)(!memcmp(gr, hr, sizeof(PetscInt) * nroots) && !memcmp(gl, hl, sizeof(PetscInt) * leafspan) && !memcmp(gsr, sr, sizeof(PetscScalar) * nroots) && !memcmp(gsl, sl, sizeof(PetscScalar) * leafspan));
        PetscCall(PetscSNPrintf(name, sizeof(name), "sf_%s_%s_int_and_scalar", dir ? "reduce" : "bcast", ops[o].name));
        PetscCall(PetscFree4(gr, gl, gsr, gsl));
        CHECK(same, name);
      }
    }
    PetscCall(PetscObjectQueryFunction((PetscObject)sf, "PetscSFB200GetCounts_C", &counts));
    CHECK(counts != NULL, "petscsf_default_type_is_the_b200_subclass");
    PetscCall((*counts)(sf, &native, &staged));
    CHECK(isb200 ? (native == 20 && staged == 0) : (native == 0 && staged == 0), "sf_operations_ran_where_expected");
    /* a new graph on the same SF: the plans follow */
    for (PetscInt k = 0; k < nleaves; k++) remote[k].index = (remote[k].index * 7 + 3) % nroots;
    PetscCall(PetscSFSetGraph(sf, nroots, nleaves, local, PETSC_COPY_VALUES, remote, PETSC_COPY_VALUES));
    PetscCall(PetscArraycpy(sr, sr0, nroots));
    PetscCall(PetscArraycpy(sl, sl0, leafspan));
    PetscCall(PetscSFReduceBegin(sf, MPIU_SCALAR, sl, sr, MPI_SUM));
    PetscCall(PetscSFReduceEnd(sf, MPIU_SCALAR, sl, sr, MPI_SUM));
    if (isb200) {
      PetscScalar *g;
      PetscCall(PetscMalloc1(nroots, &g));
      UP(d_sr, sr0, sizeof(PetscScalar) * nroots);
      UP(d_sl, sl0, sizeof(PetscScalar) * leafspan);
      PetscCall(PetscSFReduceWithMemTypeBegin(sf, MPIU_SCALAR, mt, d_sl, mt, d_sr, MPI_SUM));
      PetscCall(PetscSFReduceEnd(sf, MPIU_SCALAR, d_sl, d_sr, MPI_SUM));
      DOWN(g, d_sr, sizeof(PetscScalar) * nroots);
      same = (PetscBool)!memcmp(g, sr, sizeof(PetscScalar) * nroots);
      PetscCall(PetscFree(g));
      CHECK(same, "sf_new_graph_replans");
      b200Free(h, d_r); b200Free(h, d_l); b200Free(h, d_sr); b200Free(h, d_sl);
      b200Destroy(h);
    }
    PetscCall(PetscSFDestroy(&sf));
    PetscCall(PetscFree2(local, remote));
    PetscCall(PetscFree4(hr, hl, hr0, hl0));
    PetscCall(PetscFree4(sr, sl, sr0, sl0));
  }
  PetscCall(PetscFree3(from, to, perm));
  PetscCall(PetscPrintf(PETSC_COMM_SELF, "all ok\n"));
  PetscCall(PetscFinalize());
  return 0;
}
