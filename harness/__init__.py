"""TEST HARNESS -- not the product.

A stand-alone C mini-PETSc (harness/host/*.c -> harness/libb200harness.so, declared in harness/petscb200_host.h) that mimics
the slice of PETSc's Vec/Mat/PC/KSP interface the Krylov hot path uses, on top of the product's C ABI (include/petscb200.h).
It exists so that pytest can drive every kernel through PETSc-shaped calls on a box that has no PETSc, and so that the
multi-rank tests have a second, independent host path to compare with.  Its Krylov callers (host/ksp.c, host/pc.c) restate
the reference's gmres.c / cg.c / precon.c control flow; they are NOT what bench.py measures and not part of the drop-in:
the measured path is the reference's own KSPSolve in libpetsc driving petsc_plugin/libpetscb200plugin.so.
"""
