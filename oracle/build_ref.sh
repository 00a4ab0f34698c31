#!/bin/bash
# Self-contained build of the UNMODIFIED REFERENCE (PETSc, /root/reference) into baseline/_ref/petsc (git-ignored, travels to the
# GPU box).  It serves three roles: the reference arm of bench.py and the golden-fixture generator run it with its own CPU types
# (-mat_type aij -vec_type standard); and it is the HOST APPLICATION the product is a plugin for: the product arm loads
# petsc_plugin/libpetscb200plugin.so into this same library (-mat_type aijb200 -vec_type b200), so KSPSolve is the
# reference's own gmres.c/cg.c while every Mat/Vec/PC operation of the hot path runs in the sm_100a kernels.
#
#   baseline/_ref/petsc/lib/libpetsc.so*     the reference library: CPU-only, MPIUNI (no MPI in this image), -O2, no -march
#                                          (so PetscSparseDensePlusDot stays the generic FMA-free loop, aij.h:609-614),
#                                          BLAS/LAPACK = the OpenBLAS 0.3.15 bundled with the opencv wheel of this image
#   baseline/_ref/petsc/include/             the GENERATED headers of that build (petscconf.h, petscfix.h, ...): what the
#                                          plugin needs beside /root/reference/include to be ABI-locked to this library
#   baseline/_ref/petsc/build.env            PETSC_DIR / PETSC_ARCH of the scratch build tree (kept for gen_golden.py)
#
# /root/reference is read-only and PETSc configures in-tree, so the tree is copied to a scratch directory under /tmp
# (never into the repo), configured and built there (~1.5 min + ~1 min on 8 cores), and only the products above are
# installed.  No reference SOURCE enters the repository.  Then build_ref_demo.sh compiles the reference's own tutorial
# programs, ref_driver and the plugin against it.  Re-running is a no-op when the library is already there
# (FORCE=1 rebuilds).  Exits 0 without doing anything when /root/reference is absent (the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference
OUT="$(dirname "$HERE")/baseline/_ref/petsc"
SCRATCH="${PETSC_SCRATCH:-/tmp/petsc-ref-build}"
ARCH=arch-ref
BLASDIR=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
[ -d "$REF/src" ] || { echo "no /root/reference here: nothing to build (prebuilt oracle/_ref is used as is)"; exit 0; }
if [ -z "$FORCE" ] && [ -e "$OUT/lib/libpetsc.so" ] && [ -e "$OUT/include/petscconf.h" ] ; then
  echo "reference library already built: $OUT/lib/libpetsc.so"
else
  mkdir -p "$SCRATCH"
  # copy the tree (sources stay outside the repo); --delete keeps a stale scratch honest
  if command -v rsync >/dev/null 2>&1; then rsync -a --delete --exclude "$ARCH" "$REF/" "$SCRATCH/"; else rm -rf "$SCRATCH"; cp -a "$REF" "$SCRATCH"; fi
  chmod -R u+w "$SCRATCH"
  cd "$SCRATCH"
  unset CC CXX FC PETSC_DIR PETSC_ARCH
  export LD_LIBRARY_PATH="$BLASDIR:$LD_LIBRARY_PATH"
  /usr/bin/python3 ./configure PETSC_ARCH=$ARCH --with-cc=/usr/bin/gcc --with-cxx=/usr/bin/g++ \
      --with-mpi=0 --with-fc=0 --with-cuda=0 --with-debugging=0 --with-x=0 \
      --with-blaslapack-lib="[$BLASDIR/libopenblasp-r0-59ffcd50.3.15.so,$BLASDIR/libgfortran-83c28eba.so.5.0.0,$BLASDIR/libquadmath-2284e583.so.0.0.0]" \
      COPTFLAGS=-O2 CXXOPTFLAGS=-O2 > "$SCRATCH/configure.out" 2>&1 || { tail -40 "$SCRATCH/configure.out"; exit 1; }
  make PETSC_DIR="$SCRATCH" PETSC_ARCH=$ARCH all -j"$(nproc)" > "$SCRATCH/make.out" 2>&1 || { tail -40 "$SCRATCH/make.out"; exit 1; }
  rm -rf "$OUT/lib" "$OUT/include"; mkdir -p "$OUT/lib" "$OUT/include" "$OUT/bin"
  real="$(readlink -f "$SCRATCH/$ARCH/lib/libpetsc.so")"
  cp -a "$real" "$OUT/lib/libpetsc.so.3.025"
  ln -sf libpetsc.so.3.025 "$OUT/lib/libpetsc.so"
  cp -a "$SCRATCH/$ARCH/include/." "$OUT/include/"
  printf 'PETSC_DIR=%s\nPETSC_ARCH=%s\n' "$SCRATCH" "$ARCH" > "$OUT/build.env"
  echo "reference library built: $OUT/lib/libpetsc.so ($(du -h "$OUT/lib/libpetsc.so.3.025" | cut -f1))"
fi
bash "$HERE/build_ref_demo.sh"
