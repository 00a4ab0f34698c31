#!/bin/bash
# Builds the reference-side demo artefacts into oracle/_ref/ (git-ignored; they travel to the GPU box with gpurun):
#   (baseline/_ref/petsc/lib/libpetsc.so*    the reference library, built by oracle/build_ref.sh from /root/reference: CPU-only, MPIUNI, -O2)
#   baseline/_ref/petsc/bin/ex2              the reference's own tutorial programs, compiled from the sources where they lie
#   baseline/_ref/petsc/bin/bench_kspsolve   under /root/reference/src/ksp/ksp/tutorials/ (never copied into the repo)
#   oracle/_ref/ref_driver                 oracle/ref_driver.c (our driver against the reference's public API)
#   petsc_plugin/libpetscb200plugin.so     the plugin, built against exactly this PETSc
#   baseline/_ref/petsc/bin/plugin_driver    petsc_plugin/plugin_driver.c (our PETSc program for device COO / transposed products)
#   baseline/_ref/petsc/bin/sf_driver        petsc_plugin/sf_driver.c (our PETSc program for VecScatter / PetscSF on device data)
#   petsc_plugin/b200_driver, libb200driver.so   petsc_plugin/b200_driver.c (our PETSc program for the BASELINE workloads)
# Only runs in the build container (needs /root/reference and the configured PETSc build).  No reference SOURCE is copied.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF=/root/reference
BLASDIR=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
OUT="$(dirname "$HERE")/baseline/_ref/petsc"; mkdir -p "$OUT/bin" "$HERE/_ref"
[ -e "$OUT/lib/libpetsc.so" ] && [ -e "$OUT/include/petscconf.h" ] && [ -d "$REF/include" ] || { echo "no reference library in $OUT (run oracle/build_ref.sh in the build container): skipping the reference demo build"; exit 0; }
# headers: the reference's own include tree where it lies + the generated headers build_ref.sh installed next to the library
INC="-I$REF/include -I$OUT/include"
LNK="-L$OUT/lib -lpetsc -Wl,-rpath,\$ORIGIN/../lib -Wl,-rpath,$BLASDIR -Wl,-rpath-link,$BLASDIR -Wl,--allow-shlib-undefined -lm"
for ex in ex2 bench_kspsolve; do
  /usr/bin/gcc -O2 -o "$OUT/bin/$ex" "$REF/src/ksp/ksp/tutorials/$ex.c" $INC $LNK
done
/usr/bin/gcc -O2 -ffp-contract=off -fopenmp -o "$HERE/_ref/ref_driver" "$HERE/ref_driver.c" "$HERE/oracle.c" -I"$HERE" $INC -L$OUT/lib -lpetsc -Wl,-rpath,\$ORIGIN/../../baseline/_ref/petsc/lib -Wl,-rpath,$BLASDIR -Wl,-rpath-link,$BLASDIR -Wl,--allow-shlib-undefined -lm
make -s -C "$ROOT/petsc_plugin" PETSC_INC="$INC" PETSC_LIBDIR="$OUT/lib"
# a PETSc program for the plugin paths the tutorials do not reach (device COO, MatMultTranspose, MatBindToCPU)
/usr/bin/gcc -O2 -o "$OUT/bin/plugin_driver" "$ROOT/petsc_plugin/plugin_driver.c" $INC -I"$ROOT/include" $LNK -L"$ROOT/petsc_b200/lib" -lpetscb200 -Wl,-rpath,\$ORIGIN/../../../../petsc_b200/lib
# VecScatter / PetscSF on device vectors against the host types (tests/test_petsc_plugin_{cpu,gpu}.py)
/usr/bin/gcc -O2 -std=gnu11 -Wall -o "$OUT/bin/sf_driver" "$ROOT/petsc_plugin/sf_driver.c" $INC -I"$ROOT/include" $LNK -L"$ROOT/petsc_b200/lib" -lpetscb200 -Wl,-rpath,\$ORIGIN/../../../../petsc_b200/lib
# one check per host/device coherence rule of the plugin (found with the reference's own programs)
/usr/bin/gcc -O2 -std=gnu11 -Wall -o "$OUT/bin/coherence_driver" "$ROOT/petsc_plugin/coherence_driver.c" $INC $LNK
# the PETSc program that runs the BASELINE workloads on the b200 types (bench.py, tools/, GPU tests): executable + shared object
DRV_LNK="-L$ROOT/petsc_plugin -lpetscb200plugin -L$ROOT/petsc_b200/lib -lpetscb200 -L$OUT/lib -lpetsc -Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../petsc_b200/lib -Wl,-rpath,\$ORIGIN/../baseline/_ref/petsc/lib -Wl,-rpath,$BLASDIR -Wl,-rpath-link,$BLASDIR -Wl,--allow-shlib-undefined -lm"
/usr/bin/gcc -O2 -g -std=gnu11 -Wall -Wno-unused-parameter -Wno-format-truncation -o "$ROOT/petsc_plugin/b200_driver" "$ROOT/petsc_plugin/b200_driver.c" $INC -I"$ROOT/include" $DRV_LNK
/usr/bin/gcc -O2 -g -std=gnu11 -fPIC -shared -Wno-format-truncation -DB200_DRIVER_NO_MAIN -o "$ROOT/petsc_plugin/libb200driver.so" "$ROOT/petsc_plugin/b200_driver.c" $INC -I"$ROOT/include" $DRV_LNK
# the reference's own device-variant test programs (tools/ref_conformance.py: compiled from the sources where they lie + manifest)
python "$ROOT/tools/ref_conformance.py" build | tail -1
echo "reference demo built in $OUT"
