#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
B=baseline/_ref/petsc/bin/reftests; P="-dll_append petsc_plugin/libpetscb200plugin.so"; O=gpurun_out/r3d.log; : > $O
E9="-t 2 -pc_type jacobi -ksp_monitor -ksp_type gmres -ksp_gmres_cgs_refinement_type refine_always -s2_ksp_type bcgs -s2_pc_type jacobi -s2_ksp_monitor -mat_type aijb200 -vec_type b200"
for extra in "" "-b200_keep_pcjacobi" "-s2_pc_jacobi_b200_fuse 0" "-vec_type standard"; do echo "=== ex9 $extra" >> $O; timeout 20 $B/ksp_ksp_tutorials_ex9 $E9 $extra $P 2>&1 | sed -n 15,21p >> $O; done
echo "=== ex254 b200" >> $O; timeout 20 $B/mat_tests_ex254 -ncoos 3 -mat_type aijb200 $P >> $O 2>&1
echo "=== ex28 icc host" >> $O; timeout 20 $B/mat_tests_ex28 -mat_solver_type petsc -mat_type aij -mat_factor_type icc $P >> $O 2>&1
echo "=== ex28 icc b200" >> $O; timeout 20 $B/mat_tests_ex28 -mat_solver_type b200 -mat_type aijb200 -mat_factor_type icc $P >> $O 2>&1
echo "=== ex217 b200" >> $O; timeout 20 $B/mat_tests_ex217 -mat_type aijb200 $P 2>&1 | head -30 >> $O
echo "=== ex132 host" >> $O; timeout 20 $B/mat_tests_ex132 -view -mat_type aij $P 2>&1 | head -70 >> $O
echo "=== ex132 b200" >> $O; timeout 20 $B/mat_tests_ex132 -view -mat_type aijb200 $P 2>&1 | head -90 >> $O
wc -l $O
