#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
B=baseline/_ref/petsc/bin; P="-dll_append petsc_plugin/libpetscb200plugin.so"; O=gpurun_out/r3e.log; : > $O
echo "=== coherence_driver" >> $O; timeout 20 $B/coherence_driver -mat_type aijb200 -vec_type b200 $P 2>&1 | grep -v "^WARNING\|unused\|Option left\|spelling" >> $O
echo "=== ex9" >> $O; timeout 15 $B/reftests/ksp_ksp_tutorials_ex9 -t 2 -pc_type jacobi -ksp_monitor -ksp_type gmres -ksp_gmres_cgs_refinement_type refine_always -s2_ksp_type bcgs -s2_pc_type jacobi -s2_ksp_monitor -mat_type aijb200 -vec_type b200 $P 2>&1 | sed -n 15,18p >> $O
echo "=== ex254" >> $O; timeout 15 $B/reftests/mat_tests_ex254 -ncoos 3 -mat_type aijb200 $P >> $O 2>&1; echo "rc $?" >> $O
echo "=== pgmresb200 vs pgmres (ex2 30x30)" >> $O; timeout 15 $B/ex2 -m 30 -n 30 -pc_type jacobi -ksp_rtol 1e-8 -ksp_type pgmresb200 -mat_type aijb200 -vec_type b200 $P 2>&1 | tail -1 >> $O; timeout 15 $B/ex2 -m 30 -n 30 -pc_type jacobi -ksp_rtol 1e-8 -ksp_type pgmres 2>&1 | tail -1 >> $O
cat $O | head -40
