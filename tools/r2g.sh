#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_icc_gpu.py -x -q 2>&1 | tail -15
B=baseline/_ref/petsc/bin
echo "== ex2 default PC (ICC) through the plugin vs CPU types"
$B/ex2 -m 20 -n 20 -ksp_monitor -dll_append petsc_plugin/libpetscb200plugin.so -mat_type aijb200 -vec_type b200 -ksp_view 2>&1 | grep -i "residual norm\|Norm of error\|type:\|package\|error" | tail -12
$B/ex2 -m 20 -n 20 -ksp_monitor 2>&1 | grep -i "residual norm\|Norm of error" | tail -4
D=petsc_plugin/b200_driver
echo "== random CSR with auto column blocks (default) vs off"
for d in 32 128; do
timeout 300 $D -bench rand -rand_n 10000000 -rand_d $d -options_left 0 2>&1 | grep B200JSON | cut -c1-300
timeout 300 $D -bench rand -rand_n 10000000 -rand_d $d -mat_b200_spmv_column_blocks 0 -options_left 0 2>&1 | grep B200JSON | cut -c1-300
done
echo "== cg27 256 with ICC on the device"
timeout 300 $D -bench cg27 -n 256 -pc_type icc -pc_factor_mat_solver_type b200 -options_left 0 2>&1 | grep "B200JSON\|ERROR" | cut -c1-700
