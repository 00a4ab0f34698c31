#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
timeout 600 python -m pytest tests/test_host_gpu.py tests/test_fullsize_gpu.py -x -q -k "ilu or pc_apply or cg or bjacobi or ex2 or ksp" 2>&1 | tail -15
D=petsc_plugin/b200_driver
for m in 1 0; do
echo "== MARCH=$m"
PETSCB200_ILU_MARCH=$m timeout 300 $D -bench cg27 -n 256 -options_left 0 2>&1 | grep B200JSON | cut -c1-700
PETSCB200_ILU_MARCH=$m timeout 300 $D -bench gmres7 -n 384 -steps 2 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 2>&1 | grep B200JSON | cut -c1-500
done
