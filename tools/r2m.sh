#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_host_gpu.py tests/test_fullsize_gpu.py tests/test_petsc_driver_gpu.py -x -q -k "ilu or pc_apply or cg or bjacobi or ex2 or ksp or history" 2>&1 | tail -6
D=petsc_plugin/b200_driver
for m in 1 0; do
echo "== PACKED=$m"
PETSCB200_ILU_PACKED=$m timeout 300 $D -bench cg27 -n 256 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print('cg27', {k:d[k] for k in ('iterations','ms_per_iteration','pcapply_ilu_ms','first_solve_incl_setup_s')})"
PETSCB200_ILU_PACKED=$m timeout 300 $D -bench gmres7 -n 512 -steps 2 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print('gmres7 512 ilu', {k:d[k] for k in ('ms_per_step','iterations_per_sec','rnorm')})"
done
