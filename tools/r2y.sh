#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
D=petsc_plugin/b200_driver
timeout 600 python -m pytest tests/test_ilu_variants_gpu.py tests/test_host_gpu.py -x -q -k "ilu or pc_apply or variants or packed" 2>&1 | tail -3
timeout 300 $D -bench cg27 -n 256 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print('cg27', {k:d[k] for k in ('iterations','ms_per_iteration','pcapply_ilu_ms','iterations_per_sec')})"
PETSCB200_ILU_PACKED=0 timeout 300 $D -bench gmres7 -n 512 -steps 2 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print('gmres7 512 ilu (pipe)', {k:d[k] for k in ('ms_per_step','iterations_per_sec')})"
