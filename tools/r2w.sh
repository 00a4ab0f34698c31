#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
D=petsc_plugin/b200_driver
run() { timeout 300 $D -bench cg27 -n 256 -ksp_max_it 20 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print(d['pcapply_ilu_ms'])"; }
for c in 1 2 3 4 5; do echo -n "ctas_per_sm=$c pcapply_ms="; PETSCB200_ILU_CTAS_PER_SM=$c run; done
