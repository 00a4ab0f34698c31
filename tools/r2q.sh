#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
D=petsc_plugin/b200_driver
run() { timeout 300 $D -bench cg27 -n 256 -ksp_max_it 20 -options_left 0 2>&1 | grep B200JSON | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[9:]); print(d['pcapply_ilu_ms'])"; }
for b in 0 40 100 200 400 1000; do for la in 1 2 4; do
  echo -n "backoff=$b lookahead=$la pcapply_ms="; PETSCB200_ILU_BACKOFF_NS=$b PETSCB200_ILU_LOOKAHEAD=$la run
done; done
