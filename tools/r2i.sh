#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_petsc_driver_gpu.py -x -q -k "history" 2>&1 | tail -8
D=petsc_plugin/b200_driver
for n in 48 128 256; do for k in cg pipecg pipecgb200; do
 echo "n=$n ksp=$k"; timeout 300 $D -bench cg27 -n $n -ksp_type $k -pc_type jacobi -options_left 0 2>&1 | grep "B200JSON\|ERROR" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('B200JSON'):
        d=json.loads(l[9:]); print({k:d[k] for k in ('iterations','reason','solve_ms','ms_per_iteration','max_error')})
    else: print(l.strip()[:300])"
done; done 2>&1 | tee gpurun_out/r2i_pipecg.log
