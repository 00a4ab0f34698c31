#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
D=petsc_plugin/b200_driver
run() { timeout 300 $D -bench rand -rand_n $1 -rand_d $2 $3 -options_left 0 2>&1 | grep "B200JSON\|ERROR" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('B200JSON'):
        d=json.loads(l[9:]); print(d['d'], round(d['ms'],4), 'ms', round(d['algorithmic_bytes']/d['ms']/1e6), 'GB/s')
    else: print(l.strip()[:200])"; }
for d in 5 32 128; do
  echo "== d=$d auto"; run 10000000 $d ""
  echo "   auto, no blocks"; run 10000000 $d "-mat_b200_spmv_column_blocks 0"
  for v in 2 4 8 16 32; do echo "   vector=$v no blocks"; PETSCB200_SPMV_VECTOR=$v run 10000000 $d "-mat_b200_spmv_column_blocks 0"; done
  echo "   vector auto + 2 blocks"; run 10000000 $d "-mat_b200_spmv_column_blocks 2"
done
echo "== d=512 auto"; run 2500000 512 ""
for v in 16 32; do echo "   vector=$v"; PETSCB200_SPMV_VECTOR=$v run 2500000 512 ""; done
echo "== sanity: 7-pt/27-pt unchanged"
timeout 300 $D -bench cg27 -n 128 -ksp_max_it 5 -options_left 0 2>&1 | grep B200JSON | cut -c1-300
python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
