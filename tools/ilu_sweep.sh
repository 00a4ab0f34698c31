set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "ilu" 2>&1 | tail -3
for M in 1 2 4; do
  PETSCB200_ILU_ROWS_PER_GROUP=$M timeout 300 python tools/bench_configs.py --what 3,4 --n27 256 --n7 320 --out gpurun_out/ilu_M$M.json 2>&1 | grep -E "^config" | cut -c1-700
done
