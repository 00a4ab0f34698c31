#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -4
python tools/kernel_to_beat.py --out gpurun_out/r2s_kernel_to_beat.json --md gpurun_out/r2s_kernel_to_beat.md 2>&1 | grep "power-law" | cut -c1-500
cat gpurun_out/r2s_kernel_to_beat.md
