#!/bin/bash
mkdir -p gpurun_out
python tools/kernel_to_beat.py --out gpurun_out/r2_kernel_to_beat.json --md gpurun_out/r2_kernel_to_beat.md 2>&1 | tail -20
bash tools/profile_round2.sh 2>&1 | tail -15
