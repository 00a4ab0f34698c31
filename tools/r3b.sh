#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sf_gpu.py tests/test_petsc_plugin_gpu.py -m gpu -q -rs -k "sf or indexed" > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc $?"
tail -4 gpurun_out/r3b_pytest.log
timeout 300 python tools/sf_bench.py --n 33554432 --out gpurun_out/r3b_sf_bench.json > gpurun_out/r3b_sf_bench.log 2>&1; echo "sf_bench rc $?"; cut -c1-220 gpurun_out/r3b_sf_bench.log | tail -8
