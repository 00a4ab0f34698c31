#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_petsc_driver_gpu.py tests/test_petsc_plugin_gpu.py -x -q -rs 2>&1 | tail -30 > gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_pytest.log
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['value'], d['e2e'])"; tail -5 gpurun_out/r2d_bench.err
