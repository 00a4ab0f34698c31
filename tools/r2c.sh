#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -30 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 6000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
