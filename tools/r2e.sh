#!/bin/bash
# 2-GPU: multi-rank parity (harness worker + real-PETSc plugin inside bench.py) and the weak-scaling step
mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -x -q -rs 2>&1 | tail -15 > gpurun_out/r2e_pytest_multigpu.log; cat gpurun_out/r2e_pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_n2.json')); print(d['value'], d['ms_per_step'], d.get('parity_check'), d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['setup_ms']))"; tail -30 gpurun_out/r2e_bench_n2.err | cut -c1-400
