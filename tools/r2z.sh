#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2z_bench_n4.json 2> gpurun_out/r2z_bench_n4.err
python -c "
import json; d=json.load(open('gpurun_out/r2z_bench_n4.json')); print('value',d['value'],'ms/step',d['ms_per_step'],'parity',d['parity_check']['passed'],d['parity_check']['failed'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'numa',d['numa_binding'])"
tail -2 gpurun_out/r2z_bench_n4.err | cut -c1-200
