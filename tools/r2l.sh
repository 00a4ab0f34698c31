#!/bin/bash
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_n$N.json')); print('value',d['value'],'ms/step',d['ms_per_step'],'parity',d.get('parity_check',{}).get('passed'),d.get('parity_check',{}).get('failed'),'e2e',d['e2e'] and d['e2e']['value']); print(json.dumps(d.get('configs'))[:1500])"
tail -5 gpurun_out/r2l_bench_n$N.err | cut -c1-300
