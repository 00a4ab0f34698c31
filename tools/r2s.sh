#!/bin/bash
# final evidence run, 1 GPU: full GPU suite, smoke, bench (product arm with configs + cpu_baseline), reference arm, kernel-to-beat table
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs 2>&1 | tail -8 > gpurun_out/r2s_pytest.log; cat gpurun_out/r2s_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n1.json')); print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['kind']); print({k:(v.get('ms') or v.get('ms_per_iteration') or v.get('us_per_iteration')) for k,v in d['configs'].items() if isinstance(v,dict)}); print(d['configs'].get('error'))"
python tools/kernel_to_beat.py --out gpurun_out/r2s_kernel_to_beat.json --md gpurun_out/r2s_kernel_to_beat.md 2>&1 | grep "power-law\|random d=32" | cut -c1-400
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err; cut -c1-900 gpurun_out/r2s_bench_reference.json
