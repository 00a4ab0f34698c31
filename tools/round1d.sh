# final batch of round 1: whole GPU suite, config-5 column blocks + PIPECG timings, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30
timeout 400 python tools/bench_configs.py --what 5,nb,p --out gpurun_out/configs_r1d.json 2>&1 | grep -E "^(config5|pipecg)|Error|error" | cut -c1-1500
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r1d.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches'],d['clocks'])"; tail -2 gpurun_out/bench_r1d.err
