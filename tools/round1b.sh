# second measurement batch of round 1: config 5 with the persisting-L2 window, transposed product, COO assembly, e2e phases
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py --what 5,tr,coo --out gpurun_out/configs_r1b.json 2>&1 | grep -E "^(config5|transpose|coo)|Error|error" | cut -c1-1200
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 3000 gpurun_out/bench_r1b.json; tail -5 gpurun_out/bench_r1b.err
