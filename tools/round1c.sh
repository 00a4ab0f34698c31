# third batch of round 1: the whole GPU suite (ordered row sums, full-size property tests) + summation-mode timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 500 python tools/bench_configs.py --what 5,tr --nasm 256 --out gpurun_out/configs_r1c.json 2>&1 | grep -E "^(config5|transpose)|Error|error" | cut -c1-900
