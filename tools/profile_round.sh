#!/bin/bash
# ncu captures for profiles/ (run under gpurun, ONE GPU).  Numbers printed by runs under ncu are never bench values.
# usage: tools/profile_round.sh <round-tag>
R=${1:-r1}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
# 1. every launch of one bench step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_$R.csv $B > $O/ncu_launches_$R.log 2>&1
# 2. the dominant kernel (CSR SpMV), full set, with source
ncu --set full --clock-control none --import-source on -k regex:csr_spmv_tile -s 10 -c 2 -f -o $O/prof_spmv_$R $B > $O/ncu_spmv_$R.log 2>&1
# 3. MDot / MAXPY at nv = 30
ncu --set full --clock-control none --import-source on -k regex:maxpy_kernel -s 29 -c 1 -f -o $O/prof_maxpy30_$R $B > $O/ncu_maxpy_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdot_kernel -s 30 -c 1 -f -o $O/prof_mdot30_$R $B > $O/ncu_mdot_$R.log 2>&1
# 4. ILU(0) sweeps (27-point 128^3) and the random-CSR SpMV (d = 32)
ncu --set full --clock-control none --import-source on -k regex:ilu_sweep -s 4 -c 2 -f -o $O/prof_ilusweep_$R python tools/bench_configs.py --what 3 --n27 128 > $O/ncu_ilu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:csr_spmv_tile -s 12 -c 1 -f -o $O/prof_spmv_rand32_$R python tools/bench_configs.py --what 5 --nrand 4000000 > $O/ncu_rand_$R.log 2>&1
ls -la $O | grep $R
