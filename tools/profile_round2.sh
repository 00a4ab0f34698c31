#!/bin/bash
# ncu evidence for the kernels AS SHIPPED (round 2): a launch list of the bench command + --set full captures per kernel family.
# Run on the GPU box:  bash tools/profile_round2.sh     (writes gpurun_out/r2_*.ncu-rep / .csv; summarise with tools/ncu_summary.py)
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
mkdir -p gpurun_out
D=petsc_plugin/b200_driver
NCU="ncu --clock-control none"
# 1. launch list of the bench command (per-launch durations; cold-cache and serialised: shares, not absolutes)
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
FULL="$NCU --set full --import-source on -f"
# 2. the GMRES(30)+Jacobi cycle at 512^3: SpMV(+Jacobi), MDot and MAXPY late in the cycle (nv ~ 25-30), reductions, scale
$FULL -k regex:csr_spmv_tile_kernel -c 2 -o gpurun_out/r2_spmv7 $D -bench gmres7 -n 512 -steps 1 -warmup 1 -kernels 0 -options_left 0 > /dev/null 2>&1
$FULL -k regex:mdot_kernel --launch-skip 24 -c 3 -o gpurun_out/r2_mdot $D -bench gmres7 -n 512 -steps 1 -warmup 1 -kernels 0 -options_left 0 > /dev/null 2>&1
$FULL -k regex:maxpy_kernel --launch-skip 24 -c 3 -o gpurun_out/r2_maxpy $D -bench gmres7 -n 512 -steps 1 -warmup 1 -kernels 0 -options_left 0 > /dev/null 2>&1
$FULL -k regex:"ew_kernel|reduce_kernel|axpy_dot" -c 6 -o gpurun_out/r2_blas1 $D -bench gmres7 -n 512 -steps 1 -warmup 1 -kernels 0 -b200_keep_pcjacobi -options_left 0 > /dev/null 2>&1
# 3. 27-point 256^3: SpMV (4 lanes/row), ILU(0) numeric + level-scheduled sweeps; ICC(0) numeric + marching sweeps
$FULL -k regex:"csr_spmv_tile_kernel|ilu_numeric_kernel|ilu_sweep_pipe_kernel" -c 6 -o gpurun_out/r2_cg27_ilu $D -bench cg27 -n 256 -ksp_max_it 3 -options_left 0 > /dev/null 2>&1
$FULL -k regex:"icc_numeric_kernel|icc_gather_kernel|sweep_march_kernel" -c 5 -o gpurun_out/r2_cg27_icc $D -bench cg27 -n 256 -pc_type icc -ksp_max_it 3 -options_left 0 > /dev/null 2>&1
# 4. 7-point ILU(0) sweeps (the per-rank block of config 4)
$FULL -k regex:"ilu_sweep_pipe_kernel" -c 2 -o gpurun_out/r2_ilu7 $D -bench gmres7 -n 384 -steps 1 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 > /dev/null 2>&1
# 5. random CSR d = 32 (column-blocked passes, chosen automatically) and d = 5
$FULL -k regex:"csr_spmv_vector_kernel|csr_spmv_tile_kernel" --launch-skip 4 -c 4 -o gpurun_out/r2_rand32 $D -bench rand -rand_n 10000000 -rand_d 32 -options_left 0 > /dev/null 2>&1
$FULL -k regex:"csr_spmv_vector_kernel|csr_spmv_tile_kernel" --launch-skip 2 -c 2 -o gpurun_out/r2_rand5 $D -bench rand -rand_n 10000000 -rand_d 5 -options_left 0 > /dev/null 2>&1
# 6. pipecgb200's fused recurrence kernel and the compressed-row / assembly helpers on small cases
$FULL -k regex:"pipecg_update_kernel" -c 2 -o gpurun_out/r2_pipecg $D -bench cg27 -n 256 -ksp_type pipecgb200 -pc_type jacobi -ksp_max_it 4 -options_left 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_*.ncu-rep > gpurun_out/r2_ncu_summary.txt 2> gpurun_out/r2_ncu_summary.err
ncu -i gpurun_out/r2_spmv7.ncu-rep --page details --csv > gpurun_out/r2_spmv7_details.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
# gpurun brings back at most 64 MiB: keep the SpMV report (the roofline kernel) and the summaries, drop the other raw reports
for f in gpurun_out/r2_*.ncu-rep; do [ "$f" = gpurun_out/r2_spmv7.ncu-rep ] || rm -f "$f"; done
cat gpurun_out/r2_ncu_summary.txt | cut -c1-330
