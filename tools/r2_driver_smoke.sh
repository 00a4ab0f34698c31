#!/bin/bash
# smoke + first numbers of petsc_plugin/b200_driver (real PETSc + plugin) on one GPU
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
D=petsc_plugin/b200_driver
mkdir -p gpurun_out
set -x
$D -bench gmres7 -n 64 -steps 2 -warmup 1 -e2e 1 -ksp_monitor_cancel 2>&1 | tail -5
$D -bench cg27 -n 32 2>&1 | tail -3
$D -bench rand -rand_n 100000 -rand_d 8 2>&1 | tail -3
$D -bench ex2 -m 20 -n 20 -ksp_type gmres -pc_type jacobi 2>&1 | tail -3
$D -bench ex2 -m 100 -n 100 -ksp_type gmres -pc_type jacobi 2>&1 | tail -3
$D -bench gmres7 -n 128 -steps 2 -warmup 1 -pc_type bjacobi -sub_pc_type ilu -sub_pc_factor_mat_solver_type b200 -kernels 0 2>&1 | tail -3
$D -bench gmres7 -n 512 -steps 5 -warmup 3 -e2e 1 -json_out gpurun_out/r2b_gmres7.jsonl 2>&1 | tail -6
$D -bench gmres7 -n 512 -steps 5 -warmup 3 -kernels 0 -b200_keep_pcjacobi -json_out gpurun_out/r2b_gmres7_unfused.jsonl 2>&1 | tail -3
$D -bench cg27 -n 256 -json_out gpurun_out/r2b_cg27.jsonl 2>&1 | tail -3
for d in 5 32 128; do $D -bench rand -rand_n 10000000 -rand_d $d -json_out gpurun_out/r2b_rand.jsonl 2>&1 | tail -2; done
$D -bench rand -rand_n 2500000 -rand_d 512 -json_out gpurun_out/r2b_rand.jsonl 2>&1 | tail -2
