#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs 2>&1 | tail -6 > gpurun_out/r2u_pytest.log; cat gpurun_out/r2u_pytest.log
D=petsc_plugin/b200_driver
FULL="ncu --clock-control none --set full -f"
$FULL -k regex:"ilu_numeric_kernel|ilu_pack_vals_kernel|ilu_pack_cols_kernel" -c 4 -o gpurun_out/r2u_ilunum $D -bench gmres7 -n 384 -steps 1 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 > /dev/null 2>&1
$FULL -k regex:"ilu_numeric_kernel" -c 1 -o gpurun_out/r2u_ilunum27 $D -bench cg27 -n 256 -ksp_max_it 1 -options_left 0 > /dev/null 2>&1
$FULL -k regex:"coo_|tr_|blk_" -c 10 -o gpurun_out/r2u_asm python tools/bench_configs.py --what tr,coo --nasm 160 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2u_ilunum.ncu-rep gpurun_out/r2u_ilunum27.ncu-rep gpurun_out/r2u_asm.ncu-rep > gpurun_out/r2u_ncu_summary.txt 2>/dev/null
rm -f gpurun_out/r2u_*.ncu-rep
cut -c1-330 gpurun_out/r2u_ncu_summary.txt
