#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
mkdir -p gpurun_out
D=petsc_plugin/b200_driver
ncu --clock-control none --set full --import-source on -f -k regex:"ilu_sweep_packed_kernel" -c 2 -o gpurun_out/r2n_packed $D -bench gmres7 -n 512 -steps 1 -warmup 1 -kernels 0 -pc_type ilu -pc_factor_mat_solver_type b200 -options_left 0 > /dev/null 2>&1
ncu -i gpurun_out/r2n_packed.ncu-rep --page details --csv > gpurun_out/r2n_packed_details.csv
python tools/ncu_summary.py gpurun_out/r2n_packed.ncu-rep | cut -c1-400
rm -f gpurun_out/r2n_packed.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2n_packed_details.csv')))
hdr=rows[0]; c={h:i for i,h in enumerate(hdr)}
seen=set()
for r in rows[1:]:
    if len(r)<len(hdr): continue
    if r[c['ID']]!='0': continue
    name=r[c['Metric Name']]; sec=r[c['Section Name']]
    if any(k in name for k in ('Throughput','Duration','Hit Rate','Sectors','Occupancy','Registers','Eligible','Issued','Stall','Active','L2','DRAM','Mem')):
        print(sec[:28].ljust(28), name[:60].ljust(60), r[c['Metric Unit']], r[c['Metric Value']])
PY
