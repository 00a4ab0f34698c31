#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
mkdir -p gpurun_out
D=petsc_plugin/b200_driver
ncu --clock-control none --set full --import-source on -f -k regex:"ilu_sweep_pipe_kernel" -c 1 -o gpurun_out/r2p_pipe27 $D -bench cg27 -n 256 -ksp_max_it 2 -options_left 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2p_pipe27.ncu-rep | cut -c1-420
ncu -i gpurun_out/r2p_pipe27.ncu-rep --page source --csv > gpurun_out/r2p_pipe27_source.csv 2>/dev/null
ncu -i gpurun_out/r2p_pipe27.ncu-rep --page details --csv > gpurun_out/r2p_pipe27_details.csv 2>/dev/null
rm -f gpurun_out/r2p_pipe27.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_pipe27_source.csv')))
hdr=rows[0]
print(hdr[:12])
# find sampling column
idx=[i for i,h in enumerate(hdr) if 'Sampl' in h]
print([hdr[i] for i in idx])
if idx:
    k=idx[0]
    def val(r):
        try: return float(r[k].replace(',',''))
        except: return 0.0
    tot=sum(val(r) for r in rows[1:])
    top=sorted(rows[1:], key=val, reverse=True)[:40]
    for r in top:
        print("%6.2f%%"%(100*val(r)/max(tot,1)), ' | '.join(x[:70] for x in r[:4]))
PY
