#!/bin/bash
# final-round confirmation: GPU suite (incl. the PetscSF type), smoke, the indexed-scatter microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/r3a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3a_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r3a_smoke.log
LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs timeout 120 baseline/_ref/petsc/bin/sf_driver -dll_append petsc_plugin/libpetscb200plugin.so -mat_type aijb200 -vec_type b200 > gpurun_out/r3a_sf_driver.log 2>&1; echo "sf_driver rc $?"; grep -c "^ok" gpurun_out/r3a_sf_driver.log; grep -v "^ok" gpurun_out/r3a_sf_driver.log | head -20
timeout 300 python tools/sf_bench.py --n 33554432 --out gpurun_out/r3a_sf_bench.json > gpurun_out/r3a_sf_bench.log 2>&1; echo "sf_bench rc $?"; cut -c1-220 gpurun_out/r3a_sf_bench.log | tail -8
