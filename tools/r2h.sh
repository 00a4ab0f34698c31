#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
