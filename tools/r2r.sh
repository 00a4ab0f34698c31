#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -8
