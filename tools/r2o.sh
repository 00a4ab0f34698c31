#!/bin/bash
timeout 900 python -m pytest tests/test_ilu_variants_gpu.py -x -q 2>&1 | tail -6
