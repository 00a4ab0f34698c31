#!/bin/bash
mkdir -p gpurun_out
timeout 260 python tools/ref_conformance.py run --json gpurun_out/r3c_ref_conformance.json --md gpurun_out/r3c_ref_conformance.md > gpurun_out/r3c_ref_conformance.log 2>&1; echo "conformance rc $?"
tail -4 gpurun_out/r3c_ref_conformance.log | cut -c1-300
grep -c " pass" gpurun_out/r3c_ref_conformance.log; grep " FAIL" gpurun_out/r3c_ref_conformance.log | cut -c1-260 | head -40
